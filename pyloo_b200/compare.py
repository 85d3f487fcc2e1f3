"""``loo_compare`` -- rank models by ELPD and compute model weights.

Drop-in for ``pyloo.loo_compare`` with ``ic in {"loo", "waic"}`` and ``observations=None``
(reference: pyloo/compare.py:23-264, ``_calculate_ics`` :285-456, ``_stacking_weights`` :477-536,
``_bb_pseudo_bma_weights`` :539-577, ``_pseudo_bma_weights`` :580-596).  Per-model pointwise ELPDs
come from the GPU path (``loo`` / ``waic`` with ``pointwise=True``); the K-dimensional weight
optimisation stays on the host with the same SciPy SLSQP call as the reference, so weights agree.
``ic="kfold"`` and subsampled comparison need PyMC refits and are out of scope.
"""

from __future__ import annotations

import warnings
from copy import deepcopy

import numpy as np
import pandas as pd
import scipy.stats as st
from scipy import optimize

from .elpd import ELPDData
from .loo import loo
from .waic import waic

__all__ = ["loo_compare"]

_SCALES = ("log", "negative_log", "deviance")
_METHODS = ("stacking", "bb-pseudo-bma", "pseudo-bma")


def _to_log_scale(values, scale):
    """Undo the output scale on a *copy* (compare.py:489-492, :556-559, :587-590)."""
    values = np.array(values, dtype=float, copy=True)
    if scale == "deviance":
        values /= -2
    elif scale == "negative_log":
        values *= -1
    return values


def _pointwise_matrix(elpds, ic_i):
    names = list(elpds)
    cols = [np.asarray(elpds[n][ic_i].values, dtype=float).ravel() for n in names]
    if any(len(c) != len(cols[0]) for c in cols):
        raise ValueError("The number of observations should be the same across all models")
    return np.stack(cols, axis=1)


def _stacking_weights(elpds, ic, scale):
    """Stacking of predictive distributions: maximise sum_i log(sum_k w_k exp(elpd_ik)) on the simplex
    (compare.py:477-536; SLSQP, ftol 1e-12, maxiter 2000, analytic gradient)."""
    names = list(elpds)
    K = len(names)
    pw = _to_log_scale(_pointwise_matrix(elpds, f"{ic}_i"), scale)
    ex = np.exp(pw - pw.max(axis=1, keepdims=True))

    def full(w):
        w = np.concatenate((w, [max(1.0 - np.sum(w), 0.0)]))
        w = np.maximum(w, 0)
        return w / np.sum(w)

    def objective(w):
        return -np.sum(np.log(np.dot(ex, full(w))))

    def gradient(w):
        denom = np.dot(ex, full(w))
        grad = np.zeros(K - 1)
        for k in range(K - 1):
            grad[k] = np.sum((ex[:, k] - ex[:, -1]) / denom)
        return -grad

    sol = optimize.minimize(
        objective, np.full(K - 1, 1.0 / K), jac=gradient, bounds=[(0.0, 1.0)] * (K - 1),
        constraints=[{"type": "ineq", "fun": lambda x: 1.0 - np.sum(x)}, {"type": "ineq", "fun": np.sum}],
        method="SLSQP", options={"ftol": 1e-12, "maxiter": 2000})
    return dict(zip(names, full(sol.x)))


def _bb_pseudo_bma_weights(elpds, ic, b_samples, alpha, seed, scale):
    """Bayesian-bootstrap pseudo-BMA (compare.py:539-577)."""
    if seed is not None:
        np.random.seed(seed)
    names = list(elpds)
    pw = _pointwise_matrix(elpds, f"{ic}_i")
    rows = pw.shape[0]
    pw = _to_log_scale(pw * rows, scale)
    rng = np.random.RandomState(seed) if isinstance(seed, int) else seed
    dirichlet = st.dirichlet.rvs(alpha=[alpha] * rows, size=b_samples, random_state=rng)
    z = dirichlet @ pw
    w = np.exp(z - z.max(axis=1, keepdims=True))
    w /= w.sum(axis=1, keepdims=True)
    return dict(zip(names, w.mean(axis=0))), pd.Series(z.std(axis=0), index=names)


def _pseudo_bma_weights(elpds, ic, scale):
    """Softmax of the totals (compare.py:580-596)."""
    names = list(elpds)
    tot = _to_log_scale([elpds[n][f"elpd_{ic}"] for n in names], scale)
    w = np.exp(tot - np.max(tot))
    return dict(zip(names, w / np.sum(w)))


def _calculate_ics(compare_dict, scale, ic, var_name):
    """Validate precomputed ELPDData and compute the missing ones (compare.py:338-456)."""
    pre = {n: e for n, e in compare_dict.items() if isinstance(e, ELPDData)}
    pre_ic = pre_scale = None
    if pre:
        first = next(reversed(pre.values()))
        pre_ic = first.index[0].split("_")[1]
        pre_scale = first["scale"]
        if any(e.index[0].split("_")[1] != pre_ic for e in pre.values()):
            raise ValueError("All information criteria to be compared must be the same")
        if any(e["scale"] != pre_scale for e in pre.values()):
            raise ValueError("All information criteria to be compared must use the same scale")
        if any(f"{pre_ic}_i" not in e for e in pre.values()):
            raise ValueError("Not all provided ELPDData have been calculated with pointwise=True")
        if ic is not None and ic.lower() != pre_ic.lower():
            warnings.warn("Provided ic argument is incompatible with precomputed elpd data. "
                          f"Using ic from precomputed elpddata: {pre_ic}", stacklevel=3)
            ic = pre_ic
        if scale is not None and scale.lower() != pre_scale:
            warnings.warn("Provided scale argument is incompatible with precomputed elpd data. "
                          f"Using scale from precomputed elpddata: {pre_scale}", stacklevel=3)
            scale = pre_scale
    ic = (ic.lower() if ic is not None else (pre_ic or "loo"))
    scale = (scale.lower() if scale is not None else (pre_scale or "log"))
    if scale not in _SCALES:
        raise ValueError(f"Scale must be one of {set(_SCALES)}, not {scale}")
    ic_func = waic if ic == "waic" else loo
    out = {}
    for name, dataset in compare_dict.items():
        if isinstance(dataset, ELPDData):
            out[name] = dataset.copy()
            continue
        try:
            out[name] = ic_func(deepcopy(dataset), pointwise=True, var_name=var_name, scale=scale)
        except Exception as err:  # compare.py:449-452
            raise err.__class__(f"Encountered error trying to compute {ic} from model {name}.") from err
    return out, scale, ic


def loo_compare(compare_dict, ic="loo", method="stacking", b_samples=1000, alpha=1, seed=None, scale=None,
                var_name=None, observations=None, estimator=None, K=None, folds=None, stratify=None,
                random_seed=None):
    """Compare models by ELPD (PSIS-LOO or WAIC) -- same signature and DataFrame as ``pyloo.loo_compare``."""
    if not isinstance(compare_dict, dict):
        raise TypeError("compare_dict must be a dictionary")
    if len(compare_dict) < 2:
        raise ValueError("You must specify at least two models for comparison")
    if scale is None:
        scale = "log"
    scale = scale.lower()
    if scale not in _SCALES:
        raise ValueError("Scale must be 'log', 'negative_log' or 'deviance'")
    method = method.lower()
    if method not in _METHODS:
        raise ValueError("Method must be 'stacking', 'BB-pseudo-BMA' or 'pseudo-BMA'")
    if ic not in ("loo", "waic", "kfold"):
        raise ValueError("ic must be 'loo', 'waic', or 'kfold'")
    if ic == "kfold" or observations is not None:
        raise NotImplementedError("ic='kfold' and subsampled comparison need model refits / subsampling "
                                  "and are outside the B200 hot path")

    elpds, scale, ic = _calculate_ics(compare_dict, scale, ic, var_name)

    names = list(elpds)
    totals = np.array([elpds[n][f"elpd_{ic}"] for n in names])
    order = np.argsort(totals) if scale != "log" else np.argsort(-totals)  # compare.py:202-207
    ranked = [names[i] for i in order]
    best = ranked[0]
    sign = {"log": 1, "negative_log": -1, "deviance": -2}[scale]
    diffs, ses, dses = [], [], []
    for name in ranked:
        if name == best:
            diffs.append(0)
            dses.append(0)
        else:
            diffs.append((elpds[name][f"elpd_{ic}"] - elpds[best][f"elpd_{ic}"]) * sign)  # :219-223
            delta = (np.asarray(elpds[name][f"{ic}_i"].values, dtype=float).ravel()
                     - np.asarray(elpds[best][f"{ic}_i"].values, dtype=float).ravel())
            dses.append(np.sqrt(len(delta) * np.var(delta)))  # :226-227
        ses.append(elpds[name]["se"])

    if method == "stacking":
        weights = _stacking_weights(elpds, ic, scale)
    elif method == "bb-pseudo-bma":
        weights, boot_se = _bb_pseudo_bma_weights(elpds, ic, b_samples, alpha, seed, scale)
        ses = [boot_se[n] for n in ranked]  # compare.py:243-245
    else:
        weights = _pseudo_bma_weights(elpds, ic, scale)

    return pd.DataFrame(
        {
            "rank": range(len(ranked)),
            f"elpd_{ic}": [elpds[n][f"elpd_{ic}"] for n in ranked],
            f"p_{ic}": [elpds[n][f"p_{ic}"] for n in ranked],
            "elpd_diff": diffs,
            "weight": [weights[n] for n in ranked],
            "se": ses,
            "dse": dses,
            "warning": [elpds[n]["warning"] for n in ranked],
            "scale": scale,
        },
        index=ranked,
    )

"""``loo_score`` -- leave-one-out CRPS / scaled CRPS.

Drop-in for ``pyloo.loo_score`` (reference: pyloo/loo_score.py:48-532).  Each of the ``permutations + 1``
expectation passes of the reference -- ``psislw`` of the (joint) log ratios followed by ``e_loo`` of
``|X - X'|`` or ``|X - y|`` (:227-237, :304-321) -- is one device-resident PSIS -> weighted-mean chain here
(``engine.psis_expectation_host``).  The shuffle uses ``np.random.permutation`` exactly like the reference
(:306), so seeding NumPy's global generator reproduces its permutations.
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import engine
from .data import SAMPLE_DIM, get_log_likelihood, to_inference_data, wrap_like
from .ess import relative_efficiency
from .rcparams import rcParams

__all__ = ["loo_score", "LooScoreResult"]


@dataclass
class LooScoreResult:
    """``estimates`` (structured array with fields ``Estimate`` / ``SE``), ``pointwise`` scores and, with
    ``pointwise=True``, ``pareto_k`` / ``good_k`` / ``warning`` (pyloo/loo_score.py:19-45)."""

    estimates: np.ndarray
    pointwise: np.ndarray
    pareto_k: Any = None
    good_k: float | None = None
    warning: bool | None = None


def _pick(idata, group, var, what):
    if not hasattr(idata, group):
        raise ValueError(f"InferenceData object does not have a {group} group")
    ds = getattr(idata, group)
    if var is None:
        names = list(ds.data_vars)
        if len(names) == 1:
            var = names[0]
        else:
            raise ValueError(f"Multiple variables found in {group} group. Please specify {what} from: {names}")
    elif var not in ds.data_vars:
        raise ValueError(f"Variable '{var}' not found in {group} group. Available variables: "
                         f"{list(ds.data_vars)}")
    return ds[var], var


def _stacked(da):
    if "chain" in da.dims and "draw" in da.dims:
        da = da.stack(__sample__=("chain", "draw"))
    return da


def _crps(exx, exy, scale=False):
    """CRPS ``0.5 E|X-X'| - E|X-y|`` or its scaled variant (pyloo/loo_score.py:326-346)."""
    if scale:
        return -exy / exx - 0.5 * np.log(exx)
    return 0.5 * exx - exy


def loo_score(data, x_group="posterior_predictive", x_var=None, x2_group=None, x2_var=None,
              y_group="observed_data", y_var=None, var_name=None, pointwise=None, permutations=1, reff=None,
              scale=False, **kwargs):
    """LOO-CRPS (``scale=False``) or LOO-SCRPS; same parameters, errors, warnings and ``LooScoreResult`` as
    ``pyloo.loo_score``."""
    idata = to_inference_data(data)
    log_lik = get_log_likelihood(idata, var_name=var_name)
    pointwise = rcParams["stats.ic_pointwise"] if pointwise is None else pointwise

    x_da, x_var = _pick(idata, x_group, x_var, "x_var")                      # loo_score.py:452-472
    x2_group = x_group if x2_group is None else x2_group
    if not hasattr(idata, x2_group):
        raise ValueError(f"InferenceData object does not have a {x2_group} group")
    x2_var = x_var if x2_var is None else x2_var
    if x2_var not in getattr(idata, x2_group).data_vars:
        raise ValueError(f"Variable '{x2_var}' not found in {x2_group} group. Available variables: "
                         f"{list(getattr(idata, x2_group).data_vars)}")
    x2_da = getattr(idata, x2_group)[x2_var]
    y_da, _ = _pick(idata, y_group, y_var, "y_var")
    x_da, x2_da, log_lik = _stacked(x_da), _stacked(x2_da), _stacked(log_lik)

    # ---- loo_score.py:349-414
    if tuple(x_da.dims) != tuple(x2_da.dims):
        raise ValueError("x and x2 must have the same dimensions")
    if x_da.shape != x2_da.shape:
        raise ValueError("x and x2 must have the same shape")
    xv, x2v, yv = (np.asarray(a.values, dtype=np.float64) for a in (x_da, x2_da, y_da))
    if np.isnan(xv).any() or np.isnan(x2v).any() or np.isnan(yv).any():
        warnings.warn("NaN values detected in input data. These may lead to unreliable results.", UserWarning,
                      stacklevel=2)
    if np.isinf(xv).any() or np.isinf(x2v).any() or np.isinf(yv).any():
        warnings.warn("Infinite values detected in input data. These may lead to unreliable results.",
                      UserWarning, stacklevel=2)
    obs_dims = [d for d in x_da.dims if d != SAMPLE_DIM]
    if set(obs_dims) != set(y_da.dims):
        raise ValueError(f"y dimensions {list(y_da.dims)} are not compatible with x dimensions {x_da.dims}")
    if SAMPLE_DIM not in log_lik.dims:
        raise ValueError("log_lik must have '__sample__' dimension")
    ll_obs = [d for d in log_lik.dims if d != SAMPLE_DIM]
    if set(ll_obs) != set(obs_dims):
        raise ValueError(f"log_lik dimensions {log_lik.dims} are not compatible with x dimensions {x_da.dims}")

    # sample axis last, observation dims in x's order
    def rows(da, dims_wanted):
        vals = np.asarray(da.values, dtype=np.float64)
        perm = [da.dims.index(d) for d in dims_wanted]
        return vals.transpose(perm) if perm != list(range(vals.ndim)) else vals

    xv = rows(x_da, obs_dims + [SAMPLE_DIM])
    x2v = rows(x2_da, obs_dims + [SAMPLE_DIM])
    llv = rows(log_lik, obs_dims + [SAMPLE_DIM])
    yv = rows(y_da, obs_dims)
    obs_shape, S = xv.shape[:-1], xv.shape[-1]
    x2, x22, ll2 = (a.reshape(-1, S) for a in (xv, x2v, llv))
    y2 = yv.reshape(-1, 1)

    if reff is None:  # loo_score.py:202-217
        if not hasattr(idata, "posterior"):
            raise TypeError("Must be able to extract a posterior group from data.")
        posterior = idata.posterior
        reff = 1.0 if len(posterior.chain) == 1 else relative_efficiency(posterior, S)

    kind = kwargs.get("type", "mean")
    exx = np.zeros(x2.shape[0])
    for _ in range(permutations):  # loo_score.py:219-225, :304-323
        shuffle = np.random.permutation(S)
        joint = -ll2 - ll2[:, shuffle]
        val, _, _ = engine.psis_expectation_host(np.abs(x2 - x22[:, shuffle]), joint, reff, kind)
        exx = exx + val
    exx = exx / permutations
    exy, _, pareto_k = engine.psis_expectation_host(np.abs(x2 - y2), -ll2, reff, kind)  # :227-237

    score_pw = _crps(exx, exy, scale=scale).reshape(obs_shape)
    estimates = np.array([float(score_pw.mean()), float(score_pw.std() / np.sqrt(score_pw.size))])
    estimates.dtype = np.dtype([("Estimate", float), ("SE", float)])  # loo_score.py:244-246
    result = LooScoreResult(estimates=estimates, pointwise=score_pw)
    if pointwise:  # loo_score.py:253-272
        good_k = min(1 - 1 / np.log10(S), 0.7)
        result.pareto_k = wrap_like(log_lik, pareto_k.reshape(obs_shape), tuple(obs_dims), "pareto_shape")
        result.good_k = good_k
        if np.any(pareto_k > good_k):
            warnings.warn(
                f"Estimated shape parameter of Pareto distribution is greater than {good_k:.2f} for "
                f"{np.sum(pareto_k > good_k)} observations. This indicates that importance sampling may be "
                "unreliable because the marginal posterior and LOO posterior are very different.", UserWarning,
                stacklevel=2)
            result.warning = True
        else:
            result.warning = False
    return result

"""CPU oracle for the callers either side of ``psislw`` -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the reference (jordandeklerk/pyloo) for

* the SIS / TIS branches of ``compute_importance_weights`` (pyloo/sis.py:86-106, pyloo/tis.py:91-120,
  dispatch pyloo/base.py:146-166) and of ``loo`` (pyloo/loo.py:286-289, :305-337);
* ``e_loo`` (pyloo/e_loo.py): weighted mean / variance / sd / quantiles (:429-554) and the function
  specific Pareto k (:266-390) with its derived diagnostics (:393-426).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

Parity status: PINNED.  ``oracle/gen_golden_is.py`` runs the reference's own ``_sislw`` / ``_tislw`` /
``k_hat`` / ``_wvar_func`` / ``_weighted_quantile`` (loaded from ``/root/reference`` with the stub modules of
``oracle/_refload.py``) on seeded inputs and stores inputs + outputs in ``tests/golden/is_eloo.npz``;
``tests/test_oracle_is_golden.py`` checks this restatement against them.

Note on ``k_hat``: the reference hands ``_gpdfit`` each tail in *descending* order with the cutoff
subtracted, so the last element is exactly 0, ``1 / ary[-1]`` is inf, the likelihood profile is all-NaN, every
grid weight is dropped and the estimate collapses to the prior mean ``5 / (n + 10)`` (1/6 for the default
``tail_len = 20``).  This is the reference's observable behaviour; the restatement (and the CUDA kernel)
evaluates the same arithmetic rather than hard-coding the constant.
"""

from __future__ import annotations

import numpy as np

from .psis_oracle import gpdfit, logsumexp_row
from .psis_oracle import loo_pointwise as _loo_pointwise_oracle
from .psis_oracle import psislw as _psislw_oracle

__all__ = ["sislw_row", "tislw_row", "islw", "loo_is_pointwise", "loo_is_summary", "k_hat",
           "weighted_mean", "weighted_variance", "weighted_quantile", "e_loo_arrays", "pareto_min_ss",
           "pareto_khat_threshold", "pareto_convergence_rate", "predictive_metric",
           "loo_predictive_metric_arrays", "crps", "loo_score_arrays", "loo_group_arrays"]


# ------------------------------------------------------------------------------- SIS / TIS
def sislw_row(row):
    """pyloo/sis.py:101-106."""
    x = np.array(row, dtype=np.float64)
    x -= np.max(x)
    x -= logsumexp_row(x)
    w = np.exp(x)
    return x, float(1 / np.sum(w**2))


def tislw_row(row, n_samples: int):
    """pyloo/tis.py:108-120."""
    x = np.array(row, dtype=np.float64)
    x -= np.max(x)
    log_z = logsumexp_row(x) - np.log(n_samples)       # tis.py:112
    log_cut = log_z + 0.5 * np.log(n_samples)          # tis.py:114
    x = np.minimum(x, log_cut)                         # tis.py:115
    x -= logsumexp_row(x)                              # tis.py:116
    w = np.exp(x)
    return x, float(1 / np.sum(w**2))


def islw(lw_ns, method: str):
    """Batch driver (pyloo/base.py:146-166): rows are observations, last axis is the sample axis."""
    a = np.asarray(lw_ns, dtype=np.float64)
    flat = a.reshape(-1, a.shape[-1])
    out = np.empty_like(flat)
    ess = np.empty(flat.shape[0])
    with np.errstate(all="ignore"):
        for i, row in enumerate(flat):
            out[i], ess[i] = sislw_row(row) if method == "sis" else tislw_row(row, a.shape[-1])
    return out.reshape(a.shape), ess.reshape(a.shape[:-1])


def loo_is_pointwise(ll_sn, method: str):
    """pyloo/loo.py:218-227 (NaN -> -1e10), :286-289 (weights of -ll; lw += ll), :319-324, :329-337."""
    ll = np.asarray(ll_sn, dtype=np.float64).T.copy()
    n_nan = int(np.isnan(ll).sum())
    ll = np.where(np.isnan(ll), -1e10, ll)
    S = ll.shape[-1]
    lw, ess = islw(-ll, method)
    lw += ll
    with np.errstate(all="ignore"):
        elpd_i = np.array([logsumexp_row(r) for r in lw])
        lppd_i = np.array([logsumexp_row(r, b_inv=S) for r in ll])
    return {"elpd_i": elpd_i, "ess_i": ess, "lppd_i": lppd_i, "n_nan_in": n_nan, "n_samples": S}


def loo_is_summary(ll_sn, method: str, scale_value: float = 1.0):
    """pyloo/loo.py:305-317 (ESS warning), :326-342 (totals)."""
    pw = loo_is_pointwise(ll_sn, method)
    loo_i = scale_value * pw["elpd_i"]
    n = loo_i.size
    elpd = loo_i.sum()
    se = (n * np.var(loo_i)) ** 0.5
    lppd = np.sum(pw["lppd_i"])
    return {"elpd_loo": elpd, "se": se, "p_loo": lppd - elpd / scale_value,
            "p_loo_se": np.sqrt(np.sum(np.var(loo_i))), "looic": -2 * elpd, "looic_se": 2 * se,
            "warning": bool(np.min(pw["ess_i"]) < pw["n_samples"] * 0.1), **pw}


# ------------------------------------------------------------------------------- e_loo
def k_hat(x_vals, log_ratios, tail_len: int = 20) -> float:
    """pyloo/e_loo.py:350-390."""
    lr = np.asarray(log_ratios, dtype=np.float64)
    with np.errstate(all="ignore"):
        r = np.exp(lr - np.max(lr))                          # :350
        top_r = -np.sort(-r)[:tail_len]                      # :351
        if len(top_r) < 5 or np.allclose(top_r, top_r[0]):   # :353
            k_r = np.inf
        else:
            k_r, _ = gpdfit(top_r - top_r[-1])               # :356-357
        if x_vals is None:
            return float(k_r)
        h = np.asarray(x_vals, dtype=np.float64)
        if (np.allclose(h, h[0]) or len(np.unique(h)) == 2 or np.any(np.isnan(h))
                or np.any(np.isinf(h))):                     # :359-366
            return float(k_r)
        hr = h * r                                           # :368
        left = np.sort(hr)[:tail_len]                        # :370
        right = -np.sort(-hr)[:tail_len]                     # :371
        if len(left) < 5 or np.allclose(left, left[0]):      # :373
            k_left = -np.inf
        else:
            k_left, _ = gpdfit(-(left - left[-1]))           # :376-377
        if len(right) < 5 or np.allclose(right, right[0]):   # :379
            k_right = -np.inf
        else:
            k_right, _ = gpdfit(right - right[-1])           # :382-383
        k_hr = max(k_left, k_right)                          # :385 (Python max: NaN-order dependent)
        if np.isnan(k_hr) and np.isnan(k_r):                 # :387
            return float("nan")
        return float(max(k_hr, k_r))                         # :390


def _normalised_weights(lw_row):
    """pyloo/e_loo.py:557-559 followed by :434 / :446 / :473."""
    lw = np.asarray(lw_row, dtype=np.float64)
    return np.exp(lw - logsumexp_row(lw))


def weighted_mean(x_row, lw_row) -> float:
    """pyloo/e_loo.py:429-436."""
    return float(np.sum(_normalised_weights(lw_row) * np.asarray(x_row, dtype=np.float64)))


def weighted_variance(x_row, lw_row) -> float:
    """pyloo/e_loo.py:439-457 with ``_wvar_func`` :518-531."""
    x = np.asarray(x_row, dtype=np.float64)
    w = _normalised_weights(lw_row)
    if np.allclose(x, x[0]):
        return 0.0
    w2 = np.sum(w**2)
    if np.isclose(w2, 1.0):
        return 0.0
    mean = np.sum(w * x)
    mean_sq = np.sum(w * x**2)
    var = (mean_sq - mean**2) / (1 - w2)
    return float(max(var, 0.0))


def weighted_quantile(x_row, lw_row, prob: float) -> float:
    """pyloo/e_loo.py:466-515 with ``_weighted_quantile`` :534-554."""
    x = np.asarray(x_row, dtype=np.float64)
    w = _normalised_weights(lw_row)
    if np.allclose(w, w[0]):
        return float(np.quantile(x, prob))
    order = np.argsort(x)
    xs, ws = x[order], w[order]
    cdf = np.cumsum(ws) / np.sum(ws)
    hit = np.where(cdf >= prob)[0]
    if len(hit) == 0:
        return float(xs[-1])
    j = hit[0]
    if j == 0:
        return float(xs[0])
    return float(xs[j - 1] + (xs[j] - xs[j - 1]) * (prob - cdf[j - 1]) / (cdf[j] - cdf[j - 1]))


def pareto_min_ss(k: float) -> float:
    """pyloo/e_loo.py:393-398."""
    return 10 ** (1 / (1 - max(0, k))) if k < 1 else float("inf")


def pareto_khat_threshold(n_samples: int) -> float:
    """pyloo/e_loo.py:401-403."""
    return 1 - 1 / np.log10(n_samples)


def pareto_convergence_rate(k: float, n_samples: int) -> float:
    """pyloo/e_loo.py:406-426."""
    if k < 0:
        return 1.0
    if k > 1:
        return 0.0
    if k == 0.5:
        return 1 - 1 / np.log(n_samples)
    if 0 < k < 1:
        n = n_samples
        return max(0, (2 * (k - 1) * n ** (2 * k + 1) + (1 - 2 * k) * n ** (2 * k) + n**2)
                   / ((n - 1) * (n - n ** (2 * k))))
    return 1.0


def e_loo_arrays(x_ns, lw_ns, lr_ns=None, kind: str = "mean", probs=None, tail_len: int = 20):
    """Batch form of pyloo/e_loo.py:220-263 on ``(N, S)`` arrays: ``value`` ((N,) or (N, len(probs))),
    ``pareto_k``, ``min_ss``, ``khat_threshold``, ``convergence_rate``."""
    x = np.asarray(x_ns, dtype=np.float64)
    lw = np.asarray(lw_ns, dtype=np.float64)
    lr = lw if lr_ns is None else np.asarray(lr_ns, dtype=np.float64)
    N, S = lw.shape
    with np.errstate(all="ignore"):
        if kind == "mean":
            value = np.array([weighted_mean(x[i], lw[i]) for i in range(N)])
            h = x
        elif kind in ("variance", "sd"):
            value = np.array([weighted_variance(x[i], lw[i]) for i in range(N)])
            if kind == "sd":
                value = np.sqrt(value)
            h = x**2                                           # e_loo.py:234-236
        else:
            value = np.array([[weighted_quantile(x[i], lw[i], p) for p in np.atleast_1d(probs)]
                              for i in range(N)])
            h = None
        k = np.array([k_hat(None if h is None else h[i], lr[i], tail_len) for i in range(N)])
    return {"value": value, "pareto_k": k, "min_ss": np.array([pareto_min_ss(v) for v in k]),
            "khat_threshold": np.full(N, pareto_khat_threshold(S)),
            "convergence_rate": np.array([pareto_convergence_rate(v, S) for v in k])}


# ------------------------------------------------------------------------------- consumers of e_loo
def predictive_metric(y, yhat, metric: str) -> dict:
    """pyloo/loo_predictive_metric.py:234-356 (``_mae``, ``_mse``, ``_rmse``, ``_accuracy``,
    ``_balanced_accuracy``)."""
    y = np.asarray(y)
    yhat = np.asarray(yhat)
    n = len(y)
    if metric in ("mae", "mse", "rmse"):
        e = np.abs(y - yhat) if metric == "mae" else (y - yhat) ** 2
        est, se = np.mean(e), np.std(e, ddof=1) / np.sqrt(n)
        if metric == "rmse":                                  # :291-298
            return {"estimate": np.sqrt(est), "se": np.sqrt(se**2 / est / 4)}
        return {"estimate": est, "se": se}
    pred = (yhat > 0.5).astype(int)
    if metric == "acc":                                       # :319-326
        est = np.mean((pred == y).astype(int))
        return {"estimate": est, "se": np.sqrt(est * (1 - est) / n)}
    mask = y == 0                                             # :347-356
    tn = np.mean(pred[mask] == y[mask])
    tp = np.mean(pred[~mask] == y[~mask])
    return {"estimate": (tp + tn) / 2, "se": np.sqrt((tp * (1 - tp) + tn * (1 - tn)) / 4 / n)}


def loo_predictive_metric_arrays(x_ns, ll_ns, y, metric: str, reff: float = 1.0) -> dict:
    """pyloo/loo_predictive_metric.py:208-231 on ``(N, S)`` arrays: PSIS weights of ``-ll``, weighted mean of
    the predictive draws, metric against ``y``."""
    ll = np.asarray(ll_ns, dtype=np.float64)
    lw, _ = _psislw_oracle(-ll, reff)
    pred = e_loo_arrays(x_ns, lw, -ll, "mean")["value"]
    return predictive_metric(np.asarray(y).flatten(), pred, metric)


def crps(exx, exy, scale: bool = False):
    """pyloo/loo_score.py:343-346."""
    if scale:
        return -exy / exx - 0.5 * np.log(exx)
    return 0.5 * exx - exy


def loo_score_arrays(x_ns, x2_ns, ll_ns, y, permutations: int = 1, reff: float = 1.0, scale: bool = False):
    """pyloo/loo_score.py:219-246 and :304-323 on ``(N, S)`` arrays (``y`` of length N).  Draws the shuffles
    from NumPy's global generator like the reference (:306)."""
    x = np.asarray(x_ns, dtype=np.float64)
    x2 = np.asarray(x2_ns, dtype=np.float64)
    ll = np.asarray(ll_ns, dtype=np.float64)
    S = x.shape[-1]
    exx = 0
    for _ in range(permutations):
        shuffle = np.random.permutation(S)                     # :306
        joint = -ll - ll[:, shuffle]                           # :308-311
        lw, _ = _psislw_oracle(joint, reff)                    # :312
        exx = exx + e_loo_arrays(np.abs(x - x2[:, shuffle]), lw, joint, "mean")["value"]  # :314-321
    exx = exx / permutations                                   # :225
    lw, k = _psislw_oracle(-ll, reff)                          # :227
    exy = e_loo_arrays(np.abs(x - np.asarray(y, dtype=np.float64).reshape(-1, 1)), lw, -ll, "mean")["value"]
    pw = crps(exx, exy, scale)
    return {"pointwise": pw, "estimate": float(pw.mean()), "se": float(pw.std() / np.sqrt(pw.size)),
            "pareto_k": k}


def loo_group_arrays(ll_sn, group_ids, method: str = "psis", reff: float = 1.0, scale_value: float = 1.0):
    """pyloo/loo_group.py:162-163 (sorted unique groups), :188-197 (NaN -> -1e10), :215-222 (per-group sums of
    the stacked log-likelihood), :226-233 and :281-305 (the LOO pass per group), :283-303 (totals).
    The per-group pass is the pinned single-observation oracle applied to the group sums."""
    ll = np.asarray(ll_sn, dtype=np.float64).T           # (N, S) like .stack(__sample__=...)
    ll = np.where(np.isnan(ll), -1e10, ll)
    gid = np.asarray(group_ids)
    groups = np.unique(gid)
    sums = np.array([ll[gid == g].sum(axis=0) for g in groups])   # :216-222
    if method == "psis":
        pw = _loo_pointwise_oracle(sums.T, reff)
        diag = pw["pareto_k"]
    else:
        pw = loo_is_pointwise(sums.T, method)
        diag = pw["ess_i"]
    logo_i = scale_value * pw["elpd_i"]
    n = len(groups)
    elpd = logo_i.sum()
    se = (n * np.var(logo_i)) ** 0.5
    lppd = pw["lppd_i"].sum()
    return {"groups": groups, "group_sums": sums, "logo_i": logo_i, "diagnostic": diag, "elpd_logo": elpd, "se": se,
            "p_logo": lppd - elpd / scale_value, "p_logo_se": np.sqrt(np.sum(np.var(logo_i))),
            "logoic": -2 * elpd, "logoic_se": 2 * se}

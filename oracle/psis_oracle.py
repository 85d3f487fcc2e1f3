"""CPU oracle for the PSIS-LOO hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy restatement of the reference algorithm (jordandeklerk/pyloo) for the path
``psislw`` / ``loo(method="psis")`` / ``waic``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module, and only
as the checker or the timed CPU baseline.  The product (``pyloo_b200``) never imports it.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the *real* reference functions from
``/root/reference`` (stub ``xarray``/``arviz`` modules), runs them through the
reference's own ``make_ufunc`` loop and stores inputs + outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors bit-for-bit
(same NumPy build) / to 1e-13 (other builds), and against the known answers in SURVEY.md App. B.

Every function cites the reference lines it follows.  All arithmetic is float64 and uses the
same NumPy primitives in the same order as the reference so that, on the same NumPy build,
results are bit-identical.
"""

from __future__ import annotations

import math

import numpy as np

__all__ = [
    "tail_length",
    "logsumexp_row",
    "gpdfit",
    "gpinv",
    "psislw_row",
    "psislw",
    "loo_pointwise",
    "loo_summary",
    "waic_pointwise",
    "waic_summary",
]

CUTOFFMIN = float(np.log(np.finfo(float).tiny))  # pyloo/psis.py:90


def tail_length(n_samples: int, reff: float) -> int:
    """M = ceil(min(S/5, 3*sqrt(S/reff))) -- pyloo/psis.py:89, pyloo/base.py:139-141.

    The reference stores ``cutoff_ind = -M - 1``.
    """
    return int(np.ceil(min(n_samples / 5.0, 3 * (n_samples / reff) ** 0.5)))


def logsumexp_row(row: np.ndarray, b_inv: float | None = None) -> float:
    """1-D ``_logsumexp`` -- pyloo/utils.py:344-359 (axis=None, copy=True path)."""
    row = np.asarray(row, dtype=np.float64)
    top = row.max()
    work = row - top  # utils.py:347-349 (copy then in-place subtract)
    np.exp(work, out=work)  # utils.py:350
    total = np.log(work.sum())  # utils.py:351-352 (pairwise np.sum)
    if b_inv is not None:
        top = top - np.log(b_inv)  # utils.py:353-354
    return float(total + top)  # utils.py:357


def gpdfit(sorted_tail: np.ndarray) -> tuple[float, float]:
    """Zhang-Stephens empirical-Bayes GPD fit -- pyloo/psis.py:181-208."""
    t = np.asarray(sorted_tail, dtype=np.float64)
    n = len(t)
    m_est = 30 + int(n**0.5)  # psis.py:184

    grid = 1 - np.sqrt(m_est / (np.arange(1, m_est + 1, dtype=float) - 0.5))  # :186
    grid /= 3 * t[int(n / 4 + 0.5) - 1]  # :187 (prior_bs = 3)
    grid += 1 / t[-1]  # :188

    k_grid = np.log1p(-grid[:, None] * t).mean(axis=1)  # :190
    prof = n * (np.log(-(grid / k_grid)) - k_grid - 1)  # :191
    w = 1 / np.exp(prof - prof[:, None]).sum(axis=1)  # :192

    keep = w >= 10 * np.finfo(float).eps  # :194
    if not np.all(keep):
        w = w[keep]
        grid = grid[keep]
    w /= w.sum()  # :198

    b_post = np.sum(grid * w)  # :201
    k_post = np.log1p(-b_post * t).mean()  # :203
    sigma = -k_post / b_post  # :205
    k_post = (n * k_post + 10 * 0.5) / (n + 10)  # :206 (prior_k = 10)
    return float(k_post), float(sigma)


def gpinv(probs: np.ndarray, kappa: float, sigma: float) -> np.ndarray:
    """Inverse GPD cdf -- pyloo/psis.py:211-231."""
    probs = np.asarray(probs, dtype=np.float64)
    q = np.full_like(probs, np.nan)
    if sigma <= 0:  # :214
        return q
    eps = np.finfo(float).eps
    inside = (probs > 0) & (probs < 1)  # :216
    if np.all(inside):
        if np.abs(kappa) < eps:
            q = -np.log1p(-probs)
        else:
            q = np.expm1(-kappa * np.log1p(-probs)) / kappa
        q *= sigma
    else:
        if np.abs(kappa) < eps:
            q[inside] = -np.log1p(-probs[inside])
        else:
            q[inside] = np.expm1(-kappa * np.log1p(-probs[inside])) / kappa
        q *= sigma
        q[probs == 0] = 0
        q[probs == 1] = np.inf if kappa >= 0 else -sigma / kappa
    return q


def psislw_row(raw: np.ndarray, m_tail: int, cutoffmin: float = CUTOFFMIN,
               return_tail: bool = False):
    """One observation of PSIS -- pyloo/psis.py:133-160.

    ``raw`` is not modified (the reference mutates a deep copy, psis.py:78).
    ``m_tail`` is M, i.e. ``cutoff_ind = -m_tail - 1``.
    With ``return_tail`` the tail index set (positions with x > cutoff) is returned too.
    """
    x = np.array(raw, dtype=np.float64, copy=True)
    x -= np.max(x)  # :134
    order = np.argsort(x)  # :135
    xcut = max(x[order[-m_tail - 1]], cutoffmin)  # :136 (Python max: NaN first arg stays NaN)
    exp_cut = np.exp(xcut)  # :138
    (tail_pos,) = np.where(x > xcut)  # :139
    tail = x[tail_pos]
    n = len(tail)
    if n <= 4:  # :142
        k = np.inf
    else:
        tail_order = np.argsort(tail)  # :146
        t = np.exp(tail) - exp_cut  # :147
        k, sigma = gpdfit(t[tail_order])  # :148
        if np.isfinite(k):  # :150
            p = np.arange(0.5, n) / n  # :153
            smooth = gpinv(p, k, sigma)  # :154
            smooth = np.log(smooth + exp_cut)  # :155
            x[tail_pos[tail_order]] = smooth  # :156
            x[x > 0] = 0  # :157
    x -= logsumexp_row(x)  # :158
    if return_tail:
        return x, float(k), tail_pos
    return x, float(k)


def psislw(log_weights: np.ndarray, reff: float = 1.0):
    """Batch ``psislw`` on an ndarray with samples on the last axis.

    Follows pyloo/psis.py:78-106 with the per-observation loop of pyloo/utils.py:171-176.
    Returns ``(lw, k)`` with ``k.shape == log_weights.shape[:-1]`` (0-d for 1-D input).
    """
    lw = np.array(log_weights, dtype=np.float64, copy=True)
    n_samples = lw.shape[-1]
    m_tail = tail_length(n_samples, reff)
    k = np.empty(lw.shape[:-1])
    for idx in np.ndindex(lw.shape[:-1]):
        lw[idx], k[idx] = psislw_row(lw[idx], m_tail)
    return lw, k


def _as_obs_sample(ll_sn: np.ndarray) -> np.ndarray:
    """(S, N...) sample-major log-likelihood -> strided (N..., S) view, like
    ``DataArray.stack(__sample__=("chain","draw"))`` in pyloo/loo.py:189."""
    ll_sn = np.asarray(ll_sn, dtype=np.float64)
    return np.moveaxis(ll_sn, 0, -1)


def loo_pointwise(ll_sn: np.ndarray, reff: float):
    """Pointwise PSIS-LOO pieces for a sample-major ``(S, N)`` log-likelihood.

    Follows pyloo/loo.py:218-227 (NaN -> -1e10), :286-289 (weights, ``lw += ll``),
    :319-324 (``loo_i`` before scaling), :329-337 (``lppd_i``).
    Returns dict with ``elpd_i`` (scale = log), ``pareto_k``, ``lppd_i``, ``n_nan``.
    """
    ll = _as_obs_sample(ll_sn)
    nan_mask = np.isnan(ll)
    n_nan = int(nan_mask.sum())
    if n_nan:
        ll = np.where(nan_mask, -1e10, ll)  # loo.py:227
    n_samples = ll.shape[-1]
    m_tail = tail_length(n_samples, reff)
    obs_shape = ll.shape[:-1]
    elpd_i = np.empty(obs_shape)
    pareto_k = np.empty(obs_shape)
    lppd_i = np.empty(obs_shape)
    for idx in np.ndindex(obs_shape):
        row = ll[idx]
        lw, pareto_k[idx] = psislw_row(-row, m_tail)  # loo.py:286-288
        lw += row  # loo.py:289
        elpd_i[idx] = logsumexp_row(lw)  # loo.py:319-324
        lppd_i[idx] = logsumexp_row(row, b_inv=n_samples)  # loo.py:329-337
    return {"elpd_i": elpd_i, "pareto_k": pareto_k, "lppd_i": lppd_i, "n_nan": n_nan,
            "n_samples": n_samples}


_SCALE = {"log": 1, "negative_log": -1, "deviance": -2}  # loo.py:195-200


def loo_summary(ll_sn: np.ndarray, reff: float, scale: str = "log") -> dict:
    """All ELPDData rows of ``loo(pointwise=True)`` -- pyloo/loo.py:249, :291-293, :326-342."""
    pw = loo_pointwise(ll_sn, reff)
    sv = _SCALE[scale]
    n_samples = pw["n_samples"]
    loo_i = sv * pw["elpd_i"]
    n_points = int(np.prod(loo_i.shape))
    good_k = min(1 - 1 / np.log10(n_samples), 0.7)  # loo.py:249
    elpd = loo_i.sum()  # :326
    se = (n_points * np.var(loo_i)) ** 0.5  # :327
    lppd = np.sum(pw["lppd_i"])  # :329
    p_loo = lppd - elpd / sv  # :339
    p_loo_se = np.sqrt(np.sum(np.var(loo_i)))  # :340
    return {
        "elpd_loo": float(elpd), "se": float(se), "p_loo": float(p_loo),
        "p_loo_se": float(p_loo_se), "n_samples": n_samples, "n_data_points": n_points,
        "warning": bool(np.any(pw["pareto_k"] > good_k)),  # :292
        "n_high_k": int(np.sum(pw["pareto_k"] > good_k)),
        "loo_i": loo_i, "pareto_k": pw["pareto_k"], "scale": scale,
        "looic": float(-2 * elpd), "looic_se": float(2 * se),  # :341-342
        "good_k": float(good_k), "lppd_i": pw["lppd_i"], "n_nan": pw["n_nan"],
    }


def waic_pointwise(ll_sn: np.ndarray):
    """Pointwise WAIC pieces -- pyloo/waic.py:110-145."""
    ll = _as_obs_sample(ll_sn)
    n_nan = int(np.isnan(ll).sum())
    n_inf = int(np.isinf(ll).sum())
    if n_nan:
        ll = np.where(np.isnan(ll), -1e10, ll)  # waic.py:120
    if n_inf:
        ll = np.where(np.isinf(ll), np.where(ll > 0, 1e10, -1e10), ll)  # waic.py:129-132
    n_samples = ll.shape[-1]
    obs_shape = ll.shape[:-1]
    lppd_i = np.empty(obs_shape)
    for idx in np.ndindex(obs_shape):
        lppd_i[idx] = logsumexp_row(ll[idx], b_inv=n_samples)  # waic.py:137-143
    var_i = ll.var(axis=-1)  # waic.py:145 (ddof = 0)
    return {"lppd_i": lppd_i, "var_i": var_i, "n_nan": n_nan, "n_inf": n_inf,
            "n_samples": n_samples}


def waic_summary(ll_sn: np.ndarray, scale: str = "log") -> dict:
    """ELPDData rows of ``waic(pointwise=True)`` -- pyloo/waic.py:147-160."""
    pw = waic_pointwise(ll_sn)
    sv = _SCALE[scale]
    waic_i = sv * (pw["lppd_i"] - pw["var_i"])  # :157
    n_points = int(np.prod(waic_i.shape))
    return {
        "elpd_waic": float(np.sum(waic_i)),  # :159
        "se": float((n_points * np.var(waic_i)) ** 0.5),  # :158
        "p_waic": float(np.sum(pw["var_i"])),  # :160
        "n_samples": pw["n_samples"], "n_data_points": n_points,
        "warning": bool(np.any(pw["var_i"] > 0.4)),  # :147
        "waic_i": waic_i, "scale": scale, "var_i": pw["var_i"], "lppd_i": pw["lppd_i"],
        "n_nan": pw["n_nan"], "n_inf": pw["n_inf"],
    }


def _selfcheck() -> None:  # pragma: no cover - manual smoke
    rng = np.random.default_rng(0)
    lw, k = psislw(rng.normal(size=(3, 1000)), reff=0.9)
    assert np.allclose(np.exp(lw).sum(-1), 1.0)
    assert np.all(np.isfinite(k))
    assert math.isclose(CUTOFFMIN, -708.3964185322641)


if __name__ == "__main__":  # pragma: no cover
    _selfcheck()
    print("oracle selfcheck ok")

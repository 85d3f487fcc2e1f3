"""Relative efficiency ``r_eff`` for ``loo(reff=None)``.

The reference calls ``arviz.stats.diagnostics.ess(posterior, method="mean")`` (pyloo/loo.py:9,212-216):
third-party arithmetic that is *not* in the reference tree and not installed in this image, so this
NumPy restatement of the published algorithm (split chains, FFT autocovariance, Geyer's initial
positive + monotone sequences; Vehtari et al. 2021, as implemented by ArviZ / Stan) is "parity
unpinned" (SURVEY 8c).  When ArviZ is importable the real function is used instead.  Host-side,
O(#posterior values): not part of the accelerated path.
"""

from __future__ import annotations

import numpy as np

__all__ = ["ess_mean", "relative_efficiency"]


def _fast_len(n: int) -> int:
    try:
        from scipy.fft import next_fast_len

        return int(next_fast_len(n))
    except Exception:  # pragma: no cover
        return 1 << (n - 1).bit_length()


def _autocov(chains: np.ndarray) -> np.ndarray:
    n = chains.shape[1]
    m = _fast_len(2 * n)
    centred = chains - chains.mean(axis=1, keepdims=True)
    spec = np.fft.rfft(centred, n=m, axis=1)
    spec *= np.conjugate(spec)
    return np.fft.irfft(spec, n=m, axis=1)[:, :n] / n


def ess_mean(ary: np.ndarray) -> float:
    """Effective sample size of the mean for draws shaped ``(chain, draw)``."""
    ary = np.asarray(ary, dtype=float)
    if ary.ndim == 1:
        ary = ary[None, :]
    if not np.all(np.isfinite(ary)):
        return float("nan")
    half = ary.shape[1] // 2
    if half >= 2:  # split every chain in two halves
        ary = np.concatenate([ary[:, :half], ary[:, -half:]], axis=0)
    if (ary.max() - ary.min()) < np.finfo(float).resolution:
        return float(ary.size)
    n_chain, n_draw = ary.shape
    acov = _autocov(ary)
    mean_var = acov[:, 0].mean() * n_draw / (n_draw - 1.0)
    var_plus = mean_var * (n_draw - 1.0) / n_draw
    if n_chain > 1:
        var_plus += np.var(ary.mean(axis=1), ddof=1)
    rho = np.zeros(n_draw)
    even = 1.0
    rho[0] = even
    odd = 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = odd
    t = 1
    while t < (n_draw - 3) and (even + odd) > 0.0:  # Geyer: initial positive sequence
        even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if (even + odd) >= 0:
            rho[t + 1] = even
            rho[t + 2] = odd
        t += 2
    max_t = t - 2
    if even > 0:
        rho[max_t + 1] = even
    t = 1
    while t <= max_t - 2:  # Geyer: initial monotone sequence
        if (rho[t + 1] + rho[t + 2]) > (rho[t - 1] + rho[t]):
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    total = n_chain * n_draw
    tau = -1.0 + 2.0 * np.sum(rho[: max_t + 1]) + np.sum(rho[max_t + 1: max_t + 2])
    tau = max(tau, 1 / np.log10(total))
    return float(total / tau) if not np.isnan(rho).any() else float("nan")


def relative_efficiency(posterior, n_samples: int) -> float:
    """``mean_v ess_mean(posterior[v]) / n_samples`` over every posterior value (pyloo/loo.py:212-216)."""
    try:  # the real thing when available
        from arviz.stats.diagnostics import ess as az_ess  # type: ignore

        ess_p = az_ess(posterior, method="mean")
        return float(np.hstack([ess_p[v].values.flatten() for v in ess_p.data_vars]).mean() / n_samples)
    except ImportError:
        pass
    vals = []
    for name in posterior.data_vars:
        da = posterior[name]
        arr = np.asarray(da.values, dtype=float)
        dims = tuple(da.dims)
        arr = np.moveaxis(arr, (dims.index("chain"), dims.index("draw")), (0, 1))
        flat = arr.reshape(arr.shape[0], arr.shape[1], -1)
        vals.extend(ess_mean(flat[:, :, j]) for j in range(flat.shape[2]))
    return float(np.mean(vals) / n_samples)

"""``psislw`` -- Pareto smoothed importance sampling for a batch of observations.

Drop-in for ``pyloo.psislw`` (reference: pyloo/psis.py:25-111).  The per-observation Python loop
(pyloo/utils.py:171-176 driving ``_psislw``, pyloo/psis.py:114-160) is replaced by one call into the
CUDA library for the whole ``(N, S)`` batch.
"""

from __future__ import annotations

import numpy as np

from . import engine
from .data import SAMPLE_DIM, is_dataarray_like, wrap_like

__all__ = ["psislw"]


def _split_sample_axis(log_weights):
    """Return ``(values with the sample axis last, obs_dims or None)`` following pyloo/psis.py:79-88:
    a ``__sample__`` dimension if present, otherwise the last axis."""
    if is_dataarray_like(log_weights):
        dims = tuple(log_weights.dims)
        vals = np.asarray(log_weights.values)
        if SAMPLE_DIM in dims:
            ax = dims.index(SAMPLE_DIM)
            obs_dims = dims[:ax] + dims[ax + 1:]
            if ax != len(dims) - 1:
                vals = np.moveaxis(vals, ax, -1)
        else:
            obs_dims = dims[:-1]
        return vals, obs_dims
    return np.asarray(log_weights), None


def _batch_values(vals: np.ndarray, run):
    """ndarray ``(..., S)`` -> ``(lw (..., S), diagnostic (...))`` through ``run((N, S)) -> (lw, diag)``."""
    if vals.ndim < 1:
        raise ValueError("log_weights must have at least one dimension")
    obs_shape = vals.shape[:-1]
    S = vals.shape[-1]
    in_dtype = vals.dtype
    mat = vals if vals.ndim == 2 else vals.reshape(-1, S)
    lw, k = run(mat)
    lw = lw.reshape(*obs_shape, S)
    k = k.reshape(obs_shape)  # 0-d ndarray for 1-D input (pyloo/psis.py:92, test_psis.py:57)
    if in_dtype.kind == "f" and in_dtype != np.float64:
        lw = lw.astype(in_dtype)  # computed in FP64 regardless of the input dtype
    return lw, k


def _psislw_values(vals: np.ndarray, reff: float):
    """ndarray ``(..., S)`` -> ``(lw (..., S), k (...))`` on the GPU."""
    return _batch_values(vals, lambda mat: engine.psislw_host(mat, reff))


def _wrap_outputs(log_weights, obs_dims, lw, diag, diag_name):
    """ndarray in -> ndarrays out; DataArray in -> DataArrays named ``log_weights`` / ``diag_name``
    (pyloo/psis.py:107-110, pyloo/sis.py:79-83, pyloo/tis.py:84-88)."""
    if obs_dims is None:
        return lw, diag
    lw_da = wrap_like(log_weights, lw, (*obs_dims, SAMPLE_DIM if SAMPLE_DIM in log_weights.dims
                                         else log_weights.dims[-1]), "log_weights")
    return lw_da, wrap_like(log_weights, diag, obs_dims, diag_name)


def psislw(log_weights, reff: float = 1.0):
    """Pareto smoothed importance sampling (PSIS).

    Parameters
    ----------
    log_weights : DataArray-like or (..., S) array-like
        Log weights; a dimension named ``__sample__`` is the sample axis when present, otherwise
        the last axis (pyloo/psis.py:47-51).  Never modified (pyloo/psis.py:78).
    reff : float, default 1
        Relative MCMC efficiency ``ess / n``.

    Returns
    -------
    lw_out : same family and shape as the input (sample axis last)
        Smoothed, truncated and normalised log weights.
    kss : array of the observation shape (0-d ndarray for 1-D input)
        Pareto shape estimates; ``inf`` where the tail holds <= 4 draws (pyloo/psis.py:142-144).
    """
    vals, obs_dims = _split_sample_axis(log_weights)
    lw, k = _psislw_values(vals, reff)
    return _wrap_outputs(log_weights, obs_dims, lw, k, "pareto_shape")  # pyloo/psis.py:107-110

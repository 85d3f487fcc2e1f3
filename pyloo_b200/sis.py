"""``sislw`` -- standard importance sampling for a batch of observations.

Drop-in for ``pyloo.sislw`` (reference: pyloo/sis.py:11-106).  The per-observation loop over ``_sislw``
(pyloo/sis.py:72-78 -> :86-106) is one launch of the CUDA row kernel for the whole ``(N, S)`` batch.
"""

from __future__ import annotations

from . import engine
from .psis import _batch_values, _split_sample_axis, _wrap_outputs

__all__ = ["sislw"]


def sislw(log_weights):
    """Standard importance sampling (SIS): ``(normalised log weights, effective sample sizes)``.

    ``log_weights``: DataArray-like with a ``__sample__`` dimension, or an ``(..., S)`` array whose last
    axis is the sample axis (pyloo/sis.py:27-33).  Never modified (pyloo/sis.py:53).  DataArray outputs are
    named ``log_weights`` and ``ess`` (pyloo/sis.py:79-83)."""
    vals, obs_dims = _split_sample_axis(log_weights)
    lw, ess = _batch_values(vals, lambda mat: engine.islw_host(mat, "sis"))
    return _wrap_outputs(log_weights, obs_dims, lw, ess, "ess")

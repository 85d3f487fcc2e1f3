"""``tislw`` -- truncated importance sampling for a batch of observations.

Drop-in for ``pyloo.tislw`` (reference: pyloo/tis.py:11-120; Ionides 2008).  The per-observation loop over
``_tislw`` (pyloo/tis.py:77-83 -> :91-120) is one launch of the CUDA row kernel for the whole batch.
"""

from __future__ import annotations

from . import engine
from .psis import _batch_values, _split_sample_axis, _wrap_outputs

__all__ = ["tislw"]


def tislw(log_weights):
    """Truncated importance sampling (TIS): ``(truncated, normalised log weights, effective sample sizes)``.

    Weights are truncated at ``log_Z + 0.5 log S`` (pyloo/tis.py:111-115).  Same input conventions and output
    names as :func:`pyloo_b200.sislw` (pyloo/tis.py:57-67, :84-88)."""
    vals, obs_dims = _split_sample_axis(log_weights)
    lw, ess = _batch_values(vals, lambda mat: engine.islw_host(mat, "tis"))
    return _wrap_outputs(log_weights, obs_dims, lw, ess, "ess")

"""Importance-sampling dispatch -- the seam every LOO flavour of the reference calls
(pyloo/base.py:29-175; 16 call sites).  All three branches run on the GPU: PSIS
(base.py:138-144) through the split stream/tail kernels, SIS and TIS (base.py:146-152) through the
importance-sampling row kernel.
"""

from __future__ import annotations

from enum import Enum

from .data import SAMPLE_DIM, is_dataarray_like
from .psis import psislw
from .sis import sislw
from .tis import tislw

__all__ = ["ISMethod", "compute_importance_weights"]


class ISMethod(str, Enum):
    """Supported importance sampling methods (pyloo/base.py:18-23)."""

    PSIS = "psis"
    SIS = "sis"
    TIS = "tis"


def compute_importance_weights(log_weights=None, method=ISMethod.PSIS, reff: float = 1.0):
    """Unified importance-weight computation on the GPU.

    Same contract as ``pyloo.compute_importance_weights``: DataArray inputs need a ``__sample__``
    dimension or ``chain`` + ``draw`` (stacked automatically, base.py:93-98); ``ValueError`` for an
    unknown method (base.py:100-107) or missing weights (base.py:109-110); the diagnostic of the
    PSIS branch is named ``pareto_shape``, that of SIS / TIS ``ess`` (base.py:168-173)."""
    if is_dataarray_like(log_weights) and SAMPLE_DIM not in log_weights.dims:
        if "chain" in log_weights.dims and "draw" in log_weights.dims:
            log_weights = log_weights.stack(__sample__=("chain", "draw"))
        else:
            raise ValueError("log_weights must have a __sample__ dimension")
    if isinstance(method, str) and not isinstance(method, ISMethod):
        try:
            method = ISMethod(method.lower())
        except ValueError:
            valid = ", ".join(m.value for m in ISMethod)
            raise ValueError(f"Invalid method '{method}'. Must be one of: {valid}")
    if log_weights is None:
        raise ValueError("log_weights must be provided when variational=False")
    if method == ISMethod.PSIS:
        return psislw(log_weights, reff=reff)
    if method == ISMethod.SIS:
        return sislw(log_weights)
    if method == ISMethod.TIS:
        return tislw(log_weights)
    raise ValueError(f"Method {method} is not supported for standard importance sampling. "
                     "Use 'psis', 'sis', or 'tis' instead.")  # base.py:154-158

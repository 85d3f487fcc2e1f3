"""``e_loo`` -- importance-weighted expectations with their Pareto-k diagnostics.

Drop-in for ``pyloo.e_loo`` / ``compute_pareto_k`` / ``k_hat`` (reference: pyloo/e_loo.py:56-426), the direct
consumer of ``psislw``'s ``(N, S)`` weights.  The weighted moments (e_loo.py:429-463, :518-531) and the three
top-``tail_len`` tails + generalised Pareto fits of ``k_hat`` (e_loo.py:328-390) run in one CUDA kernel for the
whole batch (``b2l_eloo_dev_f64``); the three scalar diagnostics derived from k (e_loo.py:393-426) are O(N)
host arithmetic.  There is no CPU fallback for the kernel part.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np

from . import engine
from .data import SAMPLE_DIM, is_dataarray_like, to_inference_data, wrap_like

__all__ = ["e_loo", "ExpectationResult", "compute_pareto_k", "k_hat", "_pareto_min_ss",
           "_pareto_khat_threshold", "_pareto_convergence_rate"]


@dataclass
class ExpectationResult:
    """Results of an expectation calculation (pyloo/e_loo.py:24-53): ``value`` (with a trailing ``quantile``
    dimension for quantiles), the function-specific ``pareto_k``, and the derived ``min_ss``,
    ``khat_threshold`` and ``convergence_rate``."""

    value: Any
    pareto_k: Any
    min_ss: Any = None
    khat_threshold: Any = None
    convergence_rate: Any = None


# ------------------------------------------------------------------------------- scalar diagnostics
def _pareto_min_ss(k):
    """Minimum sample size for a reliable Pareto smoothed estimate (pyloo/e_loo.py:393-398)."""
    k = np.asarray(k, dtype=np.float64)
    with np.errstate(all="ignore"):
        out = np.where(k < 1, 10.0 ** (1.0 / (1.0 - np.maximum(0.0, np.where(k < 1, k, 0.0)))), np.inf)
    return out if out.ndim else float(out)


def _pareto_khat_threshold(n_samples: int) -> float:
    """k-hat threshold for a reliable Pareto smoothed estimate (pyloo/e_loo.py:401-403)."""
    return 1 - 1 / np.log10(n_samples)


def _pareto_convergence_rate(k, n_samples: int):
    """Relative convergence rate of the Pareto smoothed estimate (pyloo/e_loo.py:406-426)."""
    k = np.asarray(k, dtype=np.float64)
    n = float(n_samples)
    with np.errstate(all="ignore"):
        mid = (2 * (k - 1) * n ** (2 * k + 1) + (1 - 2 * k) * n ** (2 * k) + n**2) / ((n - 1) * (n - n ** (2 * k)))
        out = np.ones_like(k)                                   # k < 0, k == 0, k == 1, NaN
        inside = (k > 0) & (k < 1)
        out = np.where(inside, np.maximum(0.0, mid), out)
        out = np.where(k == 0.5, 1 - 1 / np.log(n), out)
        out = np.where(k > 1, 0.0, out)
    return out if out.ndim else float(out)


# ------------------------------------------------------------------------------- array plumbing
def _sample_last(obj, what):
    """``(values with the sample axis last, obs_dims or None)``; DataArray-likes need ``__sample__`` or
    ``chain`` + ``draw`` (stacked like e_loo.py:196-197, :203-208)."""
    if is_dataarray_like(obj):
        dims = tuple(obj.dims)
        if SAMPLE_DIM not in dims and "chain" in dims and "draw" in dims:
            obj = obj.stack(__sample__=("chain", "draw"))
            dims = tuple(obj.dims)
        vals = np.asarray(obj.values, dtype=np.float64)
        ax = dims.index(SAMPLE_DIM) if SAMPLE_DIM in dims else len(dims) - 1
        if ax != len(dims) - 1:
            vals = np.moveaxis(vals, ax, -1)
        return vals, dims[:ax] + dims[ax + 1:]
    vals = np.asarray(obj, dtype=np.float64)
    if vals.ndim < 1:
        raise ValueError(f"{what} must have at least one dimension")
    return vals, None


def _as_rows(vals, obs_shape, S, what):
    if vals.shape[-1] != S:
        raise ValueError(f"{what} has {vals.shape[-1]} samples, expected {S}")
    if vals.shape[:-1] != obs_shape:
        vals = np.broadcast_to(vals, (*obs_shape, S))
    return np.ascontiguousarray(vals.reshape(-1, S))


# ------------------------------------------------------------------------------- k_hat
def k_hat(x_vals, log_ratios_vals, tail_len: int = 20) -> float:
    """Pareto k of ``h(theta) = x_vals`` under raw log ratios for one observation (pyloo/e_loo.py:328-390).
    ``x_vals`` None fits the ratio tail only."""
    if tail_len < 5:
        raise ValueError("tail_len must be at least 5")
    lr = np.asarray(log_ratios_vals, dtype=np.float64).reshape(1, -1)
    if x_vals is None:
        _, k = engine.eloo_host(None, lr, None, "none", tail_len)
    else:
        x = np.asarray(x_vals, dtype=np.float64).reshape(1, -1)
        _, k = engine.eloo_host(x, lr, None, "mean", tail_len)
    return float(k[0])


def compute_pareto_k(x, log_ratios, tail_len: int = 20):
    """Batch Pareto k for expectation calculations (pyloo/e_loo.py:266-325): DataArray-likes with a
    ``__sample__`` dimension give a DataArray over the remaining dimensions, 1-D arrays give a float."""
    if tail_len < 5:
        raise ValueError("tail_len must be at least 5")
    if is_dataarray_like(log_ratios):
        if SAMPLE_DIM not in log_ratios.dims:
            raise ValueError("log_ratios must have '__sample__' dimension")
        if x is not None and is_dataarray_like(x) and SAMPLE_DIM not in x.dims:
            raise ValueError("x must have '__sample__' dimension")
        lr, obs_dims = _sample_last(log_ratios, "log_ratios")
        obs_shape, S = lr.shape[:-1], lr.shape[-1]
        lr2 = _as_rows(lr, obs_shape, S, "log_ratios")
        if x is not None and is_dataarray_like(x):
            xv, _ = _sample_last(x, "x")
            x2 = _as_rows(xv, obs_shape, S, "x")
        else:
            x2 = np.zeros_like(lr2)  # e_loo.py:307: constant h => ratio tail only
        _, k = engine.eloo_host(x2, lr2, None, "mean", tail_len)
        return wrap_like(log_ratios, k.reshape(obs_shape), obs_dims, "pareto_k")
    lr = np.asarray(log_ratios, dtype=np.float64)
    if x is not None and isinstance(x, np.ndarray) and x.shape != lr.shape:
        raise ValueError("x and log_ratios must have the same shape")
    return k_hat(x, lr, tail_len)


# ------------------------------------------------------------------------------- e_loo
def e_loo(data, var_name=None, group="posterior_predictive", weights=None, log_weights=None,
          log_ratios=None, type="mean", probs=None) -> ExpectationResult:  # noqa: A002 (reference signature)
    """Weighted mean / variance / sd of posterior(-predictive) draws under importance weights.

    Same parameters and errors as ``pyloo.e_loo`` (e_loo.py:56-263).  ``type="quantile"`` returns a trailing
    ``quantile`` dimension of ``len(probs)`` (e_loo.py:509-515) and supports up to 16384 draws per observation.
    """
    if type not in ["mean", "variance", "sd", "quantile"]:
        raise ValueError("type must be 'mean', 'variance', 'sd' or 'quantile'")
    if type == "quantile":
        if probs is None:
            raise ValueError("probs must be provided for quantile calculation")
        probs_array = np.array([probs]) if np.isscalar(probs) else np.asarray(probs)
        if not np.all((probs_array > 0) & (probs_array < 1)):
            raise ValueError("probs must be between 0 and 1")
    if weights is None and log_weights is None:
        raise ValueError("Either weights or log_weights must be provided")

    if is_dataarray_like(data):
        x_data = data
    else:
        idata = to_inference_data(data)
        if not hasattr(idata, group):
            raise ValueError(f"InferenceData object does not have a {group} group")
        data_group = getattr(idata, group)
        if var_name is None:
            var_names = list(data_group.data_vars)
            if len(var_names) == 1:
                var_name = var_names[0]
            else:
                raise ValueError(f"Multiple variables found in {group} group. Please specify var_name from: "
                                 f"{var_names}")
        elif var_name not in data_group.data_vars:
            raise ValueError(f"Variable '{var_name}' not found in {group} group. Available variables: "
                             f"{list(data_group.data_vars)}")
        x_data = data_group[var_name]

    xv, obs_dims = _sample_last(x_data, "data")
    obs_shape, S = xv.shape[:-1], xv.shape[-1]
    if weights is not None:  # e_loo.py:199-200
        wv, _ = _sample_last(weights, "weights")
        with np.errstate(all="ignore"):
            lwv = np.log(wv)
    else:
        lwv, _ = _sample_last(log_weights, "log_weights")
    x2 = _as_rows(xv, obs_shape, S, "data")
    lw2 = _as_rows(lwv, obs_shape, S, "log_weights")
    lr2 = None
    if log_ratios is not None:  # e_loo.py:229-230
        lrv, _ = _sample_last(log_ratios, "log_ratios")
        lr2 = _as_rows(lrv, obs_shape, S, "log_ratios")

    template = x_data if is_dataarray_like(x_data) else None
    if type == "quantile":  # e_loo.py:226-227, :232-233: h is None, k comes from the ratios alone
        value = engine.eloo_quantile_host(x2, lw2, probs_array).reshape(*obs_shape, len(probs_array))
        _, k = engine.eloo_host(None, lw2 if lr2 is None else lr2, None, "none", 20)
    else:
        value, k = engine.eloo_host(x2, lw2, lr2, type, 20)
        value = value.reshape(obs_shape)
    k = k.reshape(obs_shape)

    def wrap(arr, name, extra_dims=()):
        if template is None:
            return arr if arr.ndim else arr[()]
        return wrap_like(template, arr, (*obs_dims, *extra_dims), name)

    return ExpectationResult(
        value=wrap(value, getattr(x_data, "name", None), ("quantile",) if type == "quantile" else ()),
        pareto_k=wrap(k, "pareto_k"),
        min_ss=wrap(np.asarray(_pareto_min_ss(k)), "min_ss"),                          # e_loo.py:248
        khat_threshold=wrap(np.full(obs_shape, _pareto_khat_threshold(S)), "khat_threshold"),  # :249
        convergence_rate=wrap(np.asarray(_pareto_convergence_rate(k, S)), "convergence_rate"),  # :251-255
    )

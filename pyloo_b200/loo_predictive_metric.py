"""``loo_predictive_metric`` -- leave-one-out predictive error / accuracy metrics.

Drop-in for ``pyloo.loo_predictive_metric`` (reference: pyloo/loo_predictive_metric.py:22-372).  The reference
computes ``psislw(-log_lik)`` then ``e_loo(type="mean")`` (:208-218), materialising the ``(N, S)`` weights on the
host in between; here both stages run back to back on the GPU and the weights stay in device memory
(``engine.psis_expectation_host``).  The metric itself is O(N) host arithmetic like the reference's (:234-356).
"""

from __future__ import annotations

import numpy as np

from . import engine
from .data import to_inference_data
from .e_loo import _as_rows, _sample_last

__all__ = ["loo_predictive_metric"]

_METRICS = ("mae", "mse", "rmse", "acc", "balanced_acc")


def _validate_lengths(y, yhat) -> int:
    if len(y) != len(yhat):
        raise ValueError("y and yhat must have the same length")
    return len(y)


def _validate_binary_inputs(y, yhat) -> None:
    if not np.all((y <= 1) & (y >= 0)):
        raise ValueError("y must contain values between 0 and 1")
    if not np.all((yhat <= 1) & (yhat >= 0)):
        raise ValueError("yhat must contain values between 0 and 1")


def _mae(y, yhat):
    """Mean absolute error and its standard error (loo_predictive_metric.py:234-252)."""
    n = _validate_lengths(y, yhat)
    err = np.abs(y - yhat)
    return {"estimate": np.mean(err), "se": np.std(err, ddof=1) / np.sqrt(n)}


def _mse(y, yhat):
    """Mean squared error (loo_predictive_metric.py:255-273)."""
    n = _validate_lengths(y, yhat)
    err = (y - yhat) ** 2
    return {"estimate": np.mean(err), "se": np.std(err, ddof=1) / np.sqrt(n)}


def _rmse(y, yhat):
    """Root mean squared error, first-order delta-method SE (loo_predictive_metric.py:276-298)."""
    mse = _mse(y, yhat)
    return {"estimate": np.sqrt(mse["estimate"]), "se": np.sqrt(mse["se"] ** 2 / mse["estimate"] / 4)}


def _accuracy(y, yhat):
    """Classification accuracy at the 0.5 threshold (loo_predictive_metric.py:301-326)."""
    n = _validate_lengths(y, yhat)
    _validate_binary_inputs(y, yhat)
    hit = ((yhat > 0.5).astype(int) == y).astype(int)
    est = np.mean(hit)
    return {"estimate": est, "se": np.sqrt(est * (1 - est) / n)}


def _balanced_accuracy(y, yhat):
    """Mean of the per-class accuracies (loo_predictive_metric.py:329-356)."""
    n = _validate_lengths(y, yhat)
    _validate_binary_inputs(y, yhat)
    pred = (yhat > 0.5).astype(int)
    neg = y == 0
    tn = np.mean(pred[neg] == y[neg])
    tp = np.mean(pred[~neg] == y[~neg])
    return {"estimate": (tp + tn) / 2, "se": np.sqrt((tp * (1 - tp) + tn * (1 - tn)) / 4 / n)}


def loo_predictive_metric(data, y, var_name=None, group="posterior_predictive", log_lik_group="log_likelihood",
                          log_lik_var_name=None, metric="mae", r_eff=1.0, **kwargs):
    """Estimate a leave-one-out predictive metric; same parameters, errors and ``{"estimate", "se"}`` result
    as ``pyloo.loo_predictive_metric``."""
    y = np.asarray(y).flatten()
    idata = to_inference_data(data)
    if not hasattr(idata, group):
        raise ValueError(f"InferenceData object does not have a {group} group")
    if not hasattr(idata, log_lik_group):
        raise ValueError(f"InferenceData object does not have a {log_lik_group} group")
    pp_group = getattr(idata, group)
    ll_group = getattr(idata, log_lik_group)
    if log_lik_var_name is None:
        names = list(ll_group.data_vars)
        if len(names) == 1:
            log_lik_var_name = names[0]
        else:
            raise ValueError(f"Multiple variables found in {log_lik_group} group. Please specify "
                             f"log_lik_var_name from: {names}")
    elif log_lik_var_name not in ll_group.data_vars:
        raise ValueError(f"Variable '{log_lik_var_name}' not found in {log_lik_group} group. Available "
                         f"variables: {list(ll_group.data_vars)}")
    if var_name is None:  # the reference indexes pp_group[var_name] directly; a single variable is unambiguous
        names = list(pp_group.data_vars)
        if len(names) != 1:
            raise ValueError(f"Multiple variables found in {group} group. Please specify var_name from: {names}")
        var_name = names[0]
    elif var_name not in pp_group.data_vars:
        raise ValueError(f"Variable '{var_name}' not found in {group} group. Available variables: "
                         f"{list(pp_group.data_vars)}")

    xv, _ = _sample_last(pp_group[var_name], "data")
    llv, _ = _sample_last(ll_group[log_lik_var_name], "log_lik")
    obs_shape, S = xv.shape[:-1], xv.shape[-1]
    n_obs = obs_shape[0] if obs_shape else 1  # loo_predictive_metric.py:193-194
    if len(y) != n_obs:
        raise ValueError(f"Length of y ({len(y)}) must match the number of observations in x ({n_obs})")
    if metric not in _METRICS:
        raise ValueError(f"Invalid metric: {metric}. Must be one of: 'mae', 'mse', 'rmse', 'acc', 'balanced_acc'")

    x2 = _as_rows(xv, obs_shape, S, "data")
    lr2 = _as_rows(-llv, obs_shape, S, "log_lik")                   # log ratios = -log_lik (:208, :215)
    pred, _, _ = engine.psis_expectation_host(x2, lr2, float(r_eff), kwargs.get("type", "mean"))
    pred = pred.reshape(obs_shape)
    return {"mae": _mae, "mse": _mse, "rmse": _rmse, "acc": _accuracy,
            "balanced_acc": _balanced_accuracy}[metric](y, pred)

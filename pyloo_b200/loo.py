"""``loo`` -- PSIS-LOO-CV with the reference's signature and ELPDData output.

Drop-in for ``pyloo.loo`` with ``method`` in ``psis`` / ``sis`` / ``tis`` (reference: pyloo/loo.py:20-513).
What the reference does with five full-size NumPy temporaries and three per-observation Python loops
(loo.py:286-289, :319-324, :329-337) is one fused GPU pass here; for PSIS totals and standard errors come
from the device-side statistics record (loo.py:326-342).
"""

from __future__ import annotations

import warnings

import numpy as np

from . import engine
from .base import ISMethod
from .data import get_log_likelihood, sample_major, to_inference_data, wrap_like
from .elpd import ELPDData
from .ess import relative_efficiency
from .rcparams import rcParams

__all__ = ["loo"]

_SCALE_VALUE = {"deviance": -2, "log": 1, "negative_log": -1}  # loo.py:195-200


def _scale_value(scale):
    try:
        return _SCALE_VALUE[scale]
    except KeyError:
        raise TypeError('Valid scale values are "deviance", "log", "negative_log"') from None


def loo(data, pointwise=None, var_name=None, reff=None, scale=None, method="psis", moment_match=False,
        jacobian=None, mixture=False, **kwargs):
    """Pareto-smoothed importance sampling leave-one-out cross-validation (PSIS-LOO-CV).

    Same parameters, errors, warnings and ``ELPDData`` rows as ``pyloo.loo`` for ``method`` in
    ``"psis"`` / ``"sis"`` / ``"tis"``.  ``moment_match`` and ``mixture`` are outside the accelerated path
    and raise ``NotImplementedError``.
    """
    idata = to_inference_data(data)
    log_lik = get_log_likelihood(idata, var_name=var_name)
    pointwise = rcParams["stats.ic_pointwise"] if pointwise is None else pointwise
    if jacobian is not None and not pointwise:
        raise ValueError("Jacobian adjustment requires pointwise LOO results. "
                         "Please set pointwise=True when using jacobian_adjustment.")

    ll_sn, obs_dims, obs_shape = sample_major(log_lik)  # loo.py:189 (no copy for the ArviZ layout)
    n_samples = ll_sn.shape[0]
    n_data_points = np.prod(obs_shape) if obs_shape else np.int64(1)  # loo.py:192
    scale = rcParams["stats.ic_scale"] if scale is None else scale.lower()
    sv = _scale_value(scale)

    if reff is None:  # loo.py:204-216
        if not hasattr(idata, "posterior"):
            raise TypeError("Must be able to extract a posterior group from data.")
        posterior = idata.posterior
        n_chains = len(posterior.chain)
        reff = 1.0 if n_chains == 1 else relative_efficiency(posterior, n_samples)

    try:  # loo.py:229-233
        method = method if isinstance(method, ISMethod) else ISMethod(method.lower())
    except ValueError:
        valid = ", ".join(m.value for m in ISMethod)
        raise ValueError(f"Invalid method '{method}'. Must be one of: {valid}")
    if method != ISMethod.PSIS:  # loo.py:235-244
        warnings.warn(f"Using {method.value.upper()} for LOO computation. Note that PSIS is the recommended "
                      "method as it is typically more efficient and reliable.", UserWarning, stacklevel=2)
    if mixture:
        raise NotImplementedError("mixture=True (Mix-IS-LOO) is outside the B200 hot path")
    if moment_match:
        if not pointwise:
            raise ValueError("Moment matching requires pointwise LOO results. "
                             "Please set pointwise=True when using moment_match=True.")
        raise NotImplementedError("moment_match=True is outside the B200 hot path")

    if method != ISMethod.PSIS:
        return _loo_is(log_lik, ll_sn, obs_dims, obs_shape, method, sv, scale, pointwise, jacobian,
                       n_samples, n_data_points)

    res = engine.loo_host(ll_sn, reff)
    st = res["stats"]
    if st.n_nan_in > 0:  # loo.py:218-227 (the -1e10 replacement happens inside the kernel)
        warnings.warn("NaN values detected in log-likelihood. These will be ignored in the LOO calculation.",
                      UserWarning, stacklevel=2)

    good_k = res["good_k"]  # loo.py:249
    warn_mg = False
    if st.k_gt_good > 0:  # loo.py:291-304
        warnings.warn(
            f"Estimated shape parameter of Pareto distribution is greater than {good_k:.2f} for "
            f"{int(st.k_gt_good)} observations. This indicates that importance sampling may be unreliable "
            "because the marginal posterior and LOO posterior are very different.",
            UserWarning, stacklevel=2)
        warn_mg = True

    n_obs = float(st.n)
    loo_lppd = sv * st.elpd_sum                              # loo.py:326
    var_i = st.elpd_m2 / n_obs * sv * sv                     # np.var(loo_i), ddof 0
    loo_lppd_se = float((n_data_points * var_i) ** 0.5)      # loo.py:327
    lppd = st.lppd_sum                                       # loo.py:329-337
    p_loo = lppd - loo_lppd / sv                             # loo.py:339
    p_loo_se = float(np.sqrt(var_i))                         # loo.py:340
    looic = -2 * loo_lppd                                    # loo.py:341 (already-scaled total, App. D)
    looic_se = 2 * loo_lppd_se                               # loo.py:342

    rows = [("elpd_loo", loo_lppd), ("se", loo_lppd_se), ("p_loo", p_loo), ("p_loo_se", p_loo_se),
            ("n_samples", n_samples), ("n_data_points", n_data_points), ("warning", warn_mg)]
    if not pointwise:  # loo.py:344-367, :554-577
        rows += [("scale", scale), ("looic", looic), ("looic_se", looic_se), ("good_k", good_k),
                 ("subsample_size", n_data_points)]
        return ELPDData(data=[v for _, v in rows], index=[k for k, _ in rows])

    loo_i = (sv * res["elpd_i"]).reshape(obs_shape)
    pareto_k = res["pareto_k"].reshape(obs_shape)
    if np.allclose(loo_i, loo_i.flat[0]):  # loo.py:377-382
        warnings.warn("The point-wise LOO is the same with the sum LOO, please double check the Observed RV "
                      "in your model to make sure it returns element-wise logp.", stacklevel=2)
    loo_i_da = wrap_like(log_lik, loo_i, obs_dims, "loo_i")
    k_da = wrap_like(log_lik, pareto_k, obs_dims, "pareto_shape")
    rows += [("loo_i", loo_i_da), ("scale", scale), ("looic", looic), ("looic_se", looic_se),
             ("pareto_k", k_da), ("good_k", good_k), ("subsample_size", n_data_points)]  # loo.py:599-624, :400-410
    result = ELPDData(data=[v for _, v in rows], index=[k for k, _ in rows])

    return _apply_jacobian(result, jacobian, lppd, sv, n_data_points)


def _apply_jacobian(result, jacobian, lppd, sv, n_data_points):
    if jacobian is not None:  # loo.py:414-439 (host-only add)
        adj = np.asarray(jacobian)
        if adj.shape != result["loo_i"].shape:
            raise ValueError(f"Jacobian adjustment shape {adj.shape} does not match loo_i shape "
                             f"{result['loo_i'].shape}")
        vals = result["loo_i"].values + adj
        result["loo_i"].values = vals
        total = vals.sum()
        total_se = (n_data_points * np.var(vals)) ** 0.5
        result["elpd_loo"] = total
        result["se"] = total_se
        result["p_loo"] = lppd - total / sv
        result["p_loo_se"] = np.sqrt(np.sum(np.var(vals)))
        result["looic"] = -2 * total
        result["looic_se"] = 2 * total_se
    return result


def _loo_is(log_lik, ll_sn, obs_dims, obs_shape, method, sv, scale, pointwise, jacobian, n_samples,
            n_data_points):
    """SIS / TIS branch (loo.py:286-289, :305-342): the pointwise pass runs on the GPU, the totals are the
    reference's NumPy reductions over the N pointwise values."""
    res = engine.loo_is_host(ll_sn, method.value)
    if res["n_nan_in"] > 0:  # loo.py:218-227
        warnings.warn("NaN values detected in log-likelihood. These will be ignored in the LOO calculation.",
                      UserWarning, stacklevel=3)
    warn_mg = False
    min_ess = np.min(res["ess_i"])
    if min_ess < n_samples * 0.1:  # loo.py:306-317
        warnings.warn(f"Low effective sample size detected (minimum ESS: {min_ess:.1f}). This indicates that "
                      "the importance sampling approximation may be unreliable. Consider using PSIS which "
                      "is more robust to such cases.", UserWarning, stacklevel=3)
        warn_mg = True
    loo_i = sv * res["elpd_i"]
    loo_lppd = loo_i.sum()                                   # loo.py:326
    loo_lppd_se = (n_data_points * np.var(loo_i)) ** 0.5     # loo.py:327
    lppd = np.sum(res["lppd_i"])                             # loo.py:329-337
    p_loo = lppd - loo_lppd / sv                             # loo.py:339
    p_loo_se = np.sqrt(np.sum(np.var(loo_i)))                # loo.py:340
    looic = -2 * loo_lppd
    looic_se = 2 * loo_lppd_se
    rows = [("elpd_loo", loo_lppd), ("se", loo_lppd_se), ("p_loo", p_loo), ("p_loo_se", p_loo_se),
            ("n_samples", n_samples), ("n_data_points", n_data_points), ("warning", warn_mg)]
    if not pointwise:
        rows += [("scale", scale), ("looic", looic), ("looic_se", looic_se),
                 ("subsample_size", n_data_points)]  # no good_k row outside PSIS (loo.py:360-365)
        return ELPDData(data=[v for _, v in rows], index=[k for k, _ in rows])
    loo_i = loo_i.reshape(obs_shape)
    if np.allclose(loo_i, loo_i.flat[0]):  # loo.py:377-382
        warnings.warn("The point-wise LOO is the same with the sum LOO, please double check the Observed RV "
                      "in your model to make sure it returns element-wise logp.", stacklevel=3)
    loo_i_da = wrap_like(log_lik, loo_i, obs_dims, "loo_i")
    ess_da = wrap_like(log_lik, res["ess_i"].reshape(obs_shape), obs_dims, "ess")
    rows += [("loo_i", loo_i_da), ("scale", scale), ("looic", looic), ("looic_se", looic_se),
             ("ess", ess_da), ("subsample_size", n_data_points)]  # loo.py:400-410
    result = ELPDData(data=[v for _, v in rows], index=[k for k, _ in rows])
    return _apply_jacobian(result, jacobian, lppd, sv, n_data_points)

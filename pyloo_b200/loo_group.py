"""``loo_group`` -- leave-one-group-out cross-validation (LOGO-CV).

Drop-in for ``pyloo.loo_group`` (reference: pyloo/loo_group.py:19-380).  The reference sums the
log-likelihood of every group on the host (:215-222), loops over the groups calling
``compute_importance_weights`` one at a time (:226-233) and loops again for the two logsumexps (:281-305).
Here the group sums are one CUDA kernel over the whole ``(S, N)`` matrix and the result -- G rows of S
doubles, already row-contiguous -- goes through the same fused LOO pass as ``loo`` with one
"observation" per group.
"""

from __future__ import annotations

import warnings

import numpy as np

from . import engine
from .base import ISMethod
from .data import get_log_likelihood, sample_major, to_inference_data
from .data import LiteDataArray
from .elpd import ELPDData
from .ess import relative_efficiency
from .loo import _scale_value
from .rcparams import rcParams

__all__ = ["loo_group"]


def loo_group(data, group_ids, pointwise=None, var_name=None, reff=None, scale=None, method="psis"):
    """Leave-one-group-out CV; same parameters, errors, warnings and ``ELPDData`` rows (``elpd_logo``, ``se``,
    ``p_logo``, ``p_logo_se``, ``n_samples``, ``n_groups``, ``warning``, [``logo_i``,] ``scale``, ``logoic``,
    ``logoic_se``, [``pareto_k`` | ``ess``,] [``good_k``]) as ``pyloo.loo_group``."""
    idata = to_inference_data(data)
    log_lik = get_log_likelihood(idata, var_name=var_name)
    pointwise = rcParams["stats.ic_pointwise"] if pointwise is None else pointwise
    ll_sn, _, obs_shape = sample_major(log_lik)  # loo_group.py:150
    n_samples = ll_sn.shape[0]
    n_data_points = int(np.prod(obs_shape)) if obs_shape else 1
    scale = rcParams["stats.ic_scale"] if scale is None else scale.lower()
    group_ids = np.asarray(group_ids)
    if len(group_ids) != n_data_points:  # loo_group.py:156-160
        raise ValueError(f"Length of group_ids ({len(group_ids)}) must match the number of observations in "
                         f"log_likelihood ({n_data_points}).")
    unique_groups, inverse = np.unique(group_ids, return_inverse=True)  # loo_group.py:162-163
    n_groups = len(unique_groups)
    sv = _scale_value(scale)

    if reff is None:  # loo_group.py:174-186
        if not hasattr(idata, "posterior"):
            raise TypeError("Must be able to extract a posterior group from data.")
        posterior = idata.posterior
        reff = 1.0 if len(posterior.chain) == 1 else relative_efficiency(posterior, n_samples)

    try:  # loo_group.py:199-203
        method = method if isinstance(method, ISMethod) else ISMethod(method.lower())
    except ValueError:
        valid = ", ".join(m.value for m in ISMethod)
        raise ValueError(f"Invalid method '{method}'. Must be one of: {valid}")
    if method != ISMethod.PSIS:
        warnings.warn(f"Using {method.value.upper()} for LOGO computation. Note that PSIS is the recommended "
                      "method as it is typically more efficient and reliable.", UserWarning, stacklevel=2)

    res = engine.group_loo_host(ll_sn, np.asarray(inverse).reshape(-1), n_groups, reff, method.value)
    if res["n_nan_in"] > 0:  # loo_group.py:188-197 (replaced by -1e10 inside the group-sum kernel)
        warnings.warn("NaN values detected in log-likelihood. These will be ignored in the LOGO calculation.",
                      UserWarning, stacklevel=2)

    good_k = min(1 - 1 / np.log10(n_samples), 0.7)
    warn_mg = False
    if method == ISMethod.PSIS:  # loo_group.py:241-254
        diagnostics = res["pareto_k"]
        if np.any(diagnostics > good_k):
            warnings.warn(
                f"Estimated shape parameter of Pareto distribution is greater than {good_k:.2f} for "
                f"{np.sum(diagnostics > good_k)} groups. This indicates that importance sampling may be "
                "unreliable because the marginal posterior and LOGO posterior are very different.", UserWarning,
                stacklevel=2)
            warn_mg = True
    else:  # loo_group.py:255-267
        diagnostics = res["ess_i"]
        min_ess = np.min(diagnostics)
        if min_ess < n_samples * 0.1:
            warnings.warn(f"Low effective sample size detected (minimum ESS: {min_ess:.1f}). This indicates "
                          "that the importance sampling approximation may be unreliable. Consider using PSIS "
                          "which is more robust to such cases.", UserWarning, stacklevel=2)
            warn_mg = True

    logo_i = sv * res["elpd_i"]
    logo_lppd = logo_i.sum()                                  # loo_group.py:283
    logo_lppd_se = (n_groups * np.var(logo_i)) ** 0.5         # :284
    lppd = res["lppd_i"].sum()                                # :298
    p_logo = lppd - logo_lppd / sv                            # :300
    p_logo_se = np.sqrt(np.sum(np.var(logo_i)))               # :301
    logoic = -2 * logo_lppd
    logoic_se = 2 * logo_lppd_se
    rows = [("elpd_logo", logo_lppd), ("se", logo_lppd_se), ("p_logo", p_logo), ("p_logo_se", p_logo_se),
            ("n_samples", n_samples), ("n_groups", n_groups), ("warning", warn_mg)]
    if pointwise:
        rows.append(("logo_i", LiteDataArray(logo_i, ("group",), name="logo_i", coords={"group": unique_groups})))
    rows += [("scale", scale), ("logoic", logoic), ("logoic_se", logoic_se)]
    if pointwise:
        rows.append(("pareto_k" if method == ISMethod.PSIS else "ess", diagnostics))
    if method == ISMethod.PSIS:
        rows.append(("good_k", good_k))
    return ELPDData(data=[v for _, v in rows], index=[k for k, _ in rows])

"""``waic`` -- widely applicable information criterion with the reference's signature.

Drop-in for ``pyloo.waic`` (reference: pyloo/waic.py:16-207).  ``lppd_i`` (waic.py:137-143) and the
ddof-0 posterior variance (waic.py:145) come out of the same fused GPU pass as ``loo`` (run with
the PSIS stage switched off); sums and SE come from the device statistics record (waic.py:157-160).
"""

from __future__ import annotations

import warnings

import numpy as np

from . import engine
from .data import get_log_likelihood, sample_major, to_inference_data, wrap_like
from .elpd import ELPDData
from .loo import _scale_value
from .rcparams import rcParams

__all__ = ["waic"]


def waic(data, pointwise=None, var_name=None, scale=None):
    """Compute WAIC; same parameters, warnings and ``ELPDData`` rows as ``pyloo.waic``."""
    idata = to_inference_data(data)
    log_lik = get_log_likelihood(idata, var_name=var_name)
    pointwise = rcParams["stats.ic_pointwise"] if pointwise is None else pointwise
    ll_sn, obs_dims, obs_shape = sample_major(log_lik)  # waic.py:95
    n_samples = ll_sn.shape[0]
    n_data_points = np.prod(obs_shape) if obs_shape else np.int64(1)
    scale = rcParams["stats.ic_scale"] if scale is None else scale.lower()
    sv = _scale_value(scale)

    res = engine.loo_host(ll_sn, 1.0, waic_only=True)
    st = res["stats"]
    if st.n_nan_in > 0:  # waic.py:113-120
        warnings.warn("NaN values detected in log-likelihood. These will be ignored in the WAIC calculation.",
                      UserWarning, stacklevel=2)
    if st.n_pinf_in + st.n_ninf_in > 0:  # waic.py:122-132
        warnings.warn("Infinite values detected in log-likelihood. These will be ignored in the WAIC "
                      "calculation.", UserWarning, stacklevel=2)
    warn_mg = bool(st.var_gt_04 > 0)  # waic.py:147
    if warn_mg:
        warnings.warn("For one or more samples the posterior variance of the log predictive densities "
                      "exceeds 0.4. This could be indication of WAIC starting to fail.", UserWarning,
                      stacklevel=2)

    n_obs = float(st.n)
    waic_sum = sv * st.waic_sum                                         # waic.py:159
    waic_se = float((n_data_points * (st.waic_m2 / n_obs) * sv * sv) ** 0.5)  # waic.py:158
    p_waic = st.p_waic_sum                                              # waic.py:160
    if not pointwise:
        return ELPDData(data=[waic_sum, waic_se, p_waic, n_samples, n_data_points, warn_mg, scale],
                        index=["elpd_waic", "se", "p_waic", "n_samples", "n_data_points", "warning", "scale"])
    waic_i = (sv * (res["lppdw_i"] - res["var_i"])).reshape(obs_shape)  # waic.py:157
    if np.allclose(waic_i, waic_i.flat[0]):  # waic.py:176-182
        warnings.warn("The point-wise WAIC is the same with the sum WAIC, please double check the Observed "
                      "RV in your model to make sure it returns element-wise logp.", UserWarning, stacklevel=2)
    waic_i_da = wrap_like(log_lik, waic_i, obs_dims, "waic_i")
    return ELPDData(data=[waic_sum, waic_se, p_waic, n_samples, n_data_points, warn_mg, waic_i_da, scale],
                    index=["elpd_waic", "se", "p_waic", "n_samples", "n_data_points", "warning", "waic_i",
                           "scale"])

"""CPU-only tests: host logic of the reference-facing API (no compute calls), the C-ABI library
loading and exporting every symbol include/psisloo_b200.h declares, data ingress, ELPDData."""

import ctypes
import os
import re

import numpy as np
import pytest

import pyloo_b200 as pl
from pyloo_b200 import _native, engine
from pyloo_b200.data import LiteDataArray, from_dict, get_log_likelihood, sample_major, to_inference_data
from pyloo_b200.ess import ess_mean
from b2l_testutil import ROOT, has_cuda


def test_library_builds_loads_and_exports_header_symbols():
    _native.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "psisloo_b200.h")).read()
    declared = set(re.findall(r"\b(b2l_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b2l_version() == 100


def test_no_cpu_fallback_without_gpu():
    if has_cuda():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pl.psislw(np.random.default_rng(0).normal(size=(4, 100)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.loo_host(np.zeros((100, 3)))
    # the widened rows fail just as loudly
    x = np.random.default_rng(1).normal(size=(3, 64))
    for call in (lambda: pl.sislw(x), lambda: pl.tislw(x), lambda: pl.compute_importance_weights(x, method="tis"),
                 lambda: pl.k_hat(x[0], x[1]), lambda: engine.eloo_quantile_host(x, x, [0.5]),
                 lambda: engine.group_loo_host(x.T, [0, 1, 0], 2), lambda: engine.psis_expectation_host(x, x)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_stats_merge_is_host_arithmetic():
    rng = np.random.default_rng(0)
    x = rng.normal(size=1000)
    recs = []
    for part in (x[:300], x[300:]):
        r = np.zeros(32)
        r[0] = len(part); r[1] = part.mean(); r[2] = ((part - part.mean()) ** 2).sum(); r[3] = part.sum()
        r[6] = part.mean(); r[7] = r[2]; r[8] = part.sum()
        r[14] = r[16] = part.min(); r[15] = r[17] = part.max()
        recs.append(r)
    m = engine.stats_merge(recs)
    assert m.n == 1000
    np.testing.assert_allclose(m.elpd_mean, x.mean(), rtol=1e-13)
    np.testing.assert_allclose(m.elpd_m2, ((x - x.mean()) ** 2).sum(), rtol=1e-12)
    np.testing.assert_allclose(m.elpd_sum, x.sum(), rtol=1e-13)
    assert m.elpd_min == x.min() and m.elpd_max == x.max()


def test_tail_length_and_good_k_follow_reference_expressions():
    assert engine.tail_length(4000, 0.9) == 200       # pyloo/psis.py:89
    assert engine.tail_length(2000, 1.0) == 135
    assert engine.good_k_threshold(2000) == 0.7 - 0.0 if 1 - 1 / np.log10(2000) > 0.7 else True
    assert engine.good_k_threshold(100) == pytest.approx(0.5)   # pyloo/loo.py:249
    assert engine.CUTOFFMIN == pytest.approx(-708.3964185322641)


def test_argument_errors_match_reference_classes():
    idata = from_dict(posterior={"mu": np.zeros((2, 50))}, log_likelihood={"y": np.zeros((2, 50, 3))})
    with pytest.raises(TypeError, match="Valid scale values"):      # test_loo.py:64-68
        pl.loo(idata, scale="bad", reff=1.0)
    with pytest.raises(ValueError, match="Invalid method"):          # test_loo.py:219-221
        pl.loo(idata, method="nope", reff=1.0)
    with pytest.raises(ValueError, match="Jacobian adjustment requires pointwise"):
        pl.loo(idata, jacobian=np.zeros(3), pointwise=False, reff=1.0)
    no_ll = from_dict(posterior={"mu": np.zeros((2, 50))})
    with pytest.raises(TypeError, match="log likelihood not found"):  # test_loo.py:71-74
        pl.loo(no_ll)
    two = from_dict(log_likelihood={"a": np.zeros((2, 50, 3)), "b": np.zeros((2, 50, 3))})
    with pytest.raises(TypeError, match="Found several log likelihood arrays"):  # test_loo.py:190-198
        pl.loo(two, reff=1.0)
    with pytest.raises(TypeError, match="No log likelihood data named"):
        pl.loo(two, var_name="zzz", reff=1.0)
    only_ll = from_dict(log_likelihood={"y": np.zeros((2, 50, 3))})
    with pytest.raises(TypeError, match="Must be able to extract a posterior"):  # test_loo.py:77-86
        pl.loo(only_ll)
    with pytest.raises(ValueError, match="Invalid method"):          # base.py:100-107
        pl.compute_importance_weights(np.zeros((3, 10)), method="xx")
    with pytest.raises(ValueError, match="log_weights must be provided"):
        pl.compute_importance_weights(None)
    with pytest.raises(ValueError, match="__sample__"):
        pl.compute_importance_weights(LiteDataArray(np.zeros((3, 10)), ("a", "b")))
    with pytest.raises(TypeError, match="must be a dictionary"):     # test_compare.py:188-216
        pl.loo_compare([1, 2])
    with pytest.raises(ValueError, match="at least two models"):
        pl.loo_compare({"a": idata})
    with pytest.raises(ValueError, match="Scale must be"):
        pl.loo_compare({"a": idata, "b": idata}, scale="x")
    with pytest.raises(ValueError, match="Method must be"):
        pl.loo_compare({"a": idata, "b": idata}, method="x")
    with pytest.raises(ValueError, match="ic must be"):
        pl.loo_compare({"a": idata, "b": idata}, ic="x")
    with pytest.raises(ValueError, match="Lists and tuples"):
        to_inference_data([1, 2, 3])


def test_sample_major_is_a_view_in_stack_order():
    rng = np.random.default_rng(1)
    arr = rng.normal(size=(4, 25, 3, 2))
    da = LiteDataArray(arr, ("chain", "draw", "d1", "d2"))
    mat, obs_dims, obs_shape = sample_major(da)
    assert mat.shape == (100, 6) and obs_dims == ("d1", "d2") and obs_shape == (3, 2)
    assert np.shares_memory(mat, arr)
    stacked = da.stack(__sample__=("chain", "draw"))          # (d1, d2, sample): chain outer, draw inner
    assert stacked.dims == ("d1", "d2", "__sample__")
    assert np.array_equal(stacked.values.reshape(6, 100).T, mat)
    tr = da.transpose("draw", "chain", ...)                   # transposed model (helpers.py:79-83)
    mat2, _, _ = sample_major(tr)
    assert np.array_equal(mat2, mat)
    assert hasattr(stacked, "__sample__") and len(stacked.__sample__) == 100   # psis.py:79-80


def test_get_log_likelihood_rules():
    idata = from_dict(log_likelihood={"obs": np.zeros((2, 10, 4))})
    assert get_log_likelihood(idata).name == "obs"
    assert get_log_likelihood(idata, "obs").shape == (2, 10, 4)


def test_rcparams_contract():
    rc = pl.rcParams
    assert rc["stats.ic_pointwise"] is False and rc["stats.ic_scale"] == "log"
    with pytest.raises(ValueError):
        rc["stats.ic_scale"] = "bogus"
    with pytest.raises(KeyError):
        rc["nope"] = 1
    with pytest.raises(TypeError):
        del rc["stats.ic_scale"]
    rc["stats.ic_scale"] = "Deviance"
    assert rc["stats.ic_scale"] == "deviance"
    rc["stats.ic_scale"] = "log"
    assert sorted(rc) == ["plot.backend", "stats.ic_pointwise", "stats.ic_scale"]


def test_elpddata_report_matches_reference_layout():
    k = np.array([0.1, 0.8, 1.2, 0.3, 0.2, 0.1, 0.0, 0.5])
    e = pl.ELPDData(
        data=[-30.78, 1.35, 0.95, 0.48, 2000, 8, True, LiteDataArray(np.zeros(8), ("obs",)), "log", 61.56, 2.69,
              LiteDataArray(k, ("obs",)), 0.7, 8],
        index=["elpd_loo", "se", "p_loo", "p_loo_se", "n_samples", "n_data_points", "warning", "loo_i", "scale",
               "looic", "looic_se", "pareto_k", "good_k", "subsample_size"])
    text = str(e)
    assert "Computed from 2000 posterior samples and 8 observations log-likelihood matrix." in text
    assert "elpd_loo   -30.78      1.35" in text and "p_loo       0.95        0.48" in text
    assert "looic      61.56       2.69" in text
    assert "(-Inf, 0.70]   (good)      6   75.0%" in text and "(1, Inf)   (very bad)    1    12.5%" in text
    assert "There has been a warning during the calculation." in text
    assert e.n_samples == 2000 and e.n_data_points == 8 and e.warning
    ref_path = "/root/reference/pyloo/elpd.py"
    if os.path.exists(ref_path):  # same text as the reference's own class (build container only)
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_elpd", ref_path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ref = mod.ELPDData(data=list(e.values), index=list(e.index))
        assert str(ref) == text
        e2 = pl.ELPDData(data=list(e.values)[:11] + [0.7, 8], index=list(e.index)[:11] + ["good_k", "subsample_size"])
        ref2 = mod.ELPDData(data=list(e2.values), index=list(e2.index))
        assert str(ref2) == str(e2)


def test_ess_mean_sane():
    rng = np.random.default_rng(3)
    iid = rng.normal(size=(4, 1000))
    assert 0.7 * 4000 < ess_mean(iid) < 1.4 * 4000
    ar = np.zeros((4, 2000))
    eps = rng.normal(size=(4, 2000))
    for t in range(1, 2000):
        ar[:, t] = 0.9 * ar[:, t - 1] + eps[:, t]
    # AR(1) with phi = 0.9: ESS ~ N (1 - phi) / (1 + phi) ~ 0.053 N
    assert 0.02 * 8000 < ess_mean(ar) < 0.12 * 8000


# ------------------------------------------------------------------ widened rows: host-side contracts (no GPU)
def _small_idata():
    rng = np.random.default_rng(3)
    ll = rng.normal(size=(2, 50, 6))
    return from_dict(posterior={"mu": rng.normal(size=(2, 50))}, log_likelihood={"y": ll},
                     posterior_predictive={"y": rng.normal(size=(2, 50, 6)), "z": rng.normal(size=(2, 50, 6))},
                     observed_data={"y": rng.normal(size=6)}, dims={"y": ["obs"], "z": ["obs"]})


def test_e_loo_argument_errors_precede_any_device_work():
    idata = _small_idata()
    lw = LiteDataArray(np.zeros((6, 100)), ("obs", "__sample__"))
    with pytest.raises(ValueError, match="type must be 'mean', 'variance', 'sd' or 'quantile'"):
        pl.e_loo(idata, var_name="y", log_weights=lw, type="mode")                  # e_loo.py:151-152
    with pytest.raises(ValueError, match="probs must be provided"):
        pl.e_loo(idata, var_name="y", log_weights=lw, type="quantile")              # :155-156
    with pytest.raises(ValueError, match="probs must be between 0 and 1"):
        pl.e_loo(idata, var_name="y", log_weights=lw, type="quantile", probs=[0.0, 0.5])
    with pytest.raises(ValueError, match="Either weights or log_weights must be provided"):
        pl.e_loo(idata, var_name="y")                                               # :166-167
    with pytest.raises(ValueError, match="does not have a prior group"):
        pl.e_loo(idata, group="prior", log_weights=lw)                              # :174-175
    with pytest.raises(ValueError, match="Multiple variables found in posterior_predictive group"):
        pl.e_loo(idata, log_weights=lw)                                             # :183-187
    with pytest.raises(ValueError, match="Variable 'q' not found"):
        pl.e_loo(idata, var_name="q", log_weights=lw)                               # :188-192
    with pytest.raises(ValueError, match="tail_len must be at least 5"):
        pl.compute_pareto_k(np.zeros(10), np.zeros(10), tail_len=4)                 # :295-296
    with pytest.raises(ValueError, match="log_ratios must have '__sample__' dimension"):
        pl.compute_pareto_k(None, LiteDataArray(np.zeros((3, 10)), ("a", "b")))     # :299-300


def test_pareto_diagnostics_of_e_loo():
    from pyloo_b200.e_loo import _pareto_convergence_rate, _pareto_khat_threshold, _pareto_min_ss

    assert _pareto_min_ss(0.5) == pytest.approx(100.0) and _pareto_min_ss(-1.0) == pytest.approx(10.0)
    assert _pareto_min_ss(1.0) == np.inf and _pareto_min_ss(np.nan) == np.inf      # e_loo.py:393-398
    assert _pareto_khat_threshold(1000) == pytest.approx(1 - 1 / 3)                # :401-403
    assert _pareto_convergence_rate(-0.1, 100) == 1.0 and _pareto_convergence_rate(1.5, 100) == 0.0
    assert _pareto_convergence_rate(0.5, 100) == pytest.approx(1 - 1 / np.log(100))  # :415-416
    k, n = 0.3, 400
    expect = max(0, (2 * (k - 1) * n ** (2 * k + 1) + (1 - 2 * k) * n ** (2 * k) + n**2) / ((n - 1) * (n - n ** (2 * k))))
    assert _pareto_convergence_rate(k, n) == pytest.approx(expect, rel=1e-14)
    np.testing.assert_allclose(_pareto_convergence_rate(np.array([-1.0, 0.0, 1.0, 2.0]), 50), [1.0, 1.0, 1.0, 0.0])


def test_loo_group_and_metric_argument_errors():
    idata = _small_idata()
    with pytest.raises(ValueError, match=r"Length of group_ids \(3\) must match the number of observations"):
        pl.loo_group(idata, [0, 1, 2], reff=1.0)                                     # loo_group.py:156-160
    with pytest.raises(TypeError, match="Valid scale values"):
        pl.loo_group(idata, np.arange(6) % 2, reff=1.0, scale="bits")                # :171-172
    with pytest.raises(ValueError, match="Invalid method 'xx'"):
        pl.loo_group(idata, np.arange(6) % 2, reff=1.0, method="xx")                 # :199-203
    y = np.zeros(6)
    with pytest.raises(ValueError, match="does not have a nope group"):
        pl.loo_predictive_metric(idata, y, var_name="y", group="nope")               # loo_predictive_metric.py:158-159
    with pytest.raises(ValueError, match="Variable 'q' not found in log_likelihood group"):
        pl.loo_predictive_metric(idata, y, var_name="y", log_lik_var_name="q")       # :174-178
    with pytest.raises(ValueError, match=r"Length of y \(2\) must match"):
        pl.loo_predictive_metric(idata, y[:2], var_name="y")                         # :196-200
    with pytest.raises(ValueError, match="Invalid metric: f1"):
        pl.loo_predictive_metric(idata, y, var_name="y", metric="f1")                # :202-206
    with pytest.raises(ValueError, match="Multiple variables found in posterior_predictive group"):
        pl.loo_score(idata, y_var="y")                                               # loo_score.py:458-462
    with pytest.raises(ValueError, match="Variable 'q' not found in posterior_predictive group"):
        pl.loo_score(idata, x_var="y", x2_var="q", y_var="y")                        # :485-489
    with pytest.raises(ValueError, match="does not have a nope group"):
        pl.loo_score(idata, x_var="y", y_group="nope")                               # :493-494
    with pytest.raises(ValueError, match="Invalid method 'nope'"):
        pl.compute_importance_weights(np.zeros((2, 10)), method="nope")              # base.py:100-107
    with pytest.raises(ValueError, match="log_weights must have a __sample__ dimension"):
        pl.compute_importance_weights(LiteDataArray(np.zeros((2, 10)), ("a", "b")), method="sis")  # base.py:93-98


def test_elpddata_logo_report():
    e = pl.ELPDData(data=[-12.5, 1.25, 3.5, 0.4, 1000, 10, True, "log", 25.0, 2.5, np.array([0.1, 0.9, 1.4] + [0.2] * 7), 0.67],
                    index=["elpd_logo", "se", "p_logo", "p_logo_se", "n_samples", "n_groups", "warning", "scale",
                           "logoic", "logoic_se", "pareto_k", "good_k"])
    text = str(e)
    assert "Computed from 1000 posterior samples and 10 groups log-likelihood matrix." in text   # elpd.py:74-81
    assert "elpd_logo   -12.50      1.25" in text and "p_logo       3.50        0.40" in text
    assert "logoic      25.00       2.50" in text
    assert "There has been a warning during the calculation." in text
    assert "(-Inf, 0.67]   (good)      8   80.0%" in text and "(1, Inf)   (very bad)    1    10.0%" in text
    assert e.n_groups == 10

"""GPU parity tests for the callers either side of psislw: SIS / TIS weights,
``loo(method="sis"|"tis")`` and ``e_loo`` -- the CUDA kernels (through the C ABI) against the golden vectors
of the real reference (tests/golden/is_eloo.npz) and against the CPU oracle on seeded inputs.

Tolerance: 1e-10 relative (BASELINE.json) on log weights, ess, elpd_i, lppd_i, weighted moments; Pareto k of
``k_hat`` to 1e-12; identical NaN / inf patterns."""

import warnings

import numpy as np
import pytest

from b2l_testutil import golden

pytestmark = pytest.mark.gpu

import pyloo_b200 as pl
from pyloo_b200 import engine
from pyloo_b200.data import LiteDataArray, from_dict
from oracle import is_oracle as iso
from oracle.psis_oracle import loo_pointwise as orc_loo

RTOL = 1e-10


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol, equal_nan=True)


# ------------------------------------------------------------------------------------ SIS / TIS
@pytest.mark.parametrize("tag", ["n4000", "wide", "t15", "odd", "edge", "const"])
@pytest.mark.parametrize("method", ["sis", "tis"])
def test_islw_against_reference_vectors(tag, method):
    g = golden("is_eloo.npz")
    x = g[f"{tag}_x"]
    keep = x.copy()
    with np.errstate(all="ignore"):
        lw, ess = engine.islw_host(x, method)
    assert np.array_equal(x, keep, equal_nan=True)
    close(lw, g[f"{tag}_{method}_lw"], atol=1e-12)
    close(ess, g[f"{tag}_{method}_ess"])


@pytest.mark.parametrize("S", [1, 2, 7, 33, 255, 256, 257, 1000, 4001, 30000])
@pytest.mark.parametrize("method", ["sis", "tis"])
def test_islw_shapes_against_oracle(S, method):
    # odd S and unaligned rows take the plain staging loop; S = 30000 does not fit shared memory and runs
    # straight from global memory
    rng = np.random.default_rng(S)
    x = rng.normal(size=(5, S)) * 3.0
    lw, ess = engine.islw_host(x, method)
    ref_lw, ref_ess = iso.islw(x, method)
    close(lw, ref_lw, atol=1e-12)
    close(ess, ref_ess)


@pytest.mark.parametrize("method", ["sis", "tis"])
def test_islw_many_rows_and_batch_invariance(method):
    rng = np.random.default_rng(5)
    x = rng.standard_t(3, size=(3000, 512))
    lw, ess = engine.islw_host(x, method)
    idx = rng.choice(3000, size=40, replace=False)
    ref_lw, ref_ess = iso.islw(x[idx], method)
    close(lw[idx], ref_lw, atol=1e-12)
    close(ess[idx], ref_ess)
    lw1, ess1 = engine.islw_host(x[idx], method)            # same rows, different batch: bit-identical
    assert np.array_equal(lw1, lw[idx]) and np.array_equal(ess1, ess[idx])


def test_compute_importance_weights_dispatch_and_names():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(6, 400))
    for method, fn in (("sis", pl.sislw), ("tis", pl.tislw)):
        lw, ess = pl.compute_importance_weights(x, method=method)
        ref_lw, ref_ess = iso.islw(x, method)
        close(lw, ref_lw, atol=1e-12)
        close(ess, ref_ess)
        lw2, ess2 = fn(x)
        assert np.array_equal(lw, lw2) and np.array_equal(ess, ess2)
        da = LiteDataArray(x.T.copy(), ("__sample__", "obs"))          # sample axis first
        lw_da, ess_da = pl.compute_importance_weights(da, method=method.upper())
        assert lw_da.name == "log_weights" and ess_da.name == "ess"     # base.py:168-173
        assert lw_da.dims == ("obs", "__sample__") and ess_da.dims == ("obs",)
        close(lw_da.values, ref_lw, atol=1e-12)
    one, ess0 = pl.sislw(x[0])
    assert one.shape == (400,) and ess0.shape == ()
    with pytest.raises(ValueError, match="Invalid method"):
        pl.compute_importance_weights(x, method="nope")


@pytest.mark.parametrize("method", ["sis", "tis"])
def test_loo_is_pointwise_against_reference_vectors(method):
    g = golden("is_eloo.npz")
    res = engine.loo_is_host(g["loo_ll_sn"], method)                     # obs-fastest (S, N) layout
    close(res["elpd_i"], g[f"loo_{method}_elpd_i"])
    close(res["ess_i"], g[f"loo_{method}_ess_i"])
    close(res["lppd_i"], g["loo_lppd_i"])
    assert res["n_nan_in"] == 1
    rows = np.ascontiguousarray(g["loo_ll_sn"].T)                         # row-contiguous (N, S) view
    res2 = engine.loo_is_host(rows.T, method)                             # row kernel instead of the column form
    for key in ("elpd_i", "ess_i", "lppd_i"):
        close(res2[key], res[key], 1e-12)
    close(res2["elpd_i"], g[f"loo_{method}_elpd_i"])


@pytest.mark.parametrize("method", ["sis", "tis"])
def test_loo_method_api(method):
    rng = np.random.default_rng(11)
    ll = -1.0 + 0.7 * rng.normal(size=(4, 250, 9))
    idata = from_dict(posterior={"mu": rng.normal(size=(4, 250))}, log_likelihood={"y": ll},
                      dims={"y": ["obs"]})
    with pytest.warns(UserWarning, match=f"Using {method.upper()} for LOO computation"):
        res = pl.loo(idata, pointwise=True, method=method)
    ref = iso.loo_is_summary(ll.reshape(-1, 9), method)
    for key in ("elpd_loo", "se", "p_loo", "p_loo_se", "looic", "looic_se"):
        close(res[key], ref[key])
    close(res["loo_i"].values, ref["elpd_i"])
    close(res["ess"].values, ref["ess_i"])
    assert "pareto_k" not in res and "good_k" not in res                # loo.py:400-410
    assert bool(res["warning"]) == ref["warning"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dev = pl.loo(idata, method=method, scale="deviance")
    close(dev["elpd_loo"], -2 * ref["elpd_loo"])
    assert list(dev.index[-4:]) == ["scale", "looic", "looic_se", "subsample_size"]


@pytest.mark.parametrize("method", ["sis", "tis"])
@pytest.mark.parametrize("S,N", [(3, 5), (9, 33), (64, 31), (1001, 70), (4000, 130)])
def test_loo_is_column_form_against_oracle(S, N, method):
    """(chain, draw, obs) layout: the column-form kernel (no transposed panels) against the oracle for ragged
    shapes, wide ratios, NaN and +-inf columns (redone serially with the row kernel's arithmetic)."""
    rng = np.random.default_rng(S * 7 + N)
    ll = -1.0 + rng.normal(size=(S, N)) * rng.uniform(0.05, 4.0, size=N)
    ll[0, 0] = -60.0                                  # one dominant ratio
    ll[1 % S, 1] = np.nan
    if N > 4:
        ll[2 % S, 2] = np.inf
        ll[0, 3] = -np.inf
        ll[:, 4] = -np.inf
    res = engine.loo_is_host(ll, method)
    with np.errstate(all="ignore"):
        ref = iso.loo_is_pointwise(ll, method)
    for key in ("elpd_i", "ess_i", "lppd_i"):
        close(res[key], ref[key], 1e-10, atol=1e-13)
    assert res["n_nan_in"] == 1


def test_loo_is_low_ess_warning():
    rng = np.random.default_rng(3)
    ll = 6.0 * rng.normal(size=(2, 300, 4))                              # wild ratios: ESS collapses
    idata = from_dict(posterior={"mu": rng.normal(size=(2, 300))}, log_likelihood={"y": ll})
    with pytest.warns(UserWarning, match="Low effective sample size detected"):
        res = pl.loo(idata, method="sis")
    assert res["warning"]


# ------------------------------------------------------------------------------------ e_loo
def test_eloo_against_reference_vectors():
    g = golden("is_eloo.npz")
    x, lw, lr = g["eloo_x"], g["eloo_lw"], g["eloo_lr"]
    with np.errstate(all="ignore"):
        v, k = engine.eloo_host(x, lw, lr, "mean")
        close(v, g["eloo_mean"], atol=1e-13)
        close(k, g["eloo_k_mean"], 1e-12)
        v, k = engine.eloo_host(x, lw, lr, "variance")
        close(v, g["eloo_var"], 1e-9)
        close(k, g["eloo_k_var"], 1e-12)
        v, _ = engine.eloo_host(x, lw, lr, "sd")
        close(v, np.sqrt(g["eloo_var"]), 1e-9)
        v, k = engine.eloo_host(None, lw, lr, "none")
        assert v is None
        close(k, g["eloo_k_none"], 1e-12)
        for tl, key in ((20, "short_k"), (7, "short_k7")):
            _, k = engine.eloo_host(g["short_x"], g["short_lr"], None, "mean", tl)
            close(k, g[key], 1e-12)


@pytest.mark.parametrize("S,N", [(64, 40), (1001, 17), (4000, 700), (12000, 9)])
def test_eloo_against_oracle(S, N):
    # S = 1001: odd rows, plain staging; S = 12000 with separate ratios: three rows exceed shared memory, the
    # kernel reads global memory and keeps h * r in the workspace
    rng = np.random.default_rng(S + N)
    x = rng.normal(size=(N, S)) + rng.normal(size=(N, 1))
    lr = rng.standard_t(4, size=(N, S))
    lw, _ = iso.islw(lr, "tis")
    x[0] = np.round(x[0])                                                 # heavy ties in h
    lr[1] = np.round(lr[1], 1)                                            # heavy ties in the ratios
    sub = slice(0, min(N, 25))
    for kind in ("mean", "variance", "sd"):
        v, k = engine.eloo_host(x, lw, lr, kind)
        ref = iso.e_loo_arrays(x[sub], lw[sub], lr[sub], kind)
        close(v[sub], ref["value"], 1e-9, atol=1e-13)
        close(k[sub], ref["pareto_k"], 1e-12)
    v, k = engine.eloo_host(x, lw, None, "mean")                          # ratios default to the weights
    ref = iso.e_loo_arrays(x[sub], lw[sub], None, "mean")
    close(v[sub], ref["value"], 1e-9, atol=1e-13)
    close(k[sub], ref["pareto_k"], 1e-12)


def test_eloo_tie_heavy_rows_take_the_exact_extraction():
    # hundreds of ties at the selection threshold overflow the candidate list of the fast path
    rng = np.random.default_rng(21)
    N, S = 12, 4000
    x = rng.normal(size=(N, S))
    lr = rng.normal(size=(N, S))
    lr[0, :600] = 5.0                          # 600 equal maxima: ratio tail all close -> k = inf
    lr[1, :10] = np.linspace(6.0, 7.0, 10)     # 10 distinct values, then 400 ties at the 11th..20th place
    lr[1, 10:410] = 5.5
    x[2, ::2] = 0.0                            # half of h is exactly 0: left tail of x^2 * r is all zeros
    x[3] = np.abs(x[3])
    x[3, :900] = 0.0                           # non-negative h with 900 zeros: left tail of h * r all zeros
    x[4, :300] = 50.0                          # 300 equal large h, ratios tied too: right tail ties
    lr[4, :300] = 1.0
    lw, _ = iso.islw(lr, "sis")
    for kind in ("mean", "variance"):
        v, k = engine.eloo_host(x, lw, lr, kind)
        ref = iso.e_loo_arrays(x, lw, lr, kind)
        close(v, ref["value"], 1e-9, atol=1e-13)
        close(k, ref["pareto_k"], 1e-12)
    assert np.isinf(ref["pareto_k"][0])


def test_eloo_quantiles_against_reference_vectors():
    g = golden("is_eloo.npz")
    with np.errstate(all="ignore"):
        q = engine.eloo_quantile_host(g["eloo_x"], g["eloo_lw"], g["eloo_probs"])
    # rows: 3 constant draws, 5 NaN / 6 inf in the draws, 7 uniform weights (np.quantile rule)
    close(q, g["eloo_quant"], 1e-10, atol=1e-13)


@pytest.mark.parametrize("S", [5, 256, 300, 1000, 4000, 8192, 12000])
def test_eloo_quantiles_against_oracle(S):
    rng = np.random.default_rng(S)
    N = 9
    x = rng.normal(size=(N, S)) * 3 + 1
    lw = rng.standard_t(3, size=(N, S))
    lw[1] = 0.7                                              # uniform weights: np.quantile's linear rule
    lw[2, : S // 2] = -800.0                                 # half of the weights underflow to 0
    probs = [0.001, 0.1, 0.5, 0.9, 0.999]
    q = engine.eloo_quantile_host(x, lw, probs)
    ref = iso.e_loo_arrays(x, lw, None, "quantile", probs=probs)["value"]
    close(q, ref, 1e-9, atol=1e-12)
    assert np.all(np.diff(q, axis=1) >= 0)                    # monotone in the probability


def test_eloo_quantile_api_and_limits():
    rng = np.random.default_rng(4)
    x = LiteDataArray(rng.normal(size=(5, 3, 600)), ("a", "b", "__sample__"), name="y")
    lw = LiteDataArray(rng.normal(size=(5, 3, 600)), ("a", "b", "__sample__"))
    res = pl.e_loo(x, log_weights=lw, type="quantile", probs=[0.25, 0.75])
    assert res.value.dims == ("a", "b", "quantile") and res.value.shape == (5, 3, 2)
    ref = iso.e_loo_arrays(x.values.reshape(15, 600), lw.values.reshape(15, 600), None, "quantile",
                           probs=[0.25, 0.75])
    close(res.value.values.reshape(15, 2), ref["value"], 1e-9, atol=1e-12)
    close(res.pareto_k.values.ravel(), ref["pareto_k"], 1e-12)      # h is None: ratio tail only
    one = pl.e_loo(x, log_weights=lw, type="quantile", probs=0.5)
    assert one.value.shape == (5, 3, 1)
    with pytest.raises(ValueError, match="probs must be between 0 and 1"):
        pl.e_loo(x, log_weights=lw, type="quantile", probs=[0.5, 1.0])
    with pytest.raises(NotImplementedError, match="S <= 16384"):
        engine.eloo_quantile_host(np.zeros((1, 17000)), np.zeros((1, 17000)), [0.5])


def test_eloo_api_and_diagnostics():
    rng = np.random.default_rng(8)
    ll = -1.0 + 0.5 * rng.normal(size=(4, 300, 6))
    yrep = rng.normal(size=(4, 300, 6)) * 2 + 1
    idata = from_dict(posterior={"mu": rng.normal(size=(4, 300))}, log_likelihood={"y": ll},
                      posterior_predictive={"y": yrep}, dims={"y": ["obs"]})
    llda = idata.log_likelihood["y"].stack(__sample__=("chain", "draw"))
    lw, _ = pl.psislw(-llda)
    res = pl.e_loo(idata, var_name="y", log_weights=lw, log_ratios=-llda, type="mean")
    x = yrep.reshape(-1, 6).T
    ref = iso.e_loo_arrays(x, np.asarray(lw.values), -ll.reshape(-1, 6).T, "mean")
    assert res.value.dims == ("obs",) and res.pareto_k.dims == ("obs",)
    close(res.value.values, ref["value"], 1e-9)
    close(res.pareto_k.values, ref["pareto_k"], 1e-12)
    close(res.min_ss.values, ref["min_ss"])
    close(res.khat_threshold.values, ref["khat_threshold"])
    close(res.convergence_rate.values, ref["convergence_rate"])
    sd = pl.e_loo(idata, log_weights=lw, type="sd")                       # ratios default to the weights
    close(sd.value.values, iso.e_loo_arrays(x, np.asarray(lw.values), None, "sd")["value"], 1e-9)
    w = LiteDataArray(np.exp(np.asarray(lw.values)), lw.dims)
    byw = pl.e_loo(idata, weights=w, type="mean")                          # e_loo.py:199-200
    close(byw.value.values, ref["value"], 1e-9)
    with pytest.raises(ValueError, match="type must be"):
        pl.e_loo(idata, log_weights=lw, type="median")
    with pytest.raises(ValueError, match="Either weights or log_weights"):
        pl.e_loo(idata)
    with pytest.raises(ValueError, match="probs must be provided"):
        pl.e_loo(idata, log_weights=lw, type="quantile")
    with pytest.raises(ValueError, match="does not have a nope group"):
        pl.e_loo(idata, group="nope", log_weights=lw)
    k1 = pl.k_hat(x[0], -ll.reshape(-1, 6).T[0])
    close(k1, ref["pareto_k"][0], 1e-12)
    kd = pl.compute_pareto_k(LiteDataArray(x, ("obs", "__sample__")), -llda)
    close(kd.values, ref["pareto_k"], 1e-12)
    with pytest.raises(ValueError, match="tail_len must be at least 5"):
        pl.compute_pareto_k(x[0], x[0], tail_len=3)


# ------------------------------------------------------------------------------------ consumers of e_loo
def _predictive_model(seed=5, chains=4, draws=300, n_obs=12):
    rng = np.random.default_rng(seed)
    ll = -1.0 + 0.6 * rng.normal(size=(chains, draws, n_obs))
    yrep = rng.normal(size=(chains, draws, n_obs)) + np.linspace(-1, 1, n_obs)
    yrep2 = yrep + 0.1 * rng.normal(size=yrep.shape)
    y = rng.normal(size=n_obs) + np.linspace(-1, 1, n_obs)
    idata = from_dict(posterior={"mu": rng.normal(size=(chains, draws))}, log_likelihood={"y": ll},
                      posterior_predictive={"y": yrep, "y2": yrep2}, observed_data={"y": y},
                      dims={"y": ["obs"], "y2": ["obs"]})
    rows = lambda a: a.reshape(-1, n_obs).T  # noqa: E731  (N, S) with stack(chain, draw) order
    return idata, rows(ll), rows(yrep), rows(yrep2), y


@pytest.mark.parametrize("metric", ["mae", "mse", "rmse"])
def test_loo_predictive_metric(metric):
    idata, ll, yrep, _, y = _predictive_model()
    res = pl.loo_predictive_metric(idata, y, var_name="y", metric=metric, r_eff=0.8)
    ref = iso.loo_predictive_metric_arrays(yrep, ll, y, metric, 0.8)
    close(res["estimate"], ref["estimate"], 1e-9)
    close(res["se"], ref["se"], 1e-9)


def test_loo_predictive_metric_binary_and_errors():
    rng = np.random.default_rng(9)
    ll = -0.7 + 0.3 * rng.normal(size=(2, 400, 30))
    prob = np.clip(rng.random(size=(2, 400, 30)), 0.01, 0.99)
    yb = (rng.random(30) < 0.5).astype(float)
    idata = from_dict(posterior={"mu": rng.normal(size=(2, 400))}, log_likelihood={"y": ll},
                      posterior_predictive={"y": prob}, dims={"y": ["obs"]})
    for metric in ("acc", "balanced_acc"):
        res = pl.loo_predictive_metric(idata, yb, metric=metric)
        ref = iso.loo_predictive_metric_arrays(prob.reshape(-1, 30).T, ll.reshape(-1, 30).T, yb, metric)
        close([res["estimate"], res["se"]], [ref["estimate"], ref["se"]], 1e-9)
    with pytest.raises(ValueError, match="Invalid metric"):
        pl.loo_predictive_metric(idata, yb, metric="f1")
    with pytest.raises(ValueError, match=r"Length of y \(3\) must match"):
        pl.loo_predictive_metric(idata, yb[:3])
    with pytest.raises(ValueError, match="does not have a nope group"):
        pl.loo_predictive_metric(idata, yb, group="nope")


@pytest.mark.parametrize("scale", [False, True])
def test_loo_score(scale):
    idata, ll, yrep, yrep2, y = _predictive_model(seed=6)
    # The shuffles come from NumPy's global stream.  Seed 131: neither permutation has a 2-cycle.  A 2-cycle
    # (s <-> s') makes joint[s] == joint[s'] exactly; if that tie falls in the PSIS tail, which of the two draws
    # receives which smoothed value follows np.argsort's unspecified order among equal keys in the reference
    # (draw-index order here), and |x - x2| differs between the two draws.
    np.random.seed(131)
    res = pl.loo_score(idata, x_var="y", x2_var="y2", y_var="y", permutations=2, reff=1.0, scale=scale,
                       pointwise=True)
    np.random.seed(131)
    ref = iso.loo_score_arrays(yrep, yrep2, ll, y, permutations=2, reff=1.0, scale=scale)
    close(res.pointwise, ref["pointwise"], 1e-8)
    close(res.estimates["Estimate"][0], ref["estimate"], 1e-8)
    close(res.estimates["SE"][0], ref["se"], 1e-8)
    close(res.pareto_k.values, ref["pareto_k"])
    assert res.good_k == pytest.approx(min(1 - 1 / np.log10(1200), 0.7)) and res.warning in (True, False)
    plain = pl.loo_score(idata, x_var="y", x2_var="y2", y_var="y", reff=1.0)
    assert plain.pareto_k is None and plain.pointwise.shape == (12,)
    with pytest.raises(ValueError, match="Variable 'zz' not found"):
        pl.loo_score(idata, x_var="zz", y_var="y")


# ------------------------------------------------------------------------------------ loo_group
@pytest.mark.parametrize("method", ["psis", "sis", "tis"])
def test_loo_group(method):
    rng = np.random.default_rng(17)
    n_obs, chains, draws = 60, 4, 250
    ll = -1.0 + 0.25 * rng.normal(size=(chains, draws, n_obs))
    ll[1, 7, 3] = np.nan
    group_ids = rng.permutation(np.repeat(np.array(["a", "b", "c", "d", "e", "f", "g", "h", "i", "j"]), 6))
    idata = from_dict(posterior={"mu": rng.normal(size=(chains, draws))}, log_likelihood={"y": ll},
                      dims={"y": ["obs"]})
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        res = pl.loo_group(idata, group_ids, pointwise=True, reff=0.9, method=method)
    msgs = " | ".join(str(w.message) for w in rec)
    assert "NaN values detected in log-likelihood" in msgs
    assert (f"Using {method.upper()} for LOGO computation" in msgs) == (method != "psis")
    ref = iso.loo_group_arrays(ll.reshape(-1, n_obs), group_ids, method, 0.9)
    for key in ("elpd_logo", "se", "p_logo", "p_logo_se", "logoic", "logoic_se"):
        close(res[key], ref[key])
    close(res["logo_i"].values, ref["logo_i"])
    assert list(res["logo_i"].coords["group"]) == list(ref["groups"])
    close(res["pareto_k" if method == "psis" else "ess"], ref["diagnostic"])
    assert res["n_groups"] == 10 and res["n_samples"] == 1000
    assert ("good_k" in res) == (method == "psis")
    assert "elpd_logo" in str(res) and "10 groups log-likelihood matrix" in str(res)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dev = pl.loo_group(idata, group_ids, reff=0.9, method=method, scale="deviance")
    close(dev["elpd_logo"], -2 * ref["elpd_logo"])
    assert list(dev.index[:7]) == ["elpd_logo", "se", "p_logo", "p_logo_se", "n_samples", "n_groups", "warning"]
    with pytest.raises(ValueError, match="Length of group_ids"):
        pl.loo_group(idata, group_ids[:5], reff=1.0)


def test_loo_group_sums_row_layout_and_many_groups():
    # engine level: contiguous integer groups of unequal size, (N, S)-row input viewed as (S, N)
    rng = np.random.default_rng(23)
    S, N = 2000, 700
    rows = -0.5 + 0.05 * rng.normal(size=(N, S))
    sizes = rng.integers(1, 12, size=200)
    gid = np.repeat(np.arange(200), sizes)[:N]
    gid = np.concatenate([gid, np.full(N - gid.size, 199)]) if gid.size < N else gid
    G = int(gid.max()) + 1
    res = engine.group_loo_host(rows.T, gid, G, 1.0, "psis")
    ref = iso.loo_group_arrays(rows.T, gid, "psis", 1.0)
    close(res["elpd_i"], ref["logo_i"])
    close(res["pareto_k"], ref["diagnostic"])
    close(res["stats"].elpd_sum, ref["elpd_logo"], 1e-11)


def test_host_wrappers_are_slab_invariant(monkeypatch):
    # host arrays larger than one staging slab go through the GPU in pieces: same bits as a single pass
    rng = np.random.default_rng(31)
    N, S = 300, 512
    lr = rng.standard_t(4, size=(N, S))
    x = rng.normal(size=(N, S))
    ll_sn = np.ascontiguousarray(-lr.T)
    gid = rng.integers(0, 25, size=N)

    def run_all():
        lw, ess = engine.islw_host(lr, "tis")
        loo_is = engine.loo_is_host(ll_sn, "sis")
        v, k = engine.eloo_host(x, lw, lr, "sd")
        q = engine.eloo_quantile_host(x, lw, [0.2, 0.8])
        pv, pk, pp = engine.psis_expectation_host(x, lr, 1.0, "mean")
        grp = engine.group_loo_host(ll_sn, gid, 25, 1.0, "psis")
        return [lw, ess, loo_is["elpd_i"], loo_is["ess_i"], loo_is["lppd_i"], v, k, q, pv, pk, pp, grp["elpd_i"],
                grp["pareto_k"]]

    whole = run_all()
    monkeypatch.setattr(engine, "_CHUNK_BYTES", 256 * 1024)   # 12-60 observations (or 100 draws) per slab
    pieces = run_all()
    for a, b in zip(whole, pieces):
        assert np.array_equal(a, b, equal_nan=True)


def test_loo_subset_gathers_on_the_device():
    # loo_subsample's PSIS stage (loo_subsample.py:330, :371-383): same numbers as the full pass at those indices
    import torch

    rng = np.random.default_rng(41)
    S, N = 2000, 500
    ll = -1.2 + 0.8 * rng.normal(size=(S, N))
    idx = np.concatenate([rng.choice(N, size=70, replace=False), [3, 3, N - 1]])   # repeats allowed
    d_ll = torch.from_numpy(ll).cuda()
    full = engine.loo_cuda(d_ll, 1.0)
    sub = engine.loo_subset_cuda(d_ll, idx, 1.0)
    rows = engine.loo_subset_cuda(d_ll.t().contiguous().t(), idx, 1.0)          # row-contiguous storage
    torch.cuda.synchronize()
    for key in ("elpd_i", "pareto_k", "lppd_i", "var_i"):
        a, b = full[key].cpu().numpy()[idx], sub[key].cpu().numpy()
        close(b, a, 1e-13)
        assert np.array_equal(rows[key].cpu().numpy(), b)
    ref = orc_loo(ll[:, idx[:5]], 1.0)
    close(sub["elpd_i"].cpu().numpy()[:5], ref["elpd_i"])
    with pytest.raises(IndexError):
        engine.loo_subset_cuda(d_ll, [0, N], 1.0)

"""GPU tests of the reference-facing Python API (psislw / compute_importance_weights / loo / waic /
loo_compare): same signatures, container behaviour, warnings and ELPDData rows as pyloo, values
against the CPU oracle (tolerance 1e-10 relative, BASELINE.json)."""

import warnings

import numpy as np
import pytest

from b2l_testutil import golden, has_cuda

pytestmark = pytest.mark.gpu

import pyloo_b200 as pl
from pyloo_b200.data import LiteDataArray, from_dict
from oracle import psis_oracle as orc

RTOL = 1e-10


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol, equal_nan=True)


def make_model(seed=10, chains=4, draws=500, obs=8, shift=0.0, width=1.0):
    rng = np.random.default_rng(seed)
    post = {"mu": rng.normal(size=(chains, draws)), "theta": rng.normal(size=(chains, draws, obs))}
    ll = shift + width * rng.normal(size=(chains, draws, obs))
    return from_dict(posterior=post, log_likelihood={"y": ll}, dims={"y": ["obs_dim"], "theta": ["school"]}), ll


# ------------------------------------------------------------------ psislw (test_psis.py:19-125)
def test_psislw_numpy_contract():
    rng = np.random.default_rng(44)
    lw_in = rng.normal(size=(8, 2000))
    keep = lw_in.copy()
    lw, k = pl.psislw(lw_in)
    assert np.array_equal(lw_in, keep)                      # never mutated (psis.py:78)
    assert lw.shape == lw_in.shape and isinstance(k, np.ndarray) and k.shape == (8,)
    close(np.exp(lw).sum(axis=-1), 1.0, 1e-12)
    ref_lw, ref_k = orc.psislw(lw_in)
    close(k, ref_k)
    close(lw, ref_lw, atol=1e-12)


def test_psislw_1d_returns_0d_k_and_reff_variants():
    x = np.random.default_rng(1).normal(size=2000)
    for reff in (0.5, 1.0, 2.0):                            # test_psis.py:49-58
        lw, k = pl.psislw(x, reff=reff)
        assert lw.shape == x.shape and isinstance(k, np.ndarray) and k.shape == ()
        ref_lw, ref_k = orc.psislw(x, reff)
        close(k, ref_k)
        close(lw, ref_lw, atol=1e-12)


def test_psislw_known_answers():
    lw, k = pl.psislw(np.array([1.0, 1.1, 1.2, 1.3]))       # test_psis.py:95-99 -> IndexError? no: S=4
    assert k == np.inf
    close(np.exp(lw).sum(), 1.0, 1e-12)
    lw, k = pl.psislw(np.ones(100))                          # test_psis.py:121-125
    close(lw, -np.log(100.0), 1e-12)
    assert k == np.inf
    lw, k = pl.psislw(np.random.default_rng(2).normal(size=(5, 8)))   # rows of 8 -> all inf (:115-118)
    assert np.all(k == np.inf)


def test_psislw_dataarray_in_dataarray_out_and_multidim():
    rng = np.random.default_rng(3)
    llm = rng.normal(size=(4, 23, 15, 2))                    # multidim_data (test_data.py:15-21)
    da = LiteDataArray(llm, ("chain", "draw", "dim1", "dim2")).stack(__sample__=("chain", "draw"))
    lw, k = pl.psislw(-da, reff=0.7)
    assert isinstance(lw, LiteDataArray) and lw.dims == ("dim1", "dim2", "__sample__") and lw.name == "log_weights"
    assert k.dims == ("dim1", "dim2") and k.name == "pareto_shape"       # psis.py:107-110
    ref_lw, ref_k = orc.psislw(-da.values, 0.7)
    close(k.values, ref_k)
    close(lw.values, ref_lw, atol=1e-12)
    lw2, k2 = pl.compute_importance_weights(-LiteDataArray(llm, ("chain", "draw", "dim1", "dim2")), reff=0.7)
    assert np.array_equal(lw2.values, lw.values) and np.array_equal(k2.values, k.values)   # auto-stack, base.py:93-98
    lw3, k3 = pl.compute_importance_weights(-da.values, method="psis", reff=0.7)
    assert np.array_equal(lw3, lw.values)


def test_psislw_float32_input_computed_in_fp64():
    x = np.random.default_rng(4).normal(size=(3, 1000)).astype(np.float32)
    lw, k = pl.psislw(x)
    assert lw.dtype == np.float32
    ref_lw, ref_k = orc.psislw(x.astype(np.float64))
    close(k, ref_k)


# ------------------------------------------------------------------ loo (test_loo.py)
@pytest.mark.parametrize("scale", ["log", "negative_log", "deviance"])
def test_loo_rows_match_oracle_all_scales(scale):
    g = golden("cfg1_create_model.npz")
    ll = g["ll_sn"].reshape(4, 500, 8)
    idata = from_dict(posterior={"mu": np.zeros((4, 500))}, log_likelihood={"y": ll})
    for reff in (1.0, 0.7):
        res = pl.loo(idata, pointwise=True, reff=reff, scale=scale)
        ref = orc.loo_summary(g["ll_sn"], reff, scale)
        for key in ("elpd_loo", "se", "p_loo", "p_loo_se", "looic", "looic_se", "good_k"):
            close(res[key], ref[key])
        assert res["n_samples"] == 2000 and res["n_data_points"] == 8 and res["scale"] == scale
        assert res["warning"] == ref["warning"] and res["subsample_size"] == 8
        close(res["loo_i"].values, ref["loo_i"])
        close(res["pareto_k"].values, ref["pareto_k"])
        assert list(res.index) == ["elpd_loo", "se", "p_loo", "p_loo_se", "n_samples", "n_data_points", "warning",
                                   "loo_i", "scale", "looic", "looic_se", "pareto_k", "good_k", "subsample_size"]
        nonpw = pl.loo(idata, pointwise=False, reff=reff, scale=scale)
        assert list(nonpw.index) == ["elpd_loo", "se", "p_loo", "p_loo_se", "n_samples", "n_data_points", "warning",
                                     "scale", "looic", "looic_se", "good_k", "subsample_size"]
        assert nonpw["elpd_loo"] == res["elpd_loo"] and "pareto_k" not in nonpw      # test_loo.py:201-216
    if scale == "log":  # SURVEY App. B known answers
        res = pl.loo(idata, reff=1.0)
        close(res["elpd_loo"], -4.09507565863278, 1e-11)
        close(res["se"], 0.0831053977837424, 1e-9)
        close(res["p_loo"], 8.04693296545902, 1e-11)
        assert "elpd_loo" in str(res)


def test_loo_reff_none_single_chain_and_multichain():
    idata, ll = make_model(chains=1, draws=800)
    res = pl.loo(idata)                                        # one chain -> reff = 1 (loo.py:209-210)
    ref = orc.loo_summary(ll.reshape(-1, 8), 1.0)
    close(res["elpd_loo"], ref["elpd_loo"])
    idata4, ll4 = make_model(chains=4)
    res4 = pl.loo(idata4)                                      # own ESS (ArviZ absent): finite, sensible
    assert np.isfinite(res4["elpd_loo"])
    res_fixed = pl.loo(idata4, reff=0.7)
    close(res_fixed["elpd_loo"], orc.loo_summary(ll4.reshape(-1, 8), 0.7)["elpd_loo"])


def test_loo_transposed_and_multidim_models():
    rng = np.random.default_rng(5)
    ll = rng.normal(size=(4, 500, 5, 7))
    idata = from_dict(log_likelihood={"y": ll}, posterior={"mu": np.zeros((4, 500))})
    res = pl.loo(idata, pointwise=True, reff=1.0)
    ref = orc.loo_summary(ll.reshape(2000, 35), 1.0)
    assert res["loo_i"].shape == (5, 7) and res["pareto_k"].shape == (5, 7) and res["n_data_points"] == 35
    close(res["loo_i"].values.ravel(), ref["loo_i"])
    close(res["elpd_loo"], ref["elpd_loo"])
    tr = from_dict(log_likelihood={"y": ll}, posterior={"mu": np.zeros((4, 500))})
    tr.log_likelihood.data_vars["y"] = tr.log_likelihood["y"].transpose("draw", "chain", ...)
    res_t = pl.loo(tr, pointwise=True, reff=1.0)
    close(res_t["elpd_loo"], res["elpd_loo"], 1e-13)


def test_loo_warnings_and_edge_data():
    rng = np.random.default_rng(6)
    ll = -1.0 + rng.normal(size=(4, 250, 6))
    ll[:, :, 1] = ll[0, 0, 1]                                  # constant column (test_loo.py:89-97)
    idata = from_dict(log_likelihood={"y": ll}, posterior={"mu": np.zeros((4, 250))})
    with pytest.warns(UserWarning, match="Estimated shape parameter of Pareto distribution is greater than"):
        res = pl.loo(idata, pointwise=True, reff=1.0)
    assert res["warning"] and np.any(res["pareto_k"].values > res["good_k"])
    const = from_dict(log_likelihood={"y": np.full((2, 300, 4), -1.3)}, posterior={"mu": np.zeros((2, 300))})
    with pytest.warns(UserWarning) as rec:                     # test_loo.py:100-108
        pl.loo(const, pointwise=True, reff=1.0)
    assert any("The point-wise LOO is the same" in str(w.message) for w in rec)
    nan_ll = -1.0 + rng.normal(size=(4, 250, 6))
    nan_ll[1, 7, 2] = np.nan                                   # test_loo.py:139-153
    nan_idata = from_dict(log_likelihood={"y": nan_ll}, posterior={"mu": np.zeros((4, 250))})
    with pytest.warns(UserWarning, match="NaN values detected in log-likelihood"):
        res = pl.loo(nan_idata, reff=1.0)
    assert np.isfinite(res["elpd_loo"])
    with np.errstate(all="ignore"):
        close(res["elpd_loo"], orc.loo_summary(nan_ll.reshape(1000, 6), 1.0)["elpd_loo"])
    big = -1.0 + rng.normal(size=(4, 250, 6))
    big[0, 0, 0] = 1e10; big[1, 1, 1] = -1e10                  # test_loo.py:156-171
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = pl.loo(from_dict(log_likelihood={"y": big}, posterior={"mu": np.zeros((4, 250))}), reff=1.0)
    assert np.isfinite(res["elpd_loo"])


def test_loo_jacobian_semantics():
    idata, ll = make_model()
    base = pl.loo(idata, pointwise=True, reff=1.0)
    adj = np.linspace(-0.5, 0.5, 8)
    res = pl.loo(idata, pointwise=True, reff=1.0, jacobian=adj)   # test_loo.py:307-336
    close(res["loo_i"].values, base["loo_i"].values + adj, 1e-14)
    close(res["elpd_loo"], (base["loo_i"].values + adj).sum(), 1e-13)
    with pytest.raises(ValueError, match="does not match"):
        pl.loo(idata, pointwise=True, reff=1.0, jacobian=np.zeros(3))


# ------------------------------------------------------------------ waic (test_waic.py)
@pytest.mark.parametrize("scale", ["log", "negative_log", "deviance"])
def test_waic_rows_match_oracle(scale):
    g = golden("cfg1_create_model.npz")
    idata = from_dict(log_likelihood={"y": g["ll_sn"].reshape(4, 500, 8)})
    with pytest.warns(UserWarning, match="posterior variance of the log predictive"):
        res = pl.waic(idata, pointwise=True, scale=scale)
    ref = orc.waic_summary(g["ll_sn"], scale)
    for key in ("elpd_waic", "se", "p_waic"):
        close(res[key], ref[key])
    close(res["waic_i"].values, ref["waic_i"])
    assert res["warning"] == ref["warning"] and res["n_samples"] == 2000 and res["n_data_points"] == 8
    assert list(res.index) == ["elpd_waic", "se", "p_waic", "n_samples", "n_data_points", "warning", "waic_i", "scale"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert "waic_i" not in pl.waic(idata, pointwise=False, scale=scale)


def test_waic_nan_inf_policy():
    rng = np.random.default_rng(8)
    ll = -1.0 + 0.3 * rng.normal(size=(2, 400, 5))
    ll[0, 3, 1] = np.nan; ll[1, 5, 2] = np.inf; ll[0, 9, 3] = -np.inf
    idata = from_dict(log_likelihood={"y": ll})
    with pytest.warns(UserWarning) as rec:
        res = pl.waic(idata, pointwise=True)
    msgs = " | ".join(str(w.message) for w in rec)
    assert "NaN values detected" in msgs and "Infinite values detected" in msgs   # test_waic.py:46-85
    with np.errstate(all="ignore"):
        ref = orc.waic_summary(ll.reshape(800, 5))
    close(res["waic_i"].values, ref["waic_i"])
    close(res["elpd_waic"], ref["elpd_waic"])


# ------------------------------------------------------------------ loo_compare (test_compare.py)
def _four_models(n_obs=300, draws=500):
    rng = np.random.default_rng(9)
    y = rng.normal(size=n_obs)
    models = {}
    for i in range(4):
        mu = 0.15 * i + 0.05 * rng.normal(size=(4, draws, 1))
        sd = 1.0 + 0.1 * i
        ll = -0.5 * np.log(2 * np.pi * sd**2) - 0.5 * ((y[None, None, :] - mu) / sd) ** 2
        models[f"m{i}"] = from_dict(log_likelihood={"y": ll}, posterior={"mu": mu[..., 0]})
    return models


def _oracle_compare(models, ic, scale):
    out = {}
    for name, idata in models.items():
        ll = idata.log_likelihood["y"].values
        ll_sn = ll.reshape(-1, ll.shape[-1])
        out[name] = orc.loo_summary(ll_sn, 1.0, scale) if ic == "loo" else orc.waic_summary(ll_sn, scale)
    return out


@pytest.mark.parametrize("ic", ["loo", "waic"])
@pytest.mark.parametrize("method", ["stacking", "pseudo-bma", "bb-pseudo-bma"])
def test_loo_compare_against_oracle_pointwise(ic, method):
    from pyloo_b200 import compare as cmp

    models = _four_models()
    pre = {n: (pl.loo(m, pointwise=True, reff=1.0) if ic == "loo" else pl.waic(m, pointwise=True))
           for n, m in models.items()}
    df = pl.loo_compare(pre, ic=ic, method=method, seed=3)
    ref = _oracle_compare(models, ic, "log")
    order = sorted(ref, key=lambda n: -ref[n][f"elpd_{ic}"])
    assert list(df.index) == order and list(df["rank"]) == [0, 1, 2, 3]
    for n in order:
        close(df.loc[n, f"elpd_{ic}"], ref[n][f"elpd_{ic}"])
        close(df.loc[n, f"p_{ic}"], ref[n][f"p_{ic}"], 1e-9)
    best = order[0]
    for n in order[1:]:
        d = ref[n][f"{ic}_i"] - ref[best][f"{ic}_i"]
        close(df.loc[n, "dse"], np.sqrt(len(d) * np.var(d)), 1e-9)
        close(df.loc[n, "elpd_diff"], ref[n][f"elpd_{ic}"] - ref[best][f"elpd_{ic}"], 1e-8)
    assert df.loc[best, "elpd_diff"] == 0 and df.loc[best, "dse"] == 0
    close(df["weight"].sum(), 1.0, 1e-9)
    # weights from the oracle's pointwise values through the same host optimiser
    class _Fake(dict):
        pass
    fake = {n: {f"{ic}_i": LiteDataArray(ref[n][f"{ic}_i"], ("obs",)), f"elpd_{ic}": ref[n][f"elpd_{ic}"]}
            for n in models}
    if method == "stacking":
        w_ref = cmp._stacking_weights(fake, ic, "log")
        for n in models:
            assert abs(df.loc[n, "weight"] - w_ref[n]) < 1e-6
    elif method == "pseudo-bma":
        w_ref = cmp._pseudo_bma_weights(fake, ic, "log")
        for n in models:
            close(df.loc[n, "weight"], w_ref[n], 1e-8)


def test_loo_compare_runs_models_and_scales():
    models = _four_models(n_obs=120)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df_log = pl.loo_compare(models, ic="loo", scale="log")
        df_dev = pl.loo_compare(models, ic="loo", scale="deviance")
    assert list(df_log.index) == list(df_dev.index)          # deviance ranks ascending (compare.py:202)
    close(df_dev["elpd_loo"].values, -2 * df_log["elpd_loo"].values, 1e-12)
    assert (df_log["scale"] == "log").all() and df_log.columns.tolist() == [
        "rank", "elpd_loo", "p_loo", "elpd_diff", "weight", "se", "dse", "warning", "scale"]
    nonpw = {n: pl.loo(m, pointwise=False, reff=1.0) for n, m in models.items()}
    with pytest.raises(ValueError, match="pointwise=True"):
        pl.loo_compare(nonpw)

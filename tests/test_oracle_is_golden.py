"""oracle/is_oracle.py (SIS / TIS / e_loo restatement) against the golden vectors produced by the REAL
reference code (oracle/gen_golden_is.py).  CPU-only.  Tolerance 1e-13 relative (bit-exact on the
generating NumPy build)."""

import numpy as np
import pytest

from oracle import is_oracle as iso
from oracle import _refload
from b2l_testutil import golden

RTOL = 1e-13


def _close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


@pytest.mark.parametrize("tag", ["n4000", "wide", "t15", "odd", "edge", "const"])
@pytest.mark.parametrize("method", ["sis", "tis"])
def test_islw_rows(tag, method):
    g = golden("is_eloo.npz")
    lw, ess = iso.islw(g[f"{tag}_x"], method)
    _close(lw, g[f"{tag}_{method}_lw"])
    _close(ess, g[f"{tag}_{method}_ess"])


def test_islw_known_properties():
    # test_sis.py / test_tis.py of the reference: weights normalised, ess in (0, S], constant rows -> ess = S
    g = golden("is_eloo.npz")
    for method in ("sis", "tis"):
        lw, ess = iso.islw(g["n4000_x"], method)
        _close(np.exp(lw).sum(axis=1), np.ones(lw.shape[0]), 1e-12)
        assert np.all((ess > 0) & (ess <= 4000))
        lw, ess = iso.islw(np.ones((2, 100)), method)
        _close(ess, [100.0, 100.0], 1e-12)
        _close(lw, np.full((2, 100), -np.log(100)), 1e-12)
    # truncation can only raise the effective sample size
    assert np.all(iso.islw(g["t15_x"], "tis")[1] >= iso.islw(g["t15_x"], "sis")[1])


@pytest.mark.parametrize("method", ["sis", "tis"])
def test_loo_is_pointwise(method):
    g = golden("is_eloo.npz")
    pw = iso.loo_is_pointwise(g["loo_ll_sn"], method)
    _close(pw["elpd_i"], g[f"loo_{method}_elpd_i"])
    _close(pw["ess_i"], g[f"loo_{method}_ess_i"])
    _close(pw["lppd_i"], g["loo_lppd_i"])
    assert pw["n_nan_in"] == 1


def test_eloo_pieces():
    g = golden("is_eloo.npz")
    x, lw, lr = g["eloo_x"], g["eloo_lw"], g["eloo_lr"]
    r = iso.e_loo_arrays(x, lw, lr, "mean")
    _close(r["value"], g["eloo_mean"])
    _close(r["pareto_k"], g["eloo_k_mean"])
    r = iso.e_loo_arrays(x, lw, lr, "variance")
    _close(r["value"], g["eloo_var"])
    _close(r["pareto_k"], g["eloo_k_var"])
    _close(iso.e_loo_arrays(x, lw, lr, "sd")["value"], np.sqrt(g["eloo_var"]))
    r = iso.e_loo_arrays(x, lw, lr, "quantile", probs=g["eloo_probs"])
    _close(r["value"], g["eloo_quant"])
    _close(r["pareto_k"], g["eloo_k_none"])
    _close([iso.k_hat(a, b) for a, b in zip(g["short_x"], g["short_lr"])], g["short_k"])
    _close([iso.k_hat(a, b, 7) for a, b in zip(g["short_x"], g["short_lr"])], g["short_k7"])


def test_khat_collapses_to_prior_mean_like_the_reference():
    # the reference's k_hat feeds _gpdfit a tail whose last element is 0 (see oracle/is_oracle.py): the
    # regular case is 5 / (n + 10); all-close ratio tails give +inf; NaN ratios give NaN
    g = golden("is_eloo.npz")
    k = g["eloo_k_mean"]
    assert k[0] == pytest.approx(5 / 30, rel=1e-15)
    assert np.isinf(k[7]) and np.isnan(k[8])
    assert g["short_k"][0] == pytest.approx(5 / 22, rel=1e-15)
    assert g["short_k7"][0] == pytest.approx(5 / 17, rel=1e-15)


def test_pareto_diagnostics():
    g = golden("is_eloo.npz")
    _close([iso.pareto_min_ss(k) for k in g["diag_k"]], g["diag_min_ss"])
    _close([iso.pareto_convergence_rate(k, 4000) for k in g["diag_k"]], g["diag_rate"])
    _close(iso.pareto_khat_threshold(4000), g["diag_thr"][0])


def test_metric_and_score_formulas():
    g = golden("is_eloo.npz")
    for name in ("mae", "mse", "rmse"):
        r = iso.predictive_metric(g["metric_y"], g["metric_yhat"], name)
        _close([r["estimate"], r["se"]], g[f"metric_{name}"])
    for name in ("acc", "balanced_acc"):
        r = iso.predictive_metric(g["metric_yb"], g["metric_phat"], name)
        _close([r["estimate"], r["se"]], g[f"metric_{name}"])
    _close(iso.crps(g["crps_exx"], g["crps_exy"]), g["crps_plain"])
    _close(iso.crps(g["crps_exx"], g["crps_exy"], scale=True), g["crps_scaled"])


def test_product_metric_formulas_match_reference_vectors():
    # host-side formulas of the product (no GPU needed): pyloo_b200.loo_predictive_metric / loo_score
    import importlib

    lpm = importlib.import_module("pyloo_b200.loo_predictive_metric")  # the package re-exports the functions
    lsc = importlib.import_module("pyloo_b200.loo_score")

    g = golden("is_eloo.npz")
    for name, fn, a, b in (("mae", lpm._mae, "metric_y", "metric_yhat"), ("mse", lpm._mse, "metric_y", "metric_yhat"),
                           ("rmse", lpm._rmse, "metric_y", "metric_yhat"),
                           ("acc", lpm._accuracy, "metric_yb", "metric_phat"),
                           ("balanced_acc", lpm._balanced_accuracy, "metric_yb", "metric_phat")):
        r = fn(g[a], g[b])
        _close([r["estimate"], r["se"]], g[f"metric_{name}"])
    _close(lsc._crps(g["crps_exx"], g["crps_exy"]), g["crps_plain"])
    _close(lsc._crps(g["crps_exx"], g["crps_exy"], scale=True), g["crps_scaled"])
    with pytest.raises(ValueError, match="y must contain values between 0 and 1"):
        lpm._accuracy(np.array([0.0, 2.0]), np.array([0.1, 0.9]))


@pytest.mark.skipif(not _refload.reference_available(), reason="reference tree not present")
def test_restatement_against_live_reference():
    mods = _refload.load_reference_modules(("utils", "psis", "sis", "tis", "e_loo"))
    rng = np.random.default_rng(7)
    x = rng.normal(size=(6, 300)) * 3
    for row in x:
        a, ea = mods["sis"]._sislw(row.copy())
        b, eb = iso.sislw_row(row)
        assert np.array_equal(a, b) and ea == eb
        a, ea = mods["tis"]._tislw(row.copy(), 300)
        b, eb = iso.tislw_row(row, 300)
        assert np.array_equal(a, b) and ea == eb
    h = rng.normal(size=(6, 300))
    with np.errstate(all="ignore"):
        for hr, lr in zip(h, x):
            assert mods["e_loo"].k_hat(hr, lr) == iso.k_hat(hr, lr)
            w = np.exp(lr - mods["utils"]._logsumexp(lr))
            assert mods["e_loo"]._wvar_func(hr, w) == iso.weighted_variance(hr, lr)
            assert mods["e_loo"]._weighted_quantile(hr, w, 0.3) == iso.weighted_quantile(hr, lr, 0.3)

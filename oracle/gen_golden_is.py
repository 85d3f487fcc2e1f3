"""Generate tests/golden/is_eloo.npz from the REAL reference (build container only).

    python oracle/gen_golden_is.py

Runs the reference's own ``_sislw`` (pyloo/sis.py:86-106), ``_tislw`` (pyloo/tis.py:91-120), ``_logsumexp``
(pyloo/utils.py:305-359), ``k_hat`` / ``_wvar_func`` / ``_weighted_quantile`` and the Pareto diagnostics
(pyloo/e_loo.py:328-426, :518-554), loaded from ``/root/reference`` with ``oracle/_refload.py``, on seeded
inputs and on the degenerate cases the reference's tests exercise (pyloo/tests/base_tests/test_e_loo.py,
test_sis.py, test_tis.py).  The xarray drivers around them cannot run here (xarray is absent): the per-row
glue (normalise, multiply, sum over the sample axis; pyloo/e_loo.py:429-436, :557-559) is restated inline
and marked.  TEST INFRASTRUCTURE.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import _refload  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    mods = _refload.load_reference_modules(("utils", "psis", "sis", "tis", "e_loo", "rcparams",
                                            "loo_predictive_metric", "loo_score"))
    utils, sis, tis, eloo = mods["utils"], mods["sis"], mods["tis"], mods["e_loo"]
    rng = np.random.default_rng(20261019)
    out = {}

    def is_case(tag, x):
        S = x.shape[-1]
        with np.errstate(all="ignore"):
            s_rows = [sis._sislw(r.copy()) for r in x]
            t_rows = [tis._tislw(r.copy(), S) for r in x]
        out[f"{tag}_x"] = x
        out[f"{tag}_sis_lw"] = np.array([r[0] for r in s_rows])
        out[f"{tag}_sis_ess"] = np.array([r[1] for r in s_rows])
        out[f"{tag}_tis_lw"] = np.array([r[0] for r in t_rows])
        out[f"{tag}_tis_ess"] = np.array([r[1] for r in t_rows])

    is_case("n4000", rng.normal(size=(12, 4000)))
    is_case("wide", 5.0 * rng.normal(size=(8, 1000)))
    is_case("t15", rng.standard_t(1.5, size=(8, 2000)))
    is_case("odd", rng.normal(size=(5, 33)))
    edge = rng.normal(size=(5, 64))
    edge[1, 3] = np.nan
    edge[2, 5] = np.inf
    edge[3, :] = -np.inf
    edge[4, 7] = -np.inf
    is_case("edge", edge)
    is_case("const", np.ones((2, 100)))

    # loo with SIS / TIS weights (pyloo/loo.py:286-289, :319-337) on a sample-major matrix
    ll = -1.4 + rng.normal(size=(2000, 10))
    ll[5, 2] = np.nan
    llw = np.where(np.isnan(ll), -1e10, ll).T.copy()
    lse = utils.make_ufunc(utils._logsumexp, n_dims=1, ravel=False)
    out["loo_ll_sn"] = ll
    for name, fn in (("sis", lambda r: sis._sislw(r)), ("tis", lambda r: tis._tislw(r, ll.shape[0]))):
        rows = [fn((-r).copy()) for r in llw]
        lw = np.array([r[0] for r in rows]) + llw
        out[f"loo_{name}_elpd_i"] = lse(lw)
        out[f"loo_{name}_ess_i"] = np.array([r[1] for r in rows])
    out["loo_lppd_i"] = lse(llw, b_inv=ll.shape[0])

    # e_loo pieces
    S = 4000
    x = rng.normal(size=(10, S)) * 2.0 + 0.5
    lr = rng.normal(size=(10, S)) * 1.5
    lw = lr - np.array([utils._logsumexp(r) for r in lr])[:, None] + 0.3   # any normalisation
    x[3] = 2.5                              # constant h
    x[4] = np.where(rng.random(S) < 0.5, 0.0, 1.0)  # two unique values
    x[5, 10] = np.nan
    x[6, 11] = np.inf
    lr[7] = 0.0                             # constant ratios: r tail all close
    lw[7] = -np.log(S)
    lr[8, 100] = np.nan
    probs = np.array([0.05, 0.5, 0.9])
    mean, var, kh_mean, kh_var, kh_none, quant = [], [], [], [], [], []
    with np.errstate(all="ignore"):
        for i in range(x.shape[0]):
            w = np.exp(lw[i] - utils._logsumexp(lw[i]))            # glue: e_loo.py:431-434, :557-559
            mean.append(np.sum(w * x[i]))                          # glue: e_loo.py:436
            var.append(eloo._wvar_func(x[i], w))
            quant.append([eloo._weighted_quantile(x[i], w, p) for p in probs])
            kh_mean.append(eloo.k_hat(x[i], lr[i]))
            kh_var.append(eloo.k_hat(x[i] ** 2, lr[i]))
            kh_none.append(eloo.k_hat(None, lr[i]))
    out.update(eloo_x=x, eloo_lw=lw, eloo_lr=lr, eloo_probs=probs, eloo_mean=np.array(mean),
               eloo_var=np.array(var), eloo_quant=np.array(quant), eloo_k_mean=np.array(kh_mean),
               eloo_k_var=np.array(kh_var), eloo_k_none=np.array(kh_none))
    # short rows: fewer draws than the tail length
    xs = rng.normal(size=(4, 12))
    lrs = rng.normal(size=(4, 12))
    with np.errstate(all="ignore"):
        out.update(short_x=xs, short_lr=lrs, short_k=np.array([eloo.k_hat(a, b) for a, b in zip(xs, lrs)]),
                   short_k7=np.array([eloo.k_hat(a, b, 7) for a, b in zip(xs, lrs)]))
    ks = np.array([-0.3, 0.0, 0.1666, 0.5, 0.7, 1.0, 1.5, np.inf])
    out["diag_k"] = ks
    out["diag_min_ss"] = np.array([eloo._pareto_min_ss(k) for k in ks])
    out["diag_rate"] = np.array([eloo._pareto_convergence_rate(k, S) for k in ks])
    out["diag_thr"] = np.array([eloo._pareto_khat_threshold(S)])

    # metric and score formulas of the e_loo consumers (pyloo/loo_predictive_metric.py:234-356,
    # pyloo/loo_score.py:326-346)
    lpm, lsc = mods["loo_predictive_metric"], mods["loo_score"]
    yv = rng.normal(size=40)
    yh = yv + 0.3 * rng.normal(size=40)
    yb = (rng.random(40) < 0.4).astype(float)
    ph = np.clip(0.5 + 0.4 * (yb - 0.5) + 0.3 * rng.normal(size=40), 0, 1)
    out.update(metric_y=yv, metric_yhat=yh, metric_yb=yb, metric_phat=ph)
    for name, fn, a, b in (("mae", lpm._mae, yv, yh), ("mse", lpm._mse, yv, yh), ("rmse", lpm._rmse, yv, yh),
                           ("acc", lpm._accuracy, yb, ph), ("balanced_acc", lpm._balanced_accuracy, yb, ph)):
        r = fn(a, b)
        out[f"metric_{name}"] = np.array([r["estimate"], r["se"]])
    exx, exy = rng.random(12) + 0.5, rng.random(12) + 0.2
    out.update(crps_exx=exx, crps_exy=exy, crps_plain=lsc._crps(exx, exy, scale=False),
               crps_scaled=lsc._crps(exx, exy, scale=True))

    np.savez_compressed(os.path.join(GOLDEN, "is_eloo.npz"), **out)
    print("written", os.path.join(GOLDEN, "is_eloo.npz"),
          os.path.getsize(os.path.join(GOLDEN, "is_eloo.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()

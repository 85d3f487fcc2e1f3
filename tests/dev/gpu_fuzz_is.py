"""Randomised parity sweep of the widened kernels (SIS / TIS, e_loo, quantiles, column WAIC, group sums)
against the CPU oracle: random shapes, scales, ties, NaN / inf patterns.  Dev checker, not collected by pytest.

    python tests/dev/gpu_fuzz_is.py [n_cases] [seed]
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import is_oracle as iso  # noqa: E402
from oracle import psis_oracle as orc  # noqa: E402
from pyloo_b200 import engine  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
warnings.simplefilter("ignore")
fails = 0


def check(name, got, ref, rtol, atol=0.0):
    global fails
    try:
        np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol, equal_nan=True)
    except AssertionError as err:
        fails += 1
        print("FAIL", name, str(err).splitlines()[3:8])


for case in range(n_cases):
    S = int(rng.choice([5, 17, 64, 255, 256, 257, 1000, 1023, 2048, 4000, 4097, 6000]))
    N = int(rng.integers(1, 40))
    scale = float(rng.choice([0.01, 1.0, 5.0, 50.0]))
    kind = rng.choice(["normal", "t", "ties", "shift"])
    lr = rng.normal(size=(N, S)) if kind != "t" else rng.standard_t(2, size=(N, S))
    lr *= scale
    if kind == "ties":
        lr = np.round(lr, 1)
    if kind == "shift":
        lr += rng.normal(size=(N, 1)) * 1e3
    x = rng.normal(size=(N, S)) * rng.choice([1.0, 1e-3, 1e4]) + rng.normal()
    if rng.random() < 0.3:
        x = np.round(x)
    if rng.random() < 0.2 and N > 2:
        lr[0, rng.integers(S)] = np.nan
        lr[1, rng.integers(S)] = -np.inf
        x[2 % N, rng.integers(S)] = np.nan
    tag = f"case {case} S={S} N={N} {kind} scale={scale}"
    with np.errstate(all="ignore"):
        for method in ("sis", "tis"):
            lw, ess = engine.islw_host(lr, method)
            rlw, ress = iso.islw(lr, method)
            check(f"{tag} {method} lw", lw, rlw, 1e-10, 1e-11)
            check(f"{tag} {method} ess", ess, ress, 1e-10)
        lw_ok = np.where(np.isfinite(rlw), rlw, 0.0)
        for k in ("mean", "variance"):
            v, kh = engine.eloo_host(x, lw_ok, lr, k)
            ref = iso.e_loo_arrays(x, lw_ok, lr, k)
            spread = np.nanmax(np.abs(x), axis=1) ** (2 if k == "variance" else 1)
            # nearly degenerate weights (1 - sum w^2 small): the reference's (E[x^2] - E[x]^2) / (1 - sum w^2) is
            # itself rounding noise at the 1e-5 level there, so only well-conditioned rows are compared tightly
            w = np.exp(lw_ok - np.log(np.exp(lw_ok - lw_ok.max(axis=1, keepdims=True)).sum(axis=1, keepdims=True))
                       - lw_ok.max(axis=1, keepdims=True))
            sane = (1.0 - (w**2).sum(axis=1) > 1e-3) if k == "variance" else np.ones(N, dtype=bool)
            check(f"{tag} eloo {k} value", (v / spread)[sane], (ref["value"] / spread)[sane], 1e-8, 1e-12)
            check(f"{tag} eloo {k} value (all rows, loose)", v / spread, ref["value"] / spread, 1e-3, 1e-9)
            check(f"{tag} eloo {k} khat", kh, ref["pareto_k"], 1e-12)
        if S <= 4100:
            probs = [0.03, 0.5, 0.97]
            q = engine.eloo_quantile_host(x, lw_ok, probs)
            rq = iso.e_loo_arrays(x, lw_ok, None, "quantile", probs=probs)["value"]
            if kind != "ties" and not np.any(x == np.round(x)):   # ties in x: argsort order is unspecified
                check(f"{tag} quantile", q, rq, 1e-8, 1e-9 * float(np.nanmax(np.abs(x))))
        ll_sn = np.ascontiguousarray(-lr.T)
        for method in ("sis", "tis"):   # column-form kernel on the (S, N) layout, row kernel on its transpose
            got = engine.loo_is_host(ll_sn, method)
            rows = engine.loo_is_host(np.ascontiguousarray(ll_sn.T).T, method)
            ref = iso.loo_is_pointwise(ll_sn, method)
            for key in ("elpd_i", "ess_i", "lppd_i"):
                check(f"{tag} loo_{method} cols {key}", got[key], ref[key], 1e-10, 1e-12)
                check(f"{tag} loo_{method} rows {key}", rows[key], ref[key], 1e-10, 1e-12)
        if S >= 8:
            res = engine.loo_cuda(torch.from_numpy(ll_sn).cuda(), 1.0, waic_only=True)
            ww = orc.waic_pointwise(ll_sn)
            check(f"{tag} waic lppd", res["lppdw_i"].cpu().numpy(), ww["lppd_i"], 1e-10, 1e-13)
            check(f"{tag} waic var", res["var_i"].cpu().numpy(), ww["var_i"], 1e-9, 1e-300)
print(f"{n_cases} cases, {fails} failures")
sys.exit(1 if fails else 0)

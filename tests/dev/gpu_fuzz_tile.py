"""Dev tool: the tile path (cluster kernel on the (S, N) layout) against the general row kernel (B2L_FORCE_LEGACY=1)
on random shapes and awkward distributions -- ties from rounding, chain offsets, heavy tails, constant stretches."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyloo_b200 import engine
from oracle import psis_oracle as orc   # the checker (dev tool, not a product path)

rng = np.random.default_rng(int(os.environ.get("SEED", 1)))
n_cases = int(os.environ.get("CASES", 24))
worst = {"elpd": 0.0, "k": 0.0, "lppd": 0.0, "var": 0.0}
fails = []
for case in range(n_cases):
    S = 2 * int(rng.integers(256, int(os.environ.get("SMAX", 4096)) // 2 + 1))
    if S > 4096:
        S = S // 12 * 12   # divisible by 2, 3, 4: takes the chunked units
    N = 2 * int(rng.integers(4, 1500))
    reff = float(rng.choice([float(v) for v in os.environ.get("REFFS", "1.0,0.9,0.8,0.72").split(",")]))
    kind = case % 6
    z = rng.normal(size=(S, N))
    if kind == 1:
        z = rng.standard_t(3, size=(S, N))
    elif kind == 2:
        z = np.round(z, 2)                      # many exact ties, also at the cutoff
    elif kind == 3:
        z[: S // 2] += 0.8                      # two "chains" with different locations
        z[S // 2:] *= 1.7
    elif kind == 4:
        z = rng.standard_t(5, size=(S, N)) * rng.uniform(0.2, 4.0, size=(1, N))
    elif kind == 5:
        z = np.round(rng.standard_t(4, size=(S, N)), 1)
        z[:, ::7] = np.floor(z[:, ::7])         # columns with very few distinct values
    ll = torch.from_numpy(-1.4 + z).cuda()
    engine.handover_reasons()
    a = engine.loo_cuda(ll, reff, want_diag=True)
    torch.cuda.synchronize()
    ho = engine.handover_reasons()
    os.environ["B2L_FORCE_LEGACY"] = "1"
    b = engine.loo_cuda(ll, reff, want_diag=True)
    torch.cuda.synchronize()
    del os.environ["B2L_FORCE_LEGACY"]
    def rel(x, y, floor=1e-13):
        x, y = x.cpu().numpy(), y.cpu().numpy()
        ok = np.isfinite(x) & np.isfinite(y)
        if not np.array_equal(np.isfinite(x), np.isfinite(y)) or not np.array_equal(x[~ok], y[~ok], equal_nan=True):
            return np.inf
        return float(np.max(np.abs(x[ok] - y[ok]) / np.maximum(np.abs(y[ok]), floor) * (np.abs(x[ok] - y[ok]) > floor))) if ok.any() else 0.0
    errs = {"elpd": rel(a["elpd_i"], b["elpd_i"]), "k": rel(a["pareto_k"], b["pareto_k"]),
            "lppd": rel(a["lppd_i"], b["lppd_i"]), "var": rel(a["var_i"], b["var_i"])}
    if os.environ.get("ORACLE", "1") == "1":
        with np.errstate(all="ignore"):
            pw = orc.loo_pointwise(ll.cpu().numpy(), reff)
        for name, res in (("tile", a), ("general", b)):
            for key in ("elpd_i", "pareto_k", "lppd_i"):
                x, y = res[key].cpu().numpy(), pw[key]
                ok = np.isfinite(y)
                e = float(np.max(np.abs(x[ok] - y[ok]) / np.maximum(np.abs(y[ok]), 1e-3))) if ok.any() else 0.0
                errs[f"{name}/{key}"] = e
    cut_eq = bool(torch.equal(a["diag"][:, 1], b["diag"][:, 1]) and torch.equal(a["diag"][:, 2], b["diag"][:, 2]))
    for k_ in errs:
        worst[k_] = max(worst.get(k_, 0.0), errs[k_])
    bad = (not cut_eq) or any(v > 1e-10 for v in errs.values())
    if bad:
        fails.append({"case": case, "S": S, "N": N, "reff": reff, "kind": kind, "errs": errs, "cut_eq": cut_eq})
    print(case, S, N, reff, kind, {k_: f"{v:.1e}" for k_, v in errs.items()}, "cutoff/tail equal" if cut_eq else "CUTOFF DIFFERS",
          "handover", ho, flush=True)
print(json.dumps({"cases": n_cases, "worst": worst, "fails": fails}))

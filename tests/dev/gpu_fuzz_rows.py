"""Dev tool: psislw on row-contiguous (N, S) matrices and loo on the row layout against the oracle, on random
shapes (S from 40 to 16384) and awkward distributions.  The oracle is the checker only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyloo_b200 import engine
from oracle import psis_oracle as orc

rng = np.random.default_rng(int(os.environ.get("SEED", 1)))
n_cases = int(os.environ.get("CASES", 30))
worst, fails = {}, []
for case in range(n_cases):
    S = int(np.exp(rng.uniform(np.log(40), np.log(16384))))
    N = int(rng.integers(3, 400))
    reff = float(rng.choice([1.0, 0.9, 0.6, 0.3]))
    kind = case % 6
    z = rng.normal(size=(N, S))
    if kind == 1:
        z = rng.standard_t(3, size=(N, S))
    elif kind == 2:
        z = np.round(z, 2)
    elif kind == 3:
        z[:, : S // 2] += 0.8
        z[:, S // 2:] *= 1.7
    elif kind == 4:
        z = rng.standard_t(5, size=(N, S)) * rng.uniform(0.2, 4.0, size=(N, 1))
    elif kind == 5:
        z = np.round(rng.standard_t(4, size=(N, S)), 1)
        z[::7] = np.floor(z[::7])
    lw = 1.4 - z
    with np.errstate(all="ignore"):
        ref_lw, ref_k = orc.psislw(lw, reff)
        pw = orc.loo_pointwise(np.ascontiguousarray(-lw.T), reff)
    engine.handover_reasons()
    out, k = engine.psislw_cuda(torch.from_numpy(lw).cuda(), reff)
    res = engine.loo_cuda(torch.from_numpy(-lw).cuda().T, reff)      # (S, N) view of row-contiguous data
    torch.cuda.synchronize()
    ho = engine.handover_reasons()
    def err(x, y, atol=1e-13):
        x = x.cpu().numpy()
        if not (np.array_equal(np.isnan(x), np.isnan(y)) and np.array_equal(np.isinf(x), np.isinf(y))):
            return np.inf
        ok = np.isfinite(y)
        d = np.abs(x[ok] - y[ok])
        return float(np.max(np.where(d > atol, d / np.maximum(np.abs(y[ok]), 1e-300), 0.0))) if ok.any() else 0.0
    if kind in (2, 5):
        # exact ties: which of several equal tail draws receives which smoothed value follows np.argsort's unstable
        # order in the reference (pyloo/psis.py:146,156); ours is (value, draw index).  Compare as multisets per row.
        out, ref_lw = torch.sort(out, dim=1).values, np.sort(ref_lw, axis=1)
    errs = {"lw": err(out, ref_lw), "k": err(k, ref_k), "elpd": err(res["elpd_i"], pw["elpd_i"]),
            "loo_k": err(res["pareto_k"], pw["pareto_k"]), "lppd": err(res["lppd_i"], pw["lppd_i"])}
    for k_, v in errs.items():
        worst[k_] = max(worst.get(k_, 0.0), v)
    if any(v > 1e-10 for v in errs.values()):
        fails.append({"case": case, "S": S, "N": N, "reff": reff, "kind": kind, "errs": errs})
    print(case, S, N, reff, kind, {k_: f"{v:.1e}" for k_, v in errs.items()}, "handover", ho, flush=True)
print(json.dumps({"cases": n_cases, "worst": worst, "fails": fails}))

"""Dev tool: the tile path (pl.loo on the (S, N) layout) against the oracle + per-kernel timings."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyloo_b200 import engine
from oracle import psis_oracle as orc

res = {}
S = int(os.environ.get("S", 4000)); reff = float(os.environ.get("REFF", 1.0))
if os.environ.get("PARITY", "1") == "1":
    for N in (702, 16, 4096):
        rng = np.random.default_rng(N)
        ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.3, 2.0, size=(1, N))
        t = torch.from_numpy(ll).cuda()
        r = engine.loo_cuda(t, reff, want_diag=True); torch.cuda.synchronize()
        n_chk = min(N, 160)
        pw = orc.loo_pointwise(ll[:, :n_chk], reff); ww = orc.waic_pointwise(ll[:, :n_chk])
        def rel(a, b): return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
        d = r["diag"].cpu().numpy()
        res[f"parity_N{N}"] = {
            "elpd": rel(r["elpd_i"][:n_chk].cpu().numpy(), pw["elpd_i"]), "k_abs": float(np.max(np.abs(r["pareto_k"][:n_chk].cpu().numpy() - pw["pareto_k"]))),
            "lppd": rel(r["lppd_i"][:n_chk].cpu().numpy(), pw["lppd_i"]), "var": rel(r["var_i"][:n_chk].cpu().numpy(), ww["var_i"]),
            "fallback": int(r["counters"][3]), "cand_mean": float(d[:, 3].mean()), "cand_min": float(d[:, 3].min()), "cand_max": float(d[:, 3].max()),
            "handover": engine.handover_reasons()}
N = int(os.environ.get("N", 151552))
torch.manual_seed(0)
ll = torch.randn(S, N, dtype=torch.float64, device="cuda") - 1.4
r = engine.loo_cuda(ll, reff); torch.cuda.synchronize()
ws = r["workspace"]
for _ in range(2): r = engine.loo_cuda(ll, reff, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): r = engine.loo_cuda(ll, reff, workspace=ws)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
engine.profile(True)
r = engine.loo_cuda(ll, reff, workspace=ws); torch.cuda.synchronize()
prof = engine.profile_read(); engine.profile(False)
res["time"] = {"N": N, "S": S, "ms": ms, "Mobs_s": N / ms / 1e3, "GBs": N * (8 * S + 40) / ms / 1e6, "frac_6549": N * (8 * S + 40) / ms / 1e6 / 6549.1,
               "fallback": int(r["counters"][3]), "handover": engine.handover_reasons(), "prof_ms": {k: v for k, v in prof.items() if v[1]},
               "env": {k: v for k, v in os.environ.items() if k.startswith("B2L_")}}
print(json.dumps(res, indent=1))

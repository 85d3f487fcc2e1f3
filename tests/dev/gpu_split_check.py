"""Dev tool: small split-path check against the oracle (psislw + loo), prints max errors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyloo_b200 import engine
from oracle import psis_oracle as orc
N = int(os.environ.get("N", 64)); S = int(os.environ.get("S", 4000)); reff = float(os.environ.get("REFF", 0.9))
rng = np.random.default_rng(3)
x = rng.normal(size=(N, S))
lw, k, diag = engine.psislw_cuda(torch.from_numpy(x).cuda(), reff, want_diag=True)
torch.cuda.synchronize()
ref_lw, ref_k = orc.psislw(x, reff)
lw = lw.cpu().numpy(); k = k.cpu().numpy(); diag = diag.cpu().numpy()
print("psislw k maxrel", np.max(np.abs(k - ref_k) / np.abs(ref_k)), "lw maxabs", np.max(np.abs(lw - ref_lw)))
print("C mean/min/max", diag[:, 3].mean(), diag[:, 3].min(), diag[:, 3].max(), "attempts max", diag[:, 4].max(), "n", diag[:, 2].min(), diag[:, 2].max())
ll = -1.4 + rng.normal(size=(S, N))
r = engine.loo_cuda(torch.from_numpy(ll).cuda(), 1.0)
torch.cuda.synchronize()
pw = orc.loo_pointwise(ll, 1.0); ww = orc.waic_pointwise(ll)
for a, b in (("elpd_i", pw["elpd_i"]), ("pareto_k", pw["pareto_k"]), ("lppd_i", pw["lppd_i"]), ("var_i", ww["var_i"])):
    g = r[a].cpu().numpy()
    print("loo", a, "maxrel", np.max(np.abs(g - b) / np.abs(b)))
print("fallback rows", int(r["counters"][3]))

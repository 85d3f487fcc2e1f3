"""Dev tool: a few small calls of every fast path, for `compute-sanitizer --tool memcheck python tests/dev/gpu_sanitize_case.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyloo_b200 import engine

rng = np.random.default_rng(0)
for S, N, reff in ((4000, 70, 1.0), (2000, 36, 0.5), (1000, 20, 1.0), (16000, 18, 1.0), (8000, 10, 0.25), (4000, 22, 0.2)):
    ll = torch.from_numpy(-1.4 + rng.normal(size=(S, N))).cuda()
    r = engine.loo_cuda(ll, reff, want_diag=True, want_tail_idx=True)
    torch.cuda.synchronize()
    print("loo", S, N, reff, float(r["elpd_i"].sum()), int(r["counters"][3]), flush=True)
for S, N, reff in ((4000, 40, 0.9), (600, 33, 1.0), (4000, 24, 0.1)):
    lw = torch.from_numpy(rng.normal(size=(N, S))).cuda()
    out, k = engine.psislw_cuda(lw, reff)
    torch.cuda.synchronize()
    print("psislw", S, N, reff, float(k.mean()), flush=True)
ll = -1.4 + rng.standard_t(3, size=(2000, 40))
ll[3, 5] = np.nan
r = engine.loo_host(ll, 1.0, chunk_obs=16, device=0)
print("host", float(np.nansum(r["elpd_i"])), flush=True)

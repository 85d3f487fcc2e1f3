"""Dev tool: every small observation count on every kind of plan (cluster sizes 2 / 4 / 8, chunked units, long tails):
no hang, no non-finite output, oracle parity on a sample of the calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyloo_b200 import engine
from oracle import psis_oracle as orc

rng = np.random.default_rng(9)
calls = checked = 0
worst = 0.0
for S in (512, 1000, 2000, 4000, 8000, 16000):
    for reff in (1.0, 0.5, 0.25):
        M = engine.tail_length(S, reff)
        plan = engine.tile_shape_info(S, M)
        for N in list(range(2, 66, 2)) + [130, 258, 1030]:
            ll = -1.4 + rng.normal(size=(S, N))
            r = engine.loo_cuda(torch.from_numpy(ll).cuda(), reff)
            torch.cuda.synchronize()
            calls += 1
            e = r["elpd_i"].cpu().numpy()
            assert np.isfinite(e).all() and np.isfinite(r["pareto_k"].cpu().numpy()).all(), (S, reff, N)
            if N in (2, 14, 34, 62, 130):
                pw = orc.loo_pointwise(ll[:, :8], reff)
                worst = max(worst, float(np.max(np.abs(e[:8] - pw["elpd_i"][:8]) / np.abs(pw["elpd_i"][:8]))))
                checked += 1
        print(S, reff, M, {k: plan[k] for k in ("eligible", "cluster_size", "n_chunks")}, "ok", flush=True)
print("calls", calls, "checked", checked, "worst rel err", worst)

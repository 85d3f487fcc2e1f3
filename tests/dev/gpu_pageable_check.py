"""Dev tool: end-to-end rate of loo_host / psislw_host from PAGEABLE NumPy memory (what pl.loo receives), by staging
strategy: streaming stores on / off (B2L_HOST_NT), host copy threads (B2L_HOST_THREADS)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyloo_b200 import engine

S, N = 4000, int(os.environ.get("N", 120000))
rng = np.random.default_rng(0)
pag = np.empty((S, N))
for s0 in range(0, S, 500):
    pag[s0:s0 + 500] = rng.standard_normal((500, N)) - 1.4
engine.loo_host(pag[:, :4096], 1.0, device=0)
ref = engine.loo_host(pag, 1.0, device=0)
out = {"cores": os.cpu_count()}
for nt in ("1", "0"):
    for th in ("", "12"):
        os.environ["B2L_HOST_NT"] = nt
        if th: os.environ["B2L_HOST_THREADS"] = th
        else: os.environ.pop("B2L_HOST_THREADS", None)
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            r = engine.loo_host(pag, 1.0, device=0)
            best = min(best, time.perf_counter() - t0)
        assert np.array_equal(r["elpd_i"], ref["elpd_i"], equal_nan=True) and np.array_equal(r["pareto_k"], ref["pareto_k"], equal_nan=True)
        out[f"nt={nt} threads={th or 'auto'}"] = round(N / best / 1e6, 3)
os.environ.pop("B2L_HOST_THREADS", None); os.environ.pop("B2L_HOST_NT", None)
# psislw both ways through pageable memory
Np = 30000
x = rng.standard_normal((Np, S))
engine.psislw_host(x[:2048], 0.9, device=0)
engine.psislw_host(x, 0.9, device=0)          # (allocates the full-size bounce buffers once)
for nt in ("1", "0", "1", "0"):
    os.environ["B2L_HOST_NT"] = nt
    t0 = time.perf_counter(); lw, k = engine.psislw_host(x, 0.9, device=0); t = time.perf_counter() - t0
    out[f"psislw nt={nt}"] = max(out.get(f"psislw nt={nt}", 0.0), round(Np / t / 1e6, 3))
    if nt == "1": lw1 = lw.copy()
    else: assert np.array_equal(lw, lw1)
print(json.dumps(out, indent=1))

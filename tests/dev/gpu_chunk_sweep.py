import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pyloo_b200 import engine
S, N = 4000, 120000
rng = np.random.default_rng(0)
pag = np.empty((S, N))
for s0 in range(0, S, 500): pag[s0:s0+500] = rng.standard_normal((500, N)) - 1.4
pin = torch.empty((S, N), dtype=torch.float64, pin_memory=True); pin.copy_(torch.from_numpy(pag)); pinn = pin.numpy()
for chunk in (2368, 4736, 9472, 18944):
    for name, arr in (("pageable", pag), ("pinned", pinn)):
        engine.loo_host(arr, 1.0, device=0, chunk_obs=chunk)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); engine.loo_host(arr, 1.0, device=0, chunk_obs=chunk); best = min(best, time.perf_counter() - t0)
        print(chunk, name, round(N / best / 1e6, 3), flush=True)

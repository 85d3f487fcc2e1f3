"""Device-resident timing of the SIS / TIS and e_loo kernels (CUDA events, inputs larger than L2).

    python tools/gpu_is_time.py [S] [N]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from pyloo_b200 import engine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 64 * 2
PEAK = 6549.1  # MEASURED_PEAKS.json hbm_gbs

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
lr = torch.randn((N, S), dtype=torch.float64, device=dev, generator=g)
x = torch.randn((N, S), dtype=torch.float64, device=dev, generator=g)
ll_sn = (-1.4 + torch.randn((S, N), dtype=torch.float64, device=dev, generator=g))
out = torch.empty_like(lr)
lw, _ = engine.islw_cuda(lr, "tis")


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows = []
for name, fn, bytes_per_obs in (
    ("sis weights", lambda: engine.islw_cuda(lr, "sis", out=out), 16 * S + 8),
    ("tis weights", lambda: engine.islw_cuda(lr, "tis", out=out), 16 * S + 8),
    ("sis loo (S,N)", lambda: engine.loo_is_cuda(ll_sn, "sis"), 8 * S + 24),
    ("tis loo (S,N)", lambda: engine.loo_is_cuda(ll_sn, "tis"), 8 * S + 24),
    ("tis loo rows", lambda: engine.loo_is_cuda(lr.t(), "tis"), 8 * S + 24),
    ("e_loo mean (lw only)", lambda: engine.eloo_cuda(x, lw, None, "mean"), 16 * S + 16),
    ("e_loo mean (lw + lr)", lambda: engine.eloo_cuda(x, lw, lr, "mean"), 24 * S + 16),
    ("e_loo sd (lw + lr)", lambda: engine.eloo_cuda(x, lw, lr, "sd"), 24 * S + 16),
    ("e_loo k only", lambda: engine.eloo_cuda(None, lw, lr, "none"), 16 * S + 8),
    ("e_loo quantiles x3", lambda: engine.eloo_quantile_cuda(x, lw, [0.05, 0.5, 0.95]), 16 * S + 24),
):
    ms = timeit(fn)
    gbs = bytes_per_obs * N / ms / 1e6
    rows.append({"kernel": name, "S": S, "N": N, "ms": round(ms, 4), "obs_per_s": round(N / ms * 1e3),
                 "algorithmic_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 3)})
    print(json.dumps(rows[-1]))

"""Short single-GPU program for ncu captures: a few psislw launches (S=4000) and loo launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyloo_b200 import engine
mode = sys.argv[1] if len(sys.argv) > 1 else "psislw"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8880
S = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
torch.manual_seed(0)
x = torch.randn(N, S, dtype=torch.float64, device="cuda")
if mode == "psislw":
    out = torch.empty_like(x)
    for _ in range(4):
        engine.psislw_cuda(x, 0.9, out=out)
elif mode in ("sis", "tis"):
    out = torch.empty_like(x)
    for _ in range(3):
        engine.islw_cuda(x, mode, out=out)
elif mode == "eloo":
    lw, _ = engine.islw_cuda(x, "tis")
    h = torch.randn(N, S, dtype=torch.float64, device="cuda")
    for _ in range(3):
        engine.eloo_cuda(h, lw, x, "mean")
elif mode == "quant":
    lw, _ = engine.islw_cuda(x, "tis")
    h = torch.randn(N, S, dtype=torch.float64, device="cuda")
    for _ in range(3):
        engine.eloo_quantile_cuda(h, lw, [0.05, 0.5, 0.95])
elif mode == "loo_rows":
    for _ in range(4):
        engine.loo_cuda(x.t(), 1.0)
else:
    ll = x.t().contiguous()
    for _ in range(4):
        engine.loo_cuda(ll, 1.0)
torch.cuda.synchronize()
print("done", mode, N, S)

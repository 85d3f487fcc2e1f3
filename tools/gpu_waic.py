"""Dev tool: time the WAIC-only pass (obs-fastest layout) on device-resident data."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyloo_b200 import engine
N = int(os.environ.get("N", 40000)); S = int(os.environ.get("S", 4000))
torch.manual_seed(0)
ll = torch.randn(S, N, dtype=torch.float64, device="cuda") - 1.4
for wo in (True, False):
    fn = lambda: engine.loo_cuda(ll, 1.0, waic_only=wo)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"waic_only": wo, "Mobs_s": N / ms * 1e3 / 1e6, "GBs": N * (8 * S + 40) / ms / 1e6}))

"""Dev tool: time psislw / loo on device-resident data (CUDA events), optional env sweeps."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyloo_b200 import engine

N = int(os.environ.get("N", 40000)); S = int(os.environ.get("S", 4000)); reff = float(os.environ.get("REFF", 0.9))
mode = sys.argv[1] if len(sys.argv) > 1 else "psislw"
torch.manual_seed(0)
x = torch.randn(N, S, dtype=torch.float64, device="cuda")
if os.environ.get("DIST", "normal") == "t15":   # Student-t(1.5) log-ratios (BASELINE configs[4])
    import numpy as np
    rng = np.random.default_rng(0)
    blk = 2000
    for i0 in range(0, N, blk):
        x[i0:i0 + blk] = torch.from_numpy(rng.standard_t(1.5, size=(min(blk, N - i0), S))).cuda()
res = {}
if mode in ("psislw", "both"):
    out = torch.empty_like(x)
    fn = lambda: engine.psislw_cuda(x, reff, out=out)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res["psislw_Mobs_s"] = N / ms * 1e3 / 1e6
    res["k_gt_0.7"] = float((fn()[1] > 0.7).double().mean())
    res["psislw_GBs"] = N * (16 * S + 8) / ms / 1e6
if mode in ("loo", "both"):
    ll = x.t().contiguous() if os.environ.get("OBSFAST", "1") == "1" else x.t()
    fn = lambda: engine.loo_cuda(ll, 1.0)
    for _ in range(2): r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): r = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res["loo_Mobs_s"] = N / ms * 1e3 / 1e6
    res["loo_GBs"] = N * (8 * S + 40) / ms / 1e6
    res["fallback_rows"] = int(r["counters"][3])
res["handover"] = engine.handover_reasons()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("B2L_")}, "N": N, "S": S, **res}))

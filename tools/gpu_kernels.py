"""Dev tool: per-kernel device time (b2l_profile) for psislw / loo at a given shape."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyloo_b200 import engine
N = int(os.environ.get("N", 40000)); S = int(os.environ.get("S", 4000)); reff = float(os.environ.get("REFF", 0.9))
torch.manual_seed(0)
x = torch.randn(N, S, dtype=torch.float64, device="cuda")
out = torch.empty_like(x)
for name, fn in (("psislw", lambda: engine.psislw_cuda(x, reff, out=out)), ("loo", lambda: engine.loo_cuda(x.t().contiguous(), reff))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    engine.profile(True)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    prof = engine.profile_read(); engine.profile(False)
    tot = sum(ms for ms, _ in prof.values())
    print(json.dumps({"what": name, "S": S, "N": N, "M": engine.tail_length(S, reff), "Mobs_s": N / (tot / 3) * 1e3 / 1e6,
                      "kernels_us_per_call": {k: (round(ms / 3 * 1e3, 1), c // 3) for k, (ms, c) in prof.items() if c},
                      "launch": engine.split_launch_info(S, engine.tail_length(S, reff), name if name == "psislw" else "loo", N)}))

"""Dev tool: psislw rate by round size (B2L_BATCH): small rounds keep a round's rows in the 126 MB L2 for the apply
stage's second read, at the price of more launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pyloo_b200 import engine
S, N = 4000, 100000
x = torch.randn(N, S, dtype=torch.float64, device="cuda")
out = torch.empty_like(x)
for b in os.environ.get("BATCHES", "592,1184,2368,4736,9472,18944").split(","):
    os.environ["B2L_BATCH"] = b
    ws = engine.workspace_for(S, N, 0.9, False, "cuda")
    for _ in range(3): engine.psislw_cuda(x, 0.9, out=out, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): engine.psislw_cuda(x, 0.9, out=out, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    engine.profile(True); engine.psislw_cuda(x, 0.9, out=out, workspace=ws); torch.cuda.synchronize()
    prof = engine.profile_read(); engine.profile(False)
    print("batch", b, round(ms, 3), "ms", round(N / ms / 1e3, 2), "M obs/s", {k: (round(v[0], 3), v[1]) for k, v in prof.items() if v[1]}, flush=True)

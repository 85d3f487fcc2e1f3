import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyloo_b200 import engine
torch.manual_seed(0)
N, S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 4000
x = torch.randn(N, S, dtype=torch.float64, device="cuda")
out, k, diag = engine.psislw_cuda(x, 0.9, want_diag=True)
torch.cuda.synchronize()
d = diag.cpu().numpy()
print("cand: mean %.1f std %.1f min %d max %d" % (d[:,3].mean(), d[:,3].std(), d[:,3].min(), d[:,3].max()))
print("attempts histogram:", np.bincount(d[:,4].astype(int)))
print("ntail:", np.bincount(d[:,2].astype(int))[-3:], "frac cand>512: %.3f" % (d[:,3] > 512).mean())

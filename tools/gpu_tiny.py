"""Tiny GPU case for compute-sanitizer (dev tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyloo_b200 import engine
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.normal(size=(300, 1000))).cuda()
out, k = engine.psislw_cuda(x, 1.0)
r = engine.loo_cuda(x.t().contiguous().t().t(), 1.0)
r2 = engine.loo_cuda(torch.from_numpy(rng.normal(size=(1000, 70))).cuda(), 1.0)
st = engine.stats_cuda(r2)
torch.cuda.synchronize()
print("tiny ok", float(k.mean()), float(r["elpd_i"].mean()), st[:4].tolist())

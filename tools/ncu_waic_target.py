import os, sys
sys.path.insert(0, "/root/repo")
import torch
from pyloo_b200 import engine
torch.manual_seed(0)
ll = torch.randn(4000, 60000, dtype=torch.float64, device="cuda") - 1.4
for _ in range(3):
    engine.loo_cuda(ll, 1.0, waic_only=True)
torch.cuda.synchronize()

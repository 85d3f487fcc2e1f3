"""Dev tool: heavy-tailed (S, N) matrices -- how much of the tile path's work ends in the general kernel, and what the
panel route (B2L_TILE=0) takes for the same input."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pyloo_b200 import engine

def t_dist(df, shape, gen):
    z = torch.randn(shape, dtype=torch.float64, device="cuda", generator=gen)
    g = torch.distributions.Gamma(torch.tensor(df / 2, device="cuda", dtype=torch.float64),
                                  torch.tensor(0.5, device="cuda", dtype=torch.float64)).sample(shape)
    return z / torch.sqrt(g / df)

gen = torch.Generator(device="cuda"); gen.manual_seed(1)
out = {}
for name, S, N, df, scale in (("t1.5_S8000", 8000, 20000, 1.5, 1.0), ("t3_S4000", 4000, 40000, 3.0, 1.0),
                              ("t5_S4000", 4000, 40000, 5.0, 1.0), ("t3_S4000_x3", 4000, 40000, 3.0, 3.0)):
    ll = -1.4 + scale * t_dist(df, (S, N), gen)
    row = {}
    for mode in ("tile", "panel"):
        if mode == "panel": os.environ["B2L_TILE"] = "0"
        else: os.environ.pop("B2L_TILE", None)
        engine.handover_reasons()
        r = engine.loo_cuda(ll, 1.0); torch.cuda.synchronize()
        ho = engine.handover_reasons()
        ws = r["workspace"]
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3): r = engine.loo_cuda(ll, 1.0, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        row[mode] = {"ms": e0.elapsed_time(e1) / 3, "handed_over": int(sum(ho.values())), "reasons": ho,
                     "k_gt_0.7": float((r["pareto_k"] > 0.7).double().mean())}
    os.environ.pop("B2L_TILE", None)
    out[name] = row
    del ll
print(json.dumps(out, indent=1))

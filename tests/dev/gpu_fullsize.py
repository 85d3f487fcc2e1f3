"""Dev tool: BASELINE.json configs[2] and configs[4] at FULL single-GPU size (size-independent checks + timing)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyloo_b200 import engine
from oracle import psis_oracle as orc
res = {}
# ---- configs[2]: pl.loo + waic, S = 4000 x N = 10^6, (chain, draw, obs) layout, one GPU
S, N = 4000, 1_000_000
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
ll = torch.empty(S, N, dtype=torch.float64, device="cuda")
for i0 in range(0, S, 500):
    ll[i0:i0 + 500] = torch.randn(500, N, dtype=torch.float64, device="cuda", generator=gen) - 1.4
r = engine.loo_cuda(ll, 1.0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); r = engine.loo_cuda(ll, 1.0, workspace=r["workspace"]); st = engine.stats_cuda(r); e1.record(); torch.cuda.synchronize()
rec = engine.StatsRecord(st.cpu().numpy())
idx = torch.arange(0, N, 40009, device="cuda")
sub = ll[:, idx].cpu().numpy()
pw = orc.loo_pointwise(sub, 1.0)
err_e = float(np.max(np.abs(r["elpd_i"][idx].cpu().numpy() - pw["elpd_i"]) / np.abs(pw["elpd_i"])))
err_k = float(np.max(np.abs(r["pareto_k"][idx].cpu().numpy() - pw["pareto_k"])))
res["cfg2_loo_N1e6"] = {"ms": e0.elapsed_time(e1), "Mobs_s": N / e0.elapsed_time(e1) / 1e3, "elpd_loo": rec.elpd_sum, "n": rec.n,
                        "k_gt_good": rec.k_gt_good, "handed_over": int(r["counters"][3]), "max_rel_err_elpd_vs_oracle(25 obs)": err_e,
                        "max_abs_err_k": err_k}
del ll, r
torch.cuda.empty_cache()
# ---- configs[4]: heavy-tailed stress, S = 8000 x N = 500000, Student-t(1.5) log-ratios
S, N = 8000, 500_000
x = torch.empty(N, S, dtype=torch.float64, device="cuda")
rng = np.random.default_rng(3)
blk = 5000
for i0 in range(0, N, blk):   # Student-t(1.5) = normal / sqrt(chi2_1.5 / 1.5), generated on the device
    z = torch.randn(blk, S, dtype=torch.float64, device="cuda", generator=gen)
    g = torch.distributions.Gamma(torch.tensor(0.75, device="cuda", dtype=torch.float64), torch.tensor(0.5, device="cuda", dtype=torch.float64)).sample((blk, S))
    x[i0:i0 + blk] = z / torch.sqrt(g / 1.5)
    del z, g
out = torch.empty_like(x)
engine.psislw_cuda(x[:20000], 1.0, out=out[:20000]); torch.cuda.synchronize()
engine.handover_reasons()
e0.record(); _, k = engine.psislw_cuda(x, 1.0, out=out); e1.record(); torch.cuda.synchronize()
lse = torch.logsumexp(out[::997], dim=1)
rows = x[:8].cpu().numpy()
with np.errstate(all="ignore"):
    ref_lw, ref_k = orc.psislw(rows, 1.0)
res["cfg4_t15_S8000_N500k"] = {"ms": e0.elapsed_time(e1), "Mobs_s": N / e0.elapsed_time(e1) / 1e3, "frac_k_gt_0.7": float((k > 0.7).double().mean()),
                               "max_abs_logsumexp": float(lse.abs().max()), "handover": engine.handover_reasons(),
                               "max_abs_err_k_vs_oracle(8 obs)": float(np.max(np.abs(k[:8].cpu().numpy() - ref_k))),
                               "max_abs_err_lw_vs_oracle(8 obs)": float(np.max(np.abs(out[:8].cpu().numpy() - ref_lw)))}
print(json.dumps(res, indent=1))

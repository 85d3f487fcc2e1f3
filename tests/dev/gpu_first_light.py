"""First-light GPU check: run the kernels on several shapes and print detailed discrepancies
against the CPU oracle (dev tool; the real parity tests live in tests/)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pyloo_b200 import engine
from oracle import psis_oracle as orc

def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    with np.errstate(all="ignore"):
        d = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    d = np.where((a == b) | (np.isnan(a) & np.isnan(b)), 0.0, d)
    return np.nanmax(d) if d.size else 0.0

def case_psislw(name, x, reff):
    t = torch.from_numpy(x).cuda()
    out, k, diag = engine.psislw_cuda(t, reff, want_diag=True)
    torch.cuda.synchronize()
    out = out.cpu().numpy(); k = k.cpu().numpy(); diag = diag.cpu().numpy()
    with np.errstate(all="ignore"):
        lw, kk = orc.psislw(x, reff)
    M = orc.tail_length(x.shape[1], reff)
    print(f"[psislw {name}] N={x.shape[0]} S={x.shape[1]} M={M} max rel err lw={rel(out, lw):.3e} k={rel(k, kk):.3e}"
          f" | cand mean={diag[:,3].mean():.0f} max={diag[:,3].max():.0f} attempts>0: {(diag[:,4]>0).sum()} "
          f"ntail mean={diag[:,2].mean():.1f}")
    bad = np.where(~(np.isclose(k, kk, rtol=1e-9, atol=0, equal_nan=True)))[0]
    for i in bad[:5]:
        print("   bad row", i, "k gpu", k[i], "k ref", kk[i], "diag", diag[i])

def case_loo(name, ll_sn, reff):
    t = torch.from_numpy(ll_sn).cuda()
    res = engine.loo_cuda(t, reff, want_diag=True)
    st = engine.stats_cuda(res)
    torch.cuda.synchronize()
    with np.errstate(all="ignore"):
        pw = orc.loo_pointwise(ll_sn, reff); ww = orc.waic_pointwise(ll_sn)
    e = res["elpd_i"].cpu().numpy(); k = res["pareto_k"].cpu().numpy()
    print(f"[loo {name}] S={ll_sn.shape[0]} N={ll_sn.shape[1]} rel err elpd={rel(e, pw['elpd_i']):.3e} k={rel(k, pw['pareto_k']):.3e} "
          f"lppd={rel(res['lppd_i'].cpu().numpy(), pw['lppd_i']):.3e} var={rel(res['var_i'].cpu().numpy(), ww['var_i']):.3e} "
          f"lppdw={rel(res['lppdw_i'].cpu().numpy(), ww['lppd_i']):.3e}")
    s = engine.StatsRecord(st.cpu().numpy())
    print("   stats: n", s.n, "elpd_sum", s.elpd_sum, "ref", np.nansum(pw["elpd_i"]), "se", (s.n * s.elpd_m2 / s.n) ** 0.5,
          "ref", (len(e) * np.var(pw["elpd_i"])) ** 0.5, "fallback rows", s.n_fallback)

def main():
    print(torch.cuda.get_device_name(0))
    for mode in ("psislw", "loo"):
        print(mode, engine.row_launch_info(4000, 200, mode), engine.row_launch_info(16000, 380, mode))
    rng = np.random.default_rng(1)
    case_psislw("normal-small", rng.normal(size=(64, 1000)), 1.0)
    case_psislw("cfg2", rng.normal(size=(2048, 4000)), 0.9)
    case_psislw("s8", rng.normal(size=(16, 8)), 1.0)
    case_psislw("s33", rng.normal(size=(16, 33)), 1.0)
    case_psislw("odd-s", rng.normal(size=(32, 1001)), 1.0)
    case_psislw("student", rng.standard_t(1.5, size=(512, 8000)), 1.0)
    case_psislw("ties", np.round(rng.normal(size=(64, 2000)), 1), 1.0)
    x = rng.normal(size=(8, 600)); x[1, 17] = np.nan; x[2, 5] = np.inf; x[3, 9] = -np.inf; x[4, :] = 1.0
    case_psislw("special", x, 1.0)
    case_psislw("s16000", rng.normal(size=(96, 16000)), 1.0)
    # obs-fastest psislw (transpose in and out)
    xt = np.ascontiguousarray(rng.normal(size=(2000, 300)))
    t = torch.from_numpy(xt).cuda()
    out, k = engine.psislw_cuda(t.t(), 1.0, out=torch.empty_like(t).t())
    lw, kk = orc.psislw(xt.T, 1.0)
    print("[psislw obs-fastest] rel err", rel(out.cpu().numpy(), lw), rel(k.cpu().numpy(), kk))
    case_loo("cfg3", -1.4 + rng.normal(size=(4000, 1500)), 1.0)
    case_loo("cfg1", rng.normal(size=(2000, 8)), 0.7)
    ll = -1.4 + rng.normal(size=(1000, 40)); ll[3, 2] = np.nan; ll[5, 4] = -np.inf; ll[7, 6] = np.inf
    case_loo("special", ll, 1.0)
    case_loo("student", -rng.standard_t(1.5, size=(8000, 300)), 1.0)
    case_loo("rows-layout", np.ascontiguousarray((-1.4 + rng.normal(size=(200, 4000)))).T, 1.0)
    # host entry points
    xh = rng.normal(size=(5000, 4000))
    t0 = time.time(); lwh, kh = engine.psislw_host(xh, 0.9, chunk_obs=1024); t1 = time.time()
    lw, kk = orc.psislw(xh[:64], 0.9)
    print(f"[psislw host] {t1-t0:.3f}s rel err", rel(lwh[:64], lw), rel(kh[:64], kk))
    llh = -1.4 + rng.normal(size=(4000, 3000))
    r = engine.loo_host(llh, 1.0, chunk_obs=1024)
    pw = orc.loo_pointwise(llh[:, :64], 1.0)
    print("[loo host] rel err", rel(r["elpd_i"][:64], pw["elpd_i"]), rel(r["pareto_k"][:64], pw["pareto_k"]), r["stats"])
    # quick timing
    for N, S, reff in ((20000, 4000, 0.9),):
        x = torch.randn(N, S, dtype=torch.float64, device="cuda")
        out = torch.empty_like(x)
        for _ in range(2): engine.psislw_cuda(x, reff, out=out)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(True), torch.cuda.Event(True)
        ev0.record()
        for _ in range(5): engine.psislw_cuda(x, reff, out=out)
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        print(f"[time psislw] N={N} S={S}: {ms:.3f} ms -> {N/ms*1e3:.3e} obs/s, {N*S*16/ms/1e6:.1f} GB/s algorithmic")
        ll = x.t().contiguous()  # (S, N) obs-fastest
        for _ in range(2): engine.loo_cuda(ll, 1.0)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(5): engine.loo_cuda(ll, 1.0)
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        print(f"[time loo obs-fastest] {ms:.3f} ms -> {N/ms*1e3:.3e} obs/s, {N*S*8/ms/1e6:.1f} GB/s algorithmic")
        for _ in range(2): engine.loo_cuda(x.t(), 1.0)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(5): engine.loo_cuda(x.t(), 1.0)
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        print(f"[time loo rows] {ms:.3f} ms -> {N/ms*1e3:.3e} obs/s, {N*S*8/ms/1e6:.1f} GB/s algorithmic")

if __name__ == "__main__":
    main()

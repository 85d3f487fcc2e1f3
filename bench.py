#!/usr/bin/env python
"""bench.py -- PSIS-LOO throughput on B200 (metric of BASELINE.json: observations/second at S = 4000).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload (``config.workload``): BASELINE.json configs[1] -- ``psislw`` on synthetic N(0,1) log-ratios,
S = 4000 draws x N = 100 000 observations per GPU, FP64, r_eff = 0.9 (M = 200), rows contiguous.
One step = one pass of the hot path over that batch.  Weak scaling: every rank owns its own
100 000-observation shard (observations are independent: no data-path collective).

Printed JSON (one line, rank 0):
  value      obs/s with the inputs resident in HBM (CUDA events, max over ranks)
  e2e        obs/s through the host-buffer C-ABI entry (pinned host in/out, H2D + D2H inside the timed region)
  roofline   dominant kernel (psis_stream_kernel): algorithmic bytes (16*S + 8 per observation x the
             observations its launches process) / its launch time, measured live with CUDA events around
             every launch on the launching stream (b2l_profile), vs the measured HBM copy peak;
             `path_frac` is the same bytes over the WHOLE step (all kernels), `kernels` the breakdown
  cpu_baseline  the oracle port (reference algorithm, NumPy, 1 core) on a bounded sample, same box
  loo        the fused loo + waic pass (configs[2] shard: S = 4000, (chain, draw, obs) layout), incl. the
             one NCCL exchange of the 32-double stats record when N > 1
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_DRAWS = 4000
N_OBS = 100_000
REFF = 0.9
LOO_REFF = 1.0


def env_int(name, default):
    return int(os.environ.get(name, default))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kernel_tag):
    """dram read + write bytes of one full launch (7104 observations) from the committed ncu --set full
    capture (profiles/traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            rec = json.load(fh).get(kernel_tag)
            return rec if rec is None else rec.get("dram_bytes_per_launch", rec)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        mhz, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                mhz.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(mhz)}


# ------------------------------------------------------------------------------------ CPU baseline
def _oracle_rows(args):
    """Worker: oracle psislw on a block of rows regenerated from a seed (no big pickles)."""
    seed, n_rows, S, reff = args
    from oracle import psis_oracle as orc

    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n_rows, S))
    t0 = time.perf_counter()
    orc.psislw(x, reff)
    return time.perf_counter() - t0


def cpu_baseline_one_core(budget_s=10.0):
    """Reference algorithm (oracle port, NumPy) on ONE core, bounded sample of the same workload."""
    from oracle import psis_oracle as orc

    rng = np.random.default_rng(1234)
    done, spent, block = 0, 0.0, 512
    orc.psislw(rng.normal(size=(16, S_DRAWS)), REFF)  # warm-up
    while spent < budget_s:
        x = rng.normal(size=(block, S_DRAWS))
        t0 = time.perf_counter()
        orc.psislw(x, REFF)
        spent += time.perf_counter() - t0
        done += block
    return {"value": done / spent, "unit": "obs/s", "cores": 1, "kind": "port",
            "sample": f"first {done} observations of the workload (S={S_DRAWS}, reff={REFF}), oracle/psis_oracle.py, "
                      f"NumPy {np.__version__}, {spent:.1f} s"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself cannot
    travel to the GPU box) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    import multiprocessing as mp

    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, 64))
    per_worker = 384
    n_step = cores * per_worker
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(seed0):
            # workers regenerate their rows from a seed and time only the reference algorithm; they run
            # concurrently, so the step takes as long as the slowest worker
            return max(pool.map(_oracle_rows, [(seed0 + w, per_worker, S_DRAWS, REFF) for w in range(cores)],
                                chunksize=1))

        for w in range(args.warmup):
            step(1000 + 100 * w)
        t_total = sum(step(5000 + 100 * k) for k in range(args.steps))
    value = n_step * args.steps / t_total
    line = {
        "impl": "reference", "metric": "PSIS-LOO obs/sec at S=4000", "value": value, "unit": "obs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"pl.psislw S={S_DRAWS} reff={REFF} (configs[1]); bounded sample of {n_step} obs/step"},
        "cpu_baseline": {"value": value, "unit": "obs/s", "cores": cores, "kind": "port",
                         "sample": f"{n_step} observations per step, {cores} processes x {per_worker} rows, "
                                   f"oracle/psis_oracle.py (NumPy restatement of pyloo/psis.py; the Python reference "
                                   f"is not installable on the GPU box)"},
        "e2e": {"value": value, "unit": "obs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-obs", type=int, default=N_OBS, help="observations per GPU")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-loo", action="store_true")
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from pyloo_b200 import engine

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    os.environ["B2L_DEVICE"] = str(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created: send that to stderr so that
        # stdout carries the one JSON line only
        sys.stdout.flush()
        saved_out = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_out, 1)
            os.close(saved_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N, S = args.n_obs, S_DRAWS
    M = engine.tail_length(S, REFF)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018 + rank)
    x = torch.randn(N, S, dtype=torch.float64, device=dev, generator=gen)  # 3.2 GB > L2 (126 MB)
    out = torch.empty_like(x)
    ws = engine.workspace_for(S, N, REFF, False, dev)   # sized by the library (b2l_workspace_bytes)

    def step():
        return engine.psislw_cuda(x, REFF, out=out, workspace=ws)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        # nvidia-smi needs ~0.3 s to start: keep the same step running (untimed) around the timed region so that
        # every clock / throttle sample is taken under this load
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < 0.6:
            step()
            torch.cuda.synchronize()
        barrier()
        ev0.record()
        for _ in range(args.steps):
            _, k = step()
        ev1.record()
        barrier()
        t_post = time.perf_counter()
        while time.perf_counter() - t_post < 0.3:
            step()
            torch.cuda.synchronize()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    value = world * N / (ms_step * 1e-3)
    clocks = clk.summary()
    assert bool(torch.isfinite(k).all())

    # ---- per-kernel device time: a separate pass with CUDA events around every launch (b2l_profile)
    prof_steps = 3
    engine.profile(True)
    for _ in range(prof_steps):
        step()
    torch.cuda.synchronize()
    prof = engine.profile_read()
    engine.profile(False)
    kernel_names = {"stream": "psis_stream_kernel<256,16,PSISLW> (row pass + fused apply of the previous batch)",
                    "tail": "psis_tail_kernel<8,PSISLW>", "apply": "psis_apply_kernel",
                    "row": "psis_row_kernel<256,PSISLW> (hand-over rows)", "transpose": "transpose_f64_kernel",
                    "stats": "stats kernels", "is": "is_row_kernel", "eloo": "eloo_row_kernel"}
    tot_ms = sum(ms for ms, _ in prof.values()) or 1.0
    kernels = {k: {"name": kernel_names[k], "ms_per_step": ms / prof_steps, "launches_per_step": cnt / prof_steps,
                   "share": ms / tot_ms}
               for k, (ms, cnt) in prof.items() if cnt}
    launches_per_step = sum(v["launches_per_step"] for v in kernels.values())
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])

    peak, peak_src = measured_peak()
    alg_bytes = N * (16 * S + 8)
    path_achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    dom_ms, dom_launches = kernels[dom]["ms_per_step"], kernels[dom]["launches_per_step"]
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9   # = bytes per launch / average launch duration
    try:
        split = engine.split_launch_info(S, M, "psislw", N)
    except Exception:
        split = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic("psis_stream_kernel_psislw_s4000"),
                "kernel": kernels[dom]["name"],
                "algorithmic_bytes_per_launch": alg_bytes / dom_launches,
                "avg_launch_ms": dom_ms / dom_launches, "launches_per_step": dom_launches,
                "path_achieved": path_achieved, "path_frac": path_achieved / peak,
                "kernels": kernels, "peak_source": peak_src, "launch": split}

    # ---- fused loo + waic on the configs[2] shard shape: (chain, draw, obs) layout, reff = 1
    loo = None
    if not args.skip_loo:
        n_loo = 125_000 if N >= 100_000 else N
        ll = torch.randn(S, n_loo, dtype=torch.float64, device=dev, generator=gen).sub_(1.4)
        gk = engine.good_k_threshold(S)
        gathered = [torch.empty(32, dtype=torch.float64, device=dev) for _ in range(world)]
        wsl = None

        def loo_step():
            nonlocal wsl
            res = engine.loo_cuda(ll, LOO_REFF, workspace=wsl)
            wsl = res["workspace"]  # first call sizes it (b2l_workspace_bytes), later calls reuse it
            st = engine.stats_cuda(res, gk, workspace=wsl)
            if world > 1:
                dist.all_gather(gathered, st)   # the single exchange: 32 doubles per rank over NVLink
                return torch.stack(gathered)
            return st.unsqueeze(0)

        for _ in range(3):
            loo_step()
        barrier()
        ev0.record()
        for _ in range(args.steps):
            recs = loo_step()
        ev1.record()
        barrier()
        ms_loo = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
        merged = engine.stats_merge(list(recs.cpu().numpy()))
        loo_bytes = n_loo * (8 * S + 40)
        loo = {"workload": f"pl.loo + waic fused, S={S} x N={n_loo} per GPU, (chain,draw,obs) layout, reff=1 (configs[2] shard)",
               "value": world * n_loo / (ms_loo * 1e-3), "unit": "obs/s", "ms_per_step": ms_loo,
               "roofline": {"bound": "hbm", "achieved": loo_bytes / (ms_loo * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": loo_bytes / (ms_loo * 1e-3) / 1e9 / peak},
               "elpd_loo": merged.elpd_sum, "n_total": merged.n, "collective": "all_gather(32 f64)" if world > 1 else None}
        # waic alone (pl.waic, loo_compare(ic="waic")): the one-pass column kernel, no transposed panels
        def waic_step():
            return engine.loo_cuda(ll, LOO_REFF, workspace=wsl, waic_only=True)

        for _ in range(3):
            waic_step()
        barrier()
        ev0.record()
        for _ in range(args.steps):
            waic_step()
        ev1.record()
        barrier()
        ms_waic = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
        waic_bytes = n_loo * (8 * S + 24)
        loo["waic_only"] = {"value": world * n_loo / (ms_waic * 1e-3), "unit": "obs/s", "ms_per_step": ms_waic,
                            "roofline": {"bound": "hbm", "achieved": waic_bytes / (ms_waic * 1e-3) / 1e9, "peak": peak,
                                         "unit": "GB/s", "frac": waic_bytes / (ms_waic * 1e-3) / 1e9 / peak}}
        # pl.loo end to end through the host-buffer C-ABI entry: pinned (S, N) host log-likelihood in, pointwise
        # vectors + statistics record out (H2D of the matrix inside the timed region)
        if not args.skip_e2e:
            hll = torch.empty((S, n_loo), dtype=torch.float64, pin_memory=True)
            hll.copy_(ll)
            torch.cuda.synchronize()
            hll_np = hll.numpy()
            engine.loo_host(hll_np, LOO_REFF, device=local)
            barrier()
            n_e2e = max(2, min(args.steps, 5))
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                r_host = engine.loo_host(hll_np, LOO_REFF, device=local)
            t_loo = max_over_ranks((time.perf_counter() - t0) / n_e2e * 1e3) * 1e-3
            assert abs(r_host["stats"].elpd_sum - merged.elpd_sum / world) < 1e-6 * abs(merged.elpd_sum) or world > 1
            loo["e2e"] = {"value": world * n_loo / t_loo, "unit": "obs/s", "h2d_bytes_per_step": n_loo * S * 8,
                          "d2h_bytes_per_step": n_loo * 40 + 256, "ms_per_step": t_loo * 1e3, "steps": n_e2e,
                          "api": "b2l_loo_host_f64 via pyloo_b200.engine.loo_host (pinned host in)"}
            del hll
        del ll

    # ---- the callers either side of psislw: SIS / TIS weights and e_loo on a 2-round slab
    next_rows = None
    if not args.skip_loo:
        n_nx = min(N, 148 * 64 * 2)
        xs, outs = x[:n_nx], out[:n_nx]
        hs = torch.randn(n_nx, S, dtype=torch.float64, device=dev, generator=gen)
        lw_tis, _ = engine.islw_cuda(xs, "tis")

        def timed(fn, nbytes):
            for _ in range(3):
                fn()
            barrier()
            ev0.record()
            for _ in range(args.steps):
                fn()
            ev1.record()
            barrier()
            ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
            gbs = n_nx * nbytes / (ms * 1e-3) / 1e9
            return {"value": world * n_nx / (ms * 1e-3), "unit": "obs/s", "ms_per_step": ms,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}}

        next_rows = {
            "workload": f"S={S} x N={n_nx} per GPU, device resident, 606 MB per array (larger than L2)",
            "sislw": timed(lambda: engine.islw_cuda(xs, "sis", out=outs), 16 * S + 8),
            "tislw": timed(lambda: engine.islw_cuda(xs, "tis", out=outs), 16 * S + 8),
            "e_loo_mean": timed(lambda: engine.eloo_cuda(hs, lw_tis, xs, "mean"), 24 * S + 16),
        }
        del hs, lw_tis
        lls = torch.randn(S, n_nx, dtype=torch.float64, device=dev, generator=gen).sub_(1.4)   # (chain, draw, obs)
        next_rows["loo_sis"] = timed(lambda: engine.loo_is_cuda(lls, "sis"), 8 * S + 24)
        next_rows["loo_tis"] = timed(lambda: engine.loo_is_cuda(lls, "tis"), 8 * S + 24)
        del lls

    # ---- end to end through the host-buffer C-ABI entry (pinned host memory both ways)
    e2e = None
    if not args.skip_e2e:
        hx = torch.empty((N, S), dtype=torch.float64, pin_memory=True)
        hout = torch.empty((N, S), dtype=torch.float64, pin_memory=True)
        hx.copy_(x)
        torch.cuda.synchronize()
        hx_np, hout_np = hx.numpy(), hout.numpy()
        e2e_steps = max(2, min(args.steps, 5))
        engine.psislw_host(hx_np, REFF, out=hout_np, device=local)  # warm-up (allocates staging once)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, k_host = engine.psislw_host(hx_np, REFF, out=hout_np, device=local)
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        t_e2e = max_over_ranks(t_e2e * 1e3) * 1e-3
        assert np.isfinite(k_host).all()
        e2e = {"value": world * N / t_e2e, "unit": "obs/s", "h2d_bytes_per_step": N * S * 8,
               "d2h_bytes_per_step": N * S * 8 + N * 8, "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
               "api": "b2l_psislw_host_f64 via pyloo_b200.engine.psislw_host (pinned host in/out, 3 chunk streams)"}
        del hx, hout

    cpu = None
    if rank == 0 and not args.skip_cpu:
        cpu = cpu_baseline_one_core()

    if rank == 0:
        line = {
            "metric": "PSIS-LOO obs/sec at S=4000", "value": value, "unit": "obs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"pl.psislw S={S} x N={N} per GPU, FP64, r_eff={REFF} (M={M}), rows contiguous "
                                   f"(BASELINE configs[1])", "l2_policy": "inputs (3.2 GB) larger than L2",
                       "parallelism": f"obs-sharded x{world}, no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(round(launches_per_step * args.steps)), "clocks": clocks,
            "loo": loo, "next_rows": next_rows,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

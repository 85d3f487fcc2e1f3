#!/usr/bin/env python
"""bench.py -- PSIS-LOO throughput on B200 (metric of BASELINE.json: observations/second at S = 4000).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Headline workload (``config.workload``): BASELINE.json configs[2], the north-star target -- ``pl.loo`` + ``waic``
(one fused pass) on a synthetic log-likelihood of S = 4000 draws x N = 10^6 observations in the ArviZ
(chain, draw, obs) layout, FP64, r_eff = 1 (M = 190).  N is the TOTAL over the GPUs (``"scaling": "strong"``): every
rank owns N / g observations, generated on its own device.  One step = one pass of the hot path over the rank's
shard + the statistics record (+ its all-gather when g > 1: 32 doubles per rank, the only exchange).

Printed JSON (one line, rank 0):
  value         obs/s with the inputs resident in HBM (CUDA events, max over ranks)
  roofline      dominant kernel of the step: algorithmic bytes (8 S + 40 per observation x the observations its
                launches process) / its launch time, measured live with CUDA events around every launch on the
                launching stream (b2l_profile) vs the measured HBM copy peak; ``path_frac`` = the same bytes over the
                whole step; ``kernels`` the breakdown; ``traffic`` = DRAM bytes per launch of the committed ncu capture
  e2e           obs/s through the host-buffer C-ABI entry (``engine.loo_host`` -> ``b2l_loo_host_mgpu_f64``): pinned host
                log-likelihood in, pointwise vectors + record out, H2D + D2H inside the timed region; next to it the
                box's plain-copy H2D ceiling at the same size and the same call from PAGEABLE NumPy memory
  cpu_baseline  the oracle port (reference algorithm, NumPy, 1 core) on a bounded sample of the same workload
  psislw        BASELINE configs[1]: ``pl.psislw`` S = 4000 x N = 100 000, r_eff = 0.9, with its own roofline and e2e
  compare       BASELINE configs[3]: ``loo_compare`` of 4 models, S = 16 000 x N = 262 144 each, one model at a time on
                the device, stacking weights on the host, strided-subset check against the oracle
  stress        BASELINE configs[4]: S = 8000 x N = 500 000 Student-t(1.5) log-ratios (``psislw``): share of k > 0.7,
                hand-over count, subset check against the oracle
  cfg1          BASELINE configs[0]: 4 x 500 x 8 through ``pl.loo`` against the committed golden values
  (psislw / compare / stress / cfg1 / next_rows run at --gpus 1 only.)
"""

from __future__ import annotations

import argparse
import datetime
import json
import math
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
N_TOTAL = 1_000_000      # configs[2]
LOO_REFF = 1.0
N_PSISLW = 100_000       # configs[1]
REFF = 0.9
E2E_OBS_CAP = 250_000    # host-buffer runs use at most this many observations per rank (8 GB pinned)


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
    """{dram_bytes_per_launch, obs_per_launch} of the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            rec = json.load(fh).get(kernel_tag)
        return rec if rec and "obs_per_launch" in rec and "dram_bytes_per_launch" in rec else None
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


# ------------------------------------------------------------------------------------ CPU baseline (oracle port)
def _oracle_loo_block(args):
    """Worker: the reference's loo + waic arithmetic on a block of observations in the ArviZ layout (columns of a
    sample-major matrix, i.e. strided rows as pyloo/loo.py:189 makes them), regenerated from a seed."""
    seed, n_obs, S, reff = args
    from oracle import psis_oracle as orc

    rng = np.random.default_rng(seed)
    ll = -1.4 + rng.normal(size=(S, n_obs))
    t0 = time.perf_counter()
    orc.loo_pointwise(ll, reff)
    orc.waic_pointwise(ll)
    return time.perf_counter() - t0


def cpu_baseline_one_core(budget_s=10.0):
    """Reference algorithm (oracle port, NumPy) on ONE core, bounded sample of the headline workload."""
    block, done, spent = 256, 0, 0.0
    _oracle_loo_block((1, 16, S_DRAWS, LOO_REFF))  # warm-up
    seed = 1234
    while spent < budget_s:
        spent += _oracle_loo_block((seed, block, S_DRAWS, LOO_REFF))
        done += block
        seed += 1
    return {"value": done / spent, "unit": "obs/s", "cores": 1, "kind": "port",
            "sample": f"{done} observations of the workload (S={S_DRAWS}, (chain,draw,obs) layout, reff={LOO_REFF}): "
                      f"oracle/psis_oracle.py loo_pointwise + waic_pointwise, NumPy {np.__version__}, {spent:.1f} s"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself needs ArviZ / xarray
    and cannot travel to the GPU box) on all host cores, bounded sample of the headline workload per step."""
    if rank != 0:
        return
    import multiprocessing as mp

    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, 64))
    per_worker = 192
    n_step = cores * per_worker
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(seed0):
            # workers regenerate their block from a seed and time only the reference algorithm; they run concurrently,
            # so the step takes as long as the slowest worker
            return max(pool.map(_oracle_loo_block, [(seed0 + w, per_worker, S_DRAWS, LOO_REFF) for w in range(cores)],
                                chunksize=1))

        for w in range(args.warmup):
            step(1000 + 100 * w)
        t_total = sum(step(5000 + 100 * k) for k in range(args.steps))
    value = n_step * args.steps / t_total
    line = {
        "impl": "reference", "metric": "PSIS-LOO obs/sec at S=4000", "value": value, "unit": "obs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"pl.loo + waic S={S_DRAWS} x N={N_TOTAL} (chain,draw,obs) layout reff={LOO_REFF} "
                               f"(BASELINE configs[2]); bounded sample of {n_step} obs/step"},
        "cpu_baseline": {"value": value, "unit": "obs/s", "cores": cores, "kind": "port",
                         "sample": f"{n_step} observations per step, {cores} processes x {per_worker} observations, "
                                   f"oracle/psis_oracle.py (NumPy restatement of pyloo/loo.py + psis.py + waic.py; the "
                                   f"Python reference is not installable on the GPU box)"},
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
    ap.add_argument("--n-obs", type=int, default=N_TOTAL, help="observations in total (split over the GPUs)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="headline only (no psislw / compare / stress keys)")
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from pyloo_b200 import engine
    from pyloo_b200.distributed import shard_bounds

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created: send that to stderr so that
        # stdout carries the one JSON line only
        sys.stdout.flush()
        saved_out = os.dup(1)
        os.dup2(2, 1)
        try:
            # (a step's collectives finish in milliseconds; the CPU baseline on rank 0 keeps the others waiting ~1 min)
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
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

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    warmup = max(args.warmup, 3)
    peak, peak_src = measured_peak()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    S = S_DRAWS
    M = engine.tail_length(S, LOO_REFF)
    gk = engine.good_k_threshold(S)
    lo, hi = shard_bounds(args.n_obs, world, rank)
    lo, hi = lo // 16 * 16, (hi // 16 * 16 if rank < world - 1 else hi)   # whole tiles per shard
    n_loc = hi - lo
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018 + rank)

    # ---- headline: pl.loo + waic fused, (chain, draw, obs) layout, this rank's N / g observations resident in HBM
    ll = torch.empty((S, n_loc), dtype=torch.float64, device=dev)     # 32 GB at g = 1: far larger than L2
    for s0 in range(0, S, 250):
        ll[s0:s0 + 250] = torch.randn(min(250, S - s0), n_loc, dtype=torch.float64, device=dev, generator=gen).sub_(1.4)
    ws = engine.workspace_for(S, n_loc, LOO_REFF, True, dev)
    gathered = [torch.empty(32, dtype=torch.float64, device=dev) for _ in range(world)]

    def step():
        res = engine.loo_cuda(ll, LOO_REFF, workspace=ws)
        st = engine.stats_cuda(res, gk, workspace=ws)
        if world > 1:
            dist.all_gather(gathered, st)   # the single exchange: 32 doubles per rank over NVLink
            return res, torch.stack(gathered)
        return res, st.unsqueeze(0)

    for _ in range(warmup):
        step()
    barrier()
    # untimed steps around the timed region (below): their NUMBER must be the same on every rank -- a step holds a
    # collective -- so it comes from one measured step time, maximum over the ranks, not from each rank's own clock
    ev0.record()
    step()
    ev1.record()
    torch.cuda.synchronize()
    ms_est = max(max_over_ranks(ev0.elapsed_time(ev1)), 0.05)
    n_pre, n_post = int(min(2000, math.ceil(600.0 / ms_est))), int(min(1000, math.ceil(300.0 / ms_est)))
    with ClockSampler(local) as clk:
        # nvidia-smi needs ~0.3 s to start: keep the same step running (untimed) around the timed region so that
        # every clock / throttle sample is taken under this load
        for _ in range(n_pre):
            step()
            torch.cuda.synchronize()
        barrier()
        ev0.record()
        for _ in range(args.steps):
            res, recs = step()
        ev1.record()
        barrier()
        for _ in range(n_post):
            step()
            torch.cuda.synchronize()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    n_all = int(sum_over_ranks(n_loc))
    value = n_all / (ms_step * 1e-3)
    clocks = clk.summary()
    merged = engine.stats_merge(list(recs.cpu().numpy()))
    handed_over = int(res["counters"][3].item())
    assert merged.n == n_all and np.isfinite(merged.elpd_sum)

    # strided-subset parity check of the headline run against the oracle (test infrastructure as the checker)
    parity = None
    if rank == 0:
        from oracle import psis_oracle as orc
        idx = np.arange(0, n_loc, max(1, n_loc // 24))[:24]
        sub = ll[:, torch.from_numpy(idx).to(dev)].cpu().numpy()
        pw = orc.loo_pointwise(sub, LOO_REFF)
        got_e, got_k = res["elpd_i"].cpu().numpy()[idx], res["pareto_k"].cpu().numpy()[idx]
        parity = {"observations": int(idx.size),
                  "max_rel_err_elpd_i": float(np.max(np.abs(got_e - pw["elpd_i"]) / np.abs(pw["elpd_i"]))),
                  "max_abs_err_pareto_k": float(np.max(np.abs(got_k - pw["pareto_k"]))),
                  "k_gt_0.7_flags_equal": bool(np.array_equal(got_k > 0.7, pw["pareto_k"] > 0.7))}

    # ---- per-kernel device time: a separate pass with CUDA events around every launch (b2l_profile)
    prof_steps = 2
    engine.profile(True)
    for _ in range(prof_steps):
        step()
    torch.cuda.synchronize()
    prof = engine.profile_read()
    engine.profile(False)
    kernel_names = {"stream": "loo_tile_kernel<8> (cluster of 8 CTAs, 2-D TMA tiles of the (S, N) matrix)",
                    "tail": "psis_tail_kernel<8,LOO>", "apply": "psis_apply_kernel",
                    "row": "psis_row_kernel<256,LOO> (hand-over observations)", "transpose": "transpose_f64_kernel",
                    "stats": "stats_partial_kernel + stats_final_kernel", "is": "is_row_kernel", "eloo": "eloo_row_kernel"}
    tot_ms = sum(ms for ms, _ in prof.values()) or 1.0
    kernels = {k: {"name": kernel_names[k], "ms_per_step": ms / prof_steps, "launches_per_step": cnt / prof_steps,
                   "share": ms / tot_ms}
               for k, (ms, cnt) in prof.items() if cnt}
    launches_per_step = sum(v["launches_per_step"] for v in kernels.values())
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    alg_bytes = n_loc * (8 * S + 40)
    path_achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    dom_ms, dom_launches = kernels[dom]["ms_per_step"], kernels[dom]["launches_per_step"]
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9   # = bytes per launch / average launch duration
    traffic = recorded_traffic("loo_tile_kernel_s4000" if dom == "stream" else "psis_tail_kernel_loo_s4000")
    obs_per_launch = n_loc / dom_launches
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["dram_bytes_per_launch"] * obs_per_launch / traffic["obs_per_launch"]) if traffic else None,
                "traffic_source": (f"ncu --set full of one {traffic['obs_per_launch']}-observation launch "
                                   f"({traffic['dram_bytes_per_launch']} B), scaled to this run's launch size") if traffic else None,
                "kernel": kernels[dom]["name"], "algorithmic_bytes_per_launch": alg_bytes / dom_launches,
                "obs_per_launch": obs_per_launch, "avg_launch_ms": dom_ms / dom_launches, "launches_per_step": dom_launches,
                "path_achieved": path_achieved, "path_frac": path_achieved / peak,
                "kernels": kernels, "peak_source": peak_src}
    waic_only = None
    if world == 1:
        def waic_step():
            return engine.loo_cuda(ll, LOO_REFF, workspace=ws, waic_only=True)
        for _ in range(3):
            waic_step()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.steps):
            waic_step()
        ev1.record()
        torch.cuda.synchronize()
        ms_w = ev0.elapsed_time(ev1) / args.steps
        wb = n_loc * (8 * S + 24)
        waic_only = {"value": n_loc / (ms_w * 1e-3), "unit": "obs/s", "ms_per_step": ms_w,
                     "roofline": {"bound": "hbm", "achieved": wb / (ms_w * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": wb / (ms_w * 1e-3) / 1e9 / peak}}

    # ---- end to end through the host-buffer C-ABI entry: what pl.loo does with a NumPy log-likelihood
    e2e = None
    if not args.skip_e2e:
        n_e = min(n_loc, E2E_OBS_CAP)
        hll = torch.empty((S, n_e), dtype=torch.float64, pin_memory=True)
        hll.copy_(ll[:, :n_e])
        torch.cuda.synchronize()
        hll_np = hll.numpy()
        engine.loo_host(hll_np, LOO_REFF, device=local)   # warm-up (allocates the chunk slots once)
        n_rep = max(2, min(args.steps, 4))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_rep):
            r_host = engine.loo_host(hll_np, LOO_REFF, device=local)
        t_e = max_over_ranks((time.perf_counter() - t0) / n_rep * 1e3) * 1e-3
        n_e_all = int(sum_over_ranks(n_e))
        close_ = np.allclose(r_host["elpd_i"][:64], res["elpd_i"][:64].cpu().numpy(), rtol=1e-12, atol=0)
        # the box's ceiling for this transfer: the same number of bytes as plain cudaMemcpyAsync calls (256 MB pieces of
        # the same pinned buffer viewed flat, into one device buffer), all ranks at once
        flat = hll.view(-1)
        piece = 32 * 1024 * 1024   # doubles
        d_tmp = torch.empty(piece, dtype=torch.float64, device=dev)
        st_copy = torch.cuda.Stream(device=dev)
        d_tmp.copy_(flat[:piece], non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(st_copy):
            for c0 in range(0, flat.numel(), piece):
                c1 = min(flat.numel(), c0 + piece)
                d_tmp[:c1 - c0].copy_(flat[c0:c1], non_blocking=True)
        st_copy.synchronize()
        t_copy = max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3
        ceil_obs = n_e_all / t_copy
        e2e = {"value": n_e_all / t_e, "unit": "obs/s", "h2d_bytes_per_step": n_e * S * 8,
               "d2h_bytes_per_step": n_e * 40 + 256, "ms_per_step": t_e * 1e3, "steps": n_rep,
               "observations_per_rank": n_e, "matches_device_path": bool(close_),
               "api": "b2l_loo_host_mgpu_f64 via pyloo_b200.engine.loo_host (pinned host in, 3 chunk streams per device)",
               "h2d_ceiling": {"value": ceil_obs, "unit": "obs/s", "gbs": n_e_all * S * 8 / t_copy / 1e9,
                               "how": "the same pinned bytes as plain 256 MB cudaMemcpyAsync pieces, all ranks at once"},
               "frac_of_h2d_ceiling": (n_e_all / t_e) / ceil_obs}
        del d_tmp
        if world == 1:
            # the same call from pageable NumPy memory (what pl.loo receives from ArviZ): through the library's pinned
            # bounce buffers (default), plain (the driver stages the copies) and with the array host-registered
            pag = np.empty((S, n_e), dtype=np.float64)
            np.copyto(pag, hll_np)
            engine.loo_host(pag[:, :4096], LOO_REFF, device=local)

            def timed_host(env):
                for k_, v_ in env.items():
                    os.environ[k_] = v_
                try:
                    t0_ = time.perf_counter()
                    engine.loo_host(pag, LOO_REFF, device=local)
                    return time.perf_counter() - t0_
                finally:
                    for k_ in env:
                        del os.environ[k_]

            timed_host({})
            t_pag = min(timed_host({}), timed_host({}))
            t_plain = timed_host({"B2L_HOST_BOUNCE": "0"})
            t_reg = timed_host({"B2L_HOST_BOUNCE": "0", "B2L_HOST_REGISTER": "1"})
            e2e["pageable"] = {"value": n_e / t_pag, "unit": "obs/s", "ms": t_pag * 1e3,
                               "how": "default: host threads copy each chunk into pinned bounce buffers, then async DMA"}
            e2e["pageable_driver_staged"] = {"value": n_e / t_plain, "unit": "obs/s", "ms": t_plain * 1e3,
                                             "how": "B2L_HOST_BOUNCE=0: cudaMemcpy2DAsync straight from pageable memory"}
            e2e["pageable_host_registered"] = {"value": n_e / t_reg, "unit": "obs/s", "ms": t_reg * 1e3,
                                               "how": "B2L_HOST_BOUNCE=0 B2L_HOST_REGISTER=1: cudaHostRegister for the call"}
            del pag
        del hll
    del ll, res, ws
    torch.cuda.empty_cache()

    extra = {}
    if world == 1 and not args.skip_configs:
        extra = other_configs(args, torch, engine, dev, local, gen, peak, warmup)

    cpu = None
    if rank == 0 and not args.skip_cpu:
        cpu = cpu_baseline_one_core()

    if rank == 0:
        line = {
            "metric": "PSIS-LOO obs/sec at S=4000", "value": value, "unit": "obs/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"pl.loo + waic fused, S={S} x N={n_all} in total ({n_loc} on rank 0), (chain,draw,obs) "
                                   f"layout, FP64, r_eff={LOO_REFF} (M={M}) (BASELINE configs[2])",
                       "l2_policy": f"inputs ({n_loc * S * 8 / 1e9:.1f} GB per GPU) larger than L2",
                       "parallelism": f"obs-sharded x{world}, one all_gather of 32 f64 per rank" if world > 1
                                      else "one GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(round(launches_per_step * args.steps)), "clocks": clocks,
            "elpd_loo": merged.elpd_sum, "p_loo": merged.lppd_sum - merged.elpd_sum, "n_k_gt_good": merged.k_gt_good,
            "handed_over_rank0": handed_over, "parity_spot_check": parity, "waic_only": waic_only,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def other_configs(args, torch, engine, dev, local, gen, peak, warmup):
    """BASELINE configs[0], [1], [3], [4] and the widened rows -- single GPU."""
    from oracle import psis_oracle as orc

    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(steps):
            r = fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / steps, r

    # ---- the headline path at the relative efficiencies pl.loo derives from a multi-chain posterior's ESS
    # (pyloo/loo.py:204-216): M = 3 sqrt(S / reff) grows from 190 (reff 1) to 425 draws (reff 0.2)
    S, n_rs = S_DRAWS, 148 * 64 * 8 * 2
    llr = torch.randn(S, n_rs, dtype=torch.float64, device=dev, generator=gen) - 1.4
    sweep = {}
    for reff_ in (0.5, 0.3, 0.2):
        wsr = engine.workspace_for(S, n_rs, reff_, True, dev)
        engine.handover_reasons()
        ms_, r_ = timed(lambda: engine.loo_cuda(llr, reff_, workspace=wsr), args.steps)
        hand = engine.handover_reasons()
        idx = torch.arange(0, n_rs, n_rs // 6, device=dev)[:6]
        pw = orc.loo_pointwise(llr[:, idx].cpu().numpy(), reff_)
        gbs = n_rs * (8 * S + 40) / (ms_ * 1e-3) / 1e9
        sweep[f"reff_{reff_}"] = {"tail_length": engine.tail_length(S, reff_), "value": n_rs / (ms_ * 1e-3), "unit": "obs/s",
                                  "ms_per_step": ms_, "path_frac": gbs / peak,
                                  "handed_over_per_step": int(sum(hand.values())) // (args.steps + 3),
                                  "max_rel_err_elpd_i_vs_oracle(6 obs)":
                                      float(np.max(np.abs(r_["elpd_i"][idx].cpu().numpy() - pw["elpd_i"]) / np.abs(pw["elpd_i"])))}
        del wsr, r_
    out["loo_low_reff"] = {"workload": f"pl.loo S={S} x N={n_rs}, (S, N) layout, device resident, r_eff 0.5 / 0.3 / 0.2", **sweep}
    del llr
    torch.cuda.empty_cache()

    # ---- configs[1]: pl.psislw, S = 4000 x N = 100 000, r_eff = 0.9, rows contiguous
    S, N = S_DRAWS, N_PSISLW
    x = torch.randn(N, S, dtype=torch.float64, device=dev, generator=gen)
    outb = torch.empty_like(x)
    ws = engine.workspace_for(S, N, REFF, False, dev)
    ms, (_, k) = timed(lambda: engine.psislw_cuda(x, REFF, out=outb, workspace=ws), args.steps)
    engine.profile(True)
    for _ in range(2):
        engine.psislw_cuda(x, REFF, out=outb, workspace=ws)
    torch.cuda.synchronize()
    prof = engine.profile_read()
    engine.profile(False)
    alg = N * (16 * S + 8)
    kms = {k_: v[0] / 2 for k_, v in prof.items() if v[1]}
    dom = max(kms, key=kms.get)
    traffic = recorded_traffic("psis_stream_kernel_psislw_s4000")
    launches = prof[dom][1] / 2
    ps = {"workload": f"pl.psislw S={S} x N={N}, FP64, r_eff={REFF}, rows contiguous (BASELINE configs[1])",
          "value": N / (ms * 1e-3), "unit": "obs/s", "ms_per_step": ms,
          "roofline": {"bound": "hbm", "kernel": f"{dom} kernel", "achieved": alg / (kms[dom] * 1e-3) / 1e9, "peak": peak,
                       "unit": "GB/s", "frac": alg / (kms[dom] * 1e-3) / 1e9 / peak,
                       "path_frac": alg / (ms * 1e-3) / 1e9 / peak, "kernel_ms_per_step": kms,
                       "launches_per_step": launches, "obs_per_launch": N / launches,
                       "traffic": (traffic["dram_bytes_per_launch"] * (N / launches) / traffic["obs_per_launch"]) if traffic else None}}
    assert bool(torch.isfinite(k).all())
    if not args.skip_e2e:
        hx = torch.empty((N, S), dtype=torch.float64, pin_memory=True)
        hout = torch.empty((N, S), dtype=torch.float64, pin_memory=True)
        hx.copy_(x)
        torch.cuda.synchronize()
        engine.psislw_host(hx.numpy(), REFF, out=hout.numpy(), device=local)
        t0 = time.perf_counter()
        for _ in range(3):
            engine.psislw_host(hx.numpy(), REFF, out=hout.numpy(), device=local)
        t_e = (time.perf_counter() - t0) / 3
        ps["e2e"] = {"value": N / t_e, "unit": "obs/s", "h2d_bytes_per_step": N * S * 8, "d2h_bytes_per_step": N * S * 8 + N * 8,
                     "api": "b2l_psislw_host_mgpu_f64 via engine.psislw_host (pinned host in / out)"}
        del hx, hout
    out["psislw"] = ps

    # ---- widened rows on a two-round slab of the same matrix
    n_nx = 148 * 64 * 2
    xs, outs = x[:n_nx], outb[:n_nx]
    hs = torch.randn(n_nx, S, dtype=torch.float64, device=dev, generator=gen)
    lw_tis, _ = engine.islw_cuda(xs, "tis")

    def row(fn, nbytes):
        ms_, _ = timed(fn, args.steps)
        gbs = n_nx * nbytes / (ms_ * 1e-3) / 1e9
        return {"value": n_nx / (ms_ * 1e-3), "unit": "obs/s", "ms_per_step": ms_,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}}

    out["next_rows"] = {"workload": f"S={S} x N={n_nx}, device resident",
                        "sislw": row(lambda: engine.islw_cuda(xs, "sis", out=outs), 16 * S + 8),
                        "tislw": row(lambda: engine.islw_cuda(xs, "tis", out=outs), 16 * S + 8),
                        "e_loo_mean": row(lambda: engine.eloo_cuda(hs, lw_tis, xs, "mean"), 24 * S + 16)}
    del x, outb, ws, hs, lw_tis, xs, outs
    torch.cuda.empty_cache()

    # ---- configs[3]: loo_compare of 4 models, S = 16 000 x N = 262 144 each, one model at a time
    S3, N3 = 16000, 262_144
    import pyloo_b200 as pl
    from pyloo_b200.elpd import ELPDData
    from pyloo_b200.data import LiteDataArray

    elpds, t_models, checks = {}, [], []
    ll3 = torch.empty((S3, N3), dtype=torch.float64, device=dev)
    ws3 = engine.workspace_for(S3, N3, 1.0, True, dev)
    gk3 = engine.good_k_threshold(S3)
    for kmod in range(4):
        for s0 in range(0, S3, 500):
            z = torch.randn(500, N3, dtype=torch.float64, device=dev, generator=gen)
            ll3[s0:s0 + 500] = (-1.4 - 0.1 * kmod) + (1.0 + 0.1 * kmod) * z
        del z
        torch.cuda.synchronize()
        ev0.record()
        r3 = engine.loo_cuda(ll3, 1.0, workspace=ws3)
        st3 = engine.stats_cuda(r3, gk3, workspace=ws3)
        ev1.record()
        torch.cuda.synchronize()
        t_models.append(ev0.elapsed_time(ev1))
        rec = engine.StatsRecord(st3.cpu().numpy())
        loo_i = r3["elpd_i"].cpu().numpy()
        n = rec.n
        elpds[f"m{kmod}"] = ELPDData(
            [rec.elpd_sum, float((n * rec.elpd_m2 / n) ** 0.5), rec.lppd_sum - rec.elpd_sum, S3, int(n),
             bool(rec.k_gt_good > 0), LiteDataArray(loo_i, ("obs",)), "log"],
            index=["elpd_loo", "se", "p_loo", "n_samples", "n_data_points", "warning", "loo_i", "scale"])
        idx = np.arange(0, N3, N3 // 6)[:6]
        pw = orc.loo_pointwise(ll3[:, torch.from_numpy(idx).to(dev)].cpu().numpy(), 1.0)
        checks.append(float(np.max(np.abs(loo_i[idx] - pw["elpd_i"]) / np.abs(pw["elpd_i"]))))
    t0 = time.perf_counter()
    df = pl.loo_compare(elpds, ic="loo", method="stacking")
    t_stack = time.perf_counter() - t0
    out["compare"] = {"workload": f"pl.loo_compare of 4 models, S={S3} x N={N3} each, stacking weights (BASELINE configs[3])",
                      "loo_ms_per_model": t_models, "value": 4 * N3 / (sum(t_models) * 1e-3), "unit": "obs/s",
                      "roofline": {"bound": "hbm", "achieved": 4 * N3 * (8 * S3 + 40) / (sum(t_models) * 1e-3) / 1e9,
                                   "peak": peak, "unit": "GB/s", "frac": 4 * N3 * (8 * S3 + 40) / (sum(t_models) * 1e-3) / 1e9 / peak},
                      "stacking_host_ms": t_stack * 1e3, "ranking": list(df.index),
                      "weights": [float(w) for w in df["weight"]], "elpd_loo": [float(v) for v in df["elpd_loo"]],
                      "max_rel_err_elpd_i_vs_oracle(6 obs per model)": max(checks)}
    del ll3, ws3, r3
    torch.cuda.empty_cache()

    # ---- configs[4]: heavy-tailed stress, S = 8000 x N = 500 000 Student-t(1.5) log-ratios
    S4, N4 = 8000, 500_000
    x4 = torch.empty((N4, S4), dtype=torch.float64, device=dev)
    blk = 5000
    a15 = torch.tensor(0.75, device=dev, dtype=torch.float64)
    b15 = torch.tensor(0.5, device=dev, dtype=torch.float64)
    for i0 in range(0, N4, blk):   # Student-t(1.5) = normal / sqrt(chi2_1.5 / 1.5)
        z = torch.randn(blk, S4, dtype=torch.float64, device=dev, generator=gen)
        g = torch.distributions.Gamma(a15, b15).sample((blk, S4))
        x4[i0:i0 + blk] = z / torch.sqrt(g / 1.5)
    del z, g
    out4 = torch.empty_like(x4)
    ws4 = engine.workspace_for(S4, N4, 1.0, False, dev)
    engine.psislw_cuda(x4[:20000], 1.0, out=out4[:20000])
    torch.cuda.synchronize()
    engine.handover_reasons()
    ev0.record()
    _, k4 = engine.psislw_cuda(x4, 1.0, out=out4, workspace=ws4)
    ev1.record()
    torch.cuda.synchronize()
    ms4 = ev0.elapsed_time(ev1)
    rows = x4[:6].cpu().numpy()
    with np.errstate(all="ignore"):
        ref_lw, ref_k = orc.psislw(rows, 1.0)
    hand = engine.handover_reasons()
    out["stress"] = {"workload": f"pl.psislw S={S4} x N={N4}, Student-t(1.5) log-ratios, r_eff=1 (BASELINE configs[4])",
                     "value": N4 / (ms4 * 1e-3), "unit": "obs/s", "ms": ms4,
                     "roofline": {"bound": "hbm", "achieved": N4 * (16 * S4 + 8) / (ms4 * 1e-3) / 1e9, "peak": peak,
                                  "unit": "GB/s", "frac": N4 * (16 * S4 + 8) / (ms4 * 1e-3) / 1e9 / peak},
                     "frac_k_gt_0.7": float((k4 > 0.7).double().mean()), "handed_over": int(sum(hand.values())),
                     "handover_reasons": hand,
                     "max_abs_err_k_vs_oracle(6 obs)": float(np.nanmax(np.abs(k4[:6].cpu().numpy() - ref_k))),
                     "max_abs_err_lw_vs_oracle(6 obs)": float(np.nanmax(np.abs(out4[:6].cpu().numpy() - ref_lw)))}
    del x4, out4, ws4
    torch.cuda.empty_cache()

    # ---- configs[0]: 4 chains x 500 draws x 8 observations through pl.loo against the committed golden values
    try:
        with np.load(os.path.join(ROOT, "tests", "golden", "cfg1_create_model.npz")) as g1:
            ll1 = g1["ll_sn"]
            ref1 = g1["elpd_i_r10"]
        t0 = time.perf_counter()
        r1 = engine.loo_host(ll1, 1.0, device=local)
        t1 = time.perf_counter() - t0
        out["cfg1"] = {"workload": "pl.loo 4 x 500 x 8 (BASELINE configs[0])", "ms": t1 * 1e3,
                       "max_rel_err_elpd_i_vs_reference_golden": float(np.max(np.abs(r1["elpd_i"] - ref1) / np.abs(ref1)))}
    except Exception as err:  # the golden file is part of the repo; report rather than hide a problem
        out["cfg1"] = {"error": repr(err)}
    return out


if __name__ == "__main__":
    main()

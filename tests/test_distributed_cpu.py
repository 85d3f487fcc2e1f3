"""World-size-2 gloo test (CPU) of the multi-GPU host logic: shard bounds, the single all-gather of the
32-double record, Chan merge in rank order, and the ELPD rows derived from the merged record.
Per-rank records are built from oracle pointwise values (the GPU kernels are covered by -m gpu tests)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record_from_pointwise(elpd, k, lppd, var, lppdw, good_k):
    """NumPy statement of the device statistics kernel (csrc: stats_partial_kernel) for one shard."""
    w = lppdw - var
    r = np.zeros(32)
    r[0] = len(elpd)
    r[1] = elpd.mean(); r[2] = ((elpd - elpd.mean()) ** 2).sum(); r[3] = elpd.sum()
    r[4] = lppd.sum(); r[5] = var.sum()
    r[6] = w.mean(); r[7] = ((w - w.mean()) ** 2).sum(); r[8] = w.sum()
    r[9] = (k > good_k).sum(); r[10] = (k > 1).sum(); r[11] = np.isposinf(k).sum(); r[12] = np.isnan(k).sum()
    r[13] = (var > 0.4).sum(); r[14] = elpd.min(); r[15] = elpd.max(); r[16] = w.min(); r[17] = w.max()
    return r


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import psis_oracle as orc
    from pyloo_b200 import distributed as bd, engine

    rng = np.random.default_rng(123)
    ll = -1.4 + rng.normal(size=(600, 37))              # same matrix on every rank; each takes its shard
    lo, hi = bd.shard_bounds(ll.shape[1], world, rank)
    pw = orc.loo_pointwise(ll[:, lo:hi], 1.0)
    ww = orc.waic_pointwise(ll[:, lo:hi])
    gk = engine.good_k_threshold(600)
    rec = record_from_pointwise(pw["elpd_i"], pw["pareto_k"], pw["lppd_i"], ww["var_i"], ww["lppd_i"], gk)
    merged = bd.combine_stats(torch.from_numpy(rec))
    summ = bd.summarize(merged, 600, "log")
    if rank == 0:
        np.save(os.path.join(tmp, "merged.npy"), merged.raw)
        np.save(os.path.join(tmp, "summary.npy"), np.array([summ["elpd_loo"], summ["se"], summ["p_loo"],
                                                            summ["elpd_waic"], summ["waic_se"], summ["p_waic"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_and_balance():
    from pyloo_b200.distributed import shard_bounds

    for n, w in ((10, 3), (1_000_000, 8), (7, 8), (0, 2)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_merge_equals_single_shard(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    from oracle import psis_oracle as orc

    rng = np.random.default_rng(123)
    ll = -1.4 + rng.normal(size=(600, 37))
    ref = orc.loo_summary(ll, 1.0)
    wref = orc.waic_summary(ll)
    got = np.load(tmp_path / "summary.npy")
    np.testing.assert_allclose(got[0], ref["elpd_loo"], rtol=1e-12)
    np.testing.assert_allclose(got[1], ref["se"], rtol=1e-10)
    np.testing.assert_allclose(got[2], ref["p_loo"], rtol=1e-11)
    np.testing.assert_allclose(got[3], wref["elpd_waic"], rtol=1e-12)
    np.testing.assert_allclose(got[4], wref["se"], rtol=1e-10)
    np.testing.assert_allclose(got[5], wref["p_waic"], rtol=1e-12)
    merged = np.load(tmp_path / "merged.npy")
    assert merged[0] == 37

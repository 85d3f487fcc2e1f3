"""Observation-sharded PSIS-LOO across GPUs of one box (one process per GPU, ``torch.distributed``).

Observations are independent in the reference (pyloo/utils.py:171-176 applies the 1-D function per
index), so every rank runs the fused kernel on its own contiguous block of observations and the only
exchange is an all-gather of the fixed 32-double statistics record (NCCL over NVLink on GPUs, gloo in
the CPU tests), merged with Chan's update in rank order: totals and standard errors are independent
of the GPU count.  Pointwise outputs stay sharded.
"""

from __future__ import annotations

import numpy as np

from . import _native, engine

__all__ = ["shard_bounds", "combine_stats", "summarize", "loo_sharded"]


def shard_bounds(n_obs: int, world: int, rank: int):
    """Contiguous block ``[lo, hi)`` of the observation axis owned by ``rank`` (sizes differ by <= 1)."""
    base, extra = divmod(n_obs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def combine_stats(local_record, group=None):
    """All-gather per-rank records and merge them in rank order.  ``local_record`` is a 32-element
    float64 torch tensor (any device) or array; returns an :class:`engine.StatsRecord`."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return engine.StatsRecord(np.asarray(local_record.cpu() if hasattr(local_record, "cpu") else local_record))
    rec = local_record if isinstance(local_record, torch.Tensor) else torch.as_tensor(np.asarray(local_record))
    rec = rec.to(torch.float64).contiguous()
    if rec.is_cuda and dist.get_backend(group) != "nccl":
        rec = rec.cpu()   # gloo (the CPU tests) has no all_gather on CUDA tensors; the record is 256 bytes
    world = dist.get_world_size(group)
    gathered = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec, group=group)        # the single collective of the path
    return engine.stats_merge([g.cpu().numpy() for g in gathered])


def summarize(stats, n_samples: int, scale: str = "log") -> dict:
    """ELPDData scalar rows from a merged record (pyloo/loo.py:326-342, pyloo/waic.py:157-160)."""
    sv = {"log": 1, "negative_log": -1, "deviance": -2}[scale]
    n = stats.n
    var_e = stats.elpd_m2 / n * sv * sv
    var_w = stats.waic_m2 / n * sv * sv
    elpd = sv * stats.elpd_sum
    se = float((n * var_e) ** 0.5)
    return {
        "elpd_loo": elpd, "se": se, "p_loo": stats.lppd_sum - elpd / sv, "p_loo_se": float(np.sqrt(var_e)),
        "looic": -2 * elpd, "looic_se": 2 * se, "n_samples": n_samples, "n_data_points": int(n),
        "warning": bool(stats.k_gt_good > 0), "n_high_k": int(stats.k_gt_good),
        "elpd_waic": sv * stats.waic_sum, "waic_se": float((n * var_w) ** 0.5), "p_waic": stats.p_waic_sum,
        "waic_warning": bool(stats.var_gt_04 > 0), "scale": scale,
    }


def loo_sharded(ll_sn_shard, reff: float = 1.0, scale: str = "log", group=None):
    """Fused loo + waic on this rank's observation shard (``(S, N_local)`` CUDA tensor or host array),
    then the one exchange.  Returns ``(summary dict over ALL observations, local pointwise dict)``."""
    import torch

    if isinstance(ll_sn_shard, np.ndarray):
        ll_sn_shard = torch.from_numpy(np.ascontiguousarray(ll_sn_shard)).cuda(engine.current_device())
    res = engine.loo_cuda(ll_sn_shard, reff)
    rec = engine.stats_cuda(res)
    stats = combine_stats(rec, group)
    return summarize(stats, res["n_samples"], scale), res

"""Thin Python layer over the C ABI: NumPy (host) and torch (device-resident) entry points.

The numerics live in ``csrc/``; this module only computes the two scalars the reference derives in
Python (``M`` and ``cutoffmin``, pyloo/psis.py:89-90), normalises array layouts into
(pointer, strides) and turns return codes into exceptions.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _native

__all__ = [
    "CUTOFFMIN", "tail_length", "good_k_threshold", "psislw_host", "loo_host", "psislw_cuda",
    "loo_cuda", "stats_cuda", "stats_merge", "StatsRecord", "row_launch_info", "current_device",
    "workspace_for", "profile", "profile_read", "split_launch_info", "handover_reasons", "host_devices",
]

CUTOFFMIN = float(np.log(np.finfo(float).tiny))  # pyloo/psis.py:90


def tail_length(n_samples: int, reff: float) -> int:
    """``M`` with ``cutoff_ind = -M - 1``: the reference's exact Python-float expression
    (pyloo/psis.py:89, pyloo/base.py:139-141).  Never recomputed on the device."""
    return int(np.ceil(min(n_samples / 5.0, 3 * (n_samples / reff) ** 0.5)))


def good_k_threshold(n_samples: int) -> float:
    """pyloo/loo.py:249."""
    return float(min(1 - 1 / np.log10(n_samples), 0.7))


def current_device() -> int:
    """GPU this process drives: LOCAL_RANK under torchrun, else B2L_DEVICE, else 0."""
    for key in ("LOCAL_RANK", "B2L_DEVICE"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


_MGPU_MIN_BYTES = 1 << 30  # below this one GPU finishes before the other contexts exist


def host_devices(device=None, devices=None, nbytes: int = 0) -> list:
    """Devices a host-array call is spread over (observation shards, one pipeline per device).

    ``device`` (one index) or ``devices`` (a sequence) pin the choice.  Otherwise: a process that torchrun
    gave a LOCAL_RANK (or the user a B2L_DEVICE) drives that one GPU; ``B2L_DEVICES`` ("all" or "0,2,3") names
    a list; else every visible GPU once the input is at least 1 GiB, and GPU 0 for smaller inputs."""
    if device is not None:
        return [int(device)]
    if devices is not None:
        return [int(d) for d in devices]
    if "LOCAL_RANK" in os.environ or "B2L_DEVICE" in os.environ:
        return [current_device()]
    spec = os.environ.get("B2L_DEVICES", "").strip().lower()
    if spec and spec != "all":
        return [int(d) for d in spec.split(",") if d.strip()]
    if not spec and nbytes < _MGPU_MIN_BYTES:
        return [0]
    n = _native.load().b2l_device_count()
    return list(range(max(1, n)))


def _check_tail(S: int, M: int) -> None:
    if M + 1 > S:
        # the reference fails with IndexError at x[x_sort_ind[cutoff_ind]] (pyloo/psis.py:136)
        raise IndexError(f"index {-M - 1} is out of bounds for axis 0 with size {S}")


class StatsRecord:
    """Named view of the fixed 32-double shard record (include/psisloo_b200.h)."""

    FIELDS = {
        "n": 0, "elpd_mean": 1, "elpd_m2": 2, "elpd_sum": 3, "lppd_sum": 4, "p_waic_sum": 5,
        "waic_mean": 6, "waic_m2": 7, "waic_sum": 8, "k_gt_good": 9, "k_gt_1": 10, "k_inf": 11,
        "k_nan": 12, "var_gt_04": 13, "elpd_min": 14, "elpd_max": 15, "waic_min": 16,
        "waic_max": 17, "n_nan_in": 18, "n_pinf_in": 19, "n_ninf_in": 20, "n_fallback": 21,
        "elpd_nan": 22,
    }

    def __init__(self, raw):
        self.raw = np.asarray(raw, dtype=np.float64).reshape(_native.STATS_LEN)

    def __getattr__(self, name):
        try:
            return float(self.raw[StatsRecord.FIELDS[name]])
        except KeyError as err:
            raise AttributeError(name) from err

    def __repr__(self):
        return "StatsRecord(" + ", ".join(f"{k}={getattr(self, k):g}" for k in self.FIELDS) + ")"


def stats_merge(records) -> StatsRecord:
    """Chan merge of per-shard records in the given (rank) order -- host arithmetic in the library."""
    recs = np.ascontiguousarray(np.asarray([np.asarray(getattr(r, "raw", r)) for r in records],
                                           dtype=np.float64))
    out = np.empty(_native.STATS_LEN)
    lib = _native.load()
    _native.check(lib.b2l_stats_merge(recs.ctypes.data, recs.shape[0], out.ctypes.data))
    return StatsRecord(out)


# --------------------------------------------------------------------------------- host (NumPy)
def _elem_strides(a: np.ndarray):
    return tuple(int(s // a.itemsize) for s in a.strides)


def _as_strided_f64_2d(a) -> np.ndarray:
    """float64 2-D array with one unit, non-negative-stride axis (copy only if needed)."""
    a = np.asarray(a)
    if a.dtype != np.float64:
        a = a.astype(np.float64)
    if a.ndim != 2:
        raise ValueError("expected a 2-D array")
    st = _elem_strides(a)
    ok = all(s >= 0 for s in st) and all(s * a.itemsize == b for s, b in zip(st, a.strides))
    ok = ok and (st[0] == 1 or st[1] == 1 or a.shape[0] == 1 or a.shape[1] == 1) and a.size > 0
    # the non-unit stride must not alias rows (broadcast views)
    if ok and min(a.shape) > 1 and 0 in st:
        ok = False
    return a if ok else np.ascontiguousarray(a)


def psislw_host(lw_ns: np.ndarray, reff: float = 1.0, *, out=None, device=None, devices=None, chunk_obs: int = 0):
    """Batch PSIS on a HOST ``(N, S)`` array (samples on the last axis, any unit-stride layout).

    Returns ``(lw_out, k)``: C-contiguous ``(N, S)`` smoothed log weights and ``(N,)`` Pareto k.
    The input is never modified (pyloo/psis.py:78).  Observation shards go to the GPUs of
    :func:`host_devices`."""
    lib = _native.load()
    a = _as_strided_f64_2d(lw_ns)
    N, S = a.shape
    M = tail_length(S, reff)
    _check_tail(S, M)
    if out is None:
        out = np.empty((N, S), dtype=np.float64)
    k = np.empty(N, dtype=np.float64)
    if N == 0:
        return out, k
    sn, ss = _elem_strides(a)
    osn, oss = _elem_strides(out)
    devs = np.asarray(host_devices(device, devices, a.nbytes), dtype=np.int32)
    rc = lib.b2l_psislw_host_mgpu_f64(a.ctypes.data, S, N, ss, sn, M, CUTOFFMIN, out.ctypes.data, oss, osn,
                                      k.ctypes.data, devs.ctypes.data, int(devs.size), int(chunk_obs))
    _native.check(rc)
    return out, k


def loo_host(ll_sn: np.ndarray, reff: float = 1.0, *, waic_only: bool = False, device=None, devices=None,
             chunk_obs: int = 0):
    """Fused pointwise PSIS-LOO + WAIC on a HOST sample-major ``(S, N)`` log-likelihood
    (ArviZ ``(chain, draw, obs)`` flattened; a transposed ``(N, S)``-contiguous view also works).

    Returns ``dict(elpd_i, pareto_k, lppd_i, var_i, lppdw_i, stats=StatsRecord, M, good_k)``;
    ``elpd_i`` is on the log scale."""
    lib = _native.load()
    a = _as_strided_f64_2d(ll_sn)
    S, N = a.shape
    M = tail_length(S, reff)
    _check_tail(S, M)
    gk = good_k_threshold(S)
    outs = [np.empty(N, dtype=np.float64) for _ in range(5)]
    stats = np.zeros(_native.STATS_LEN)
    ss, sn = _elem_strides(a)
    devs = np.asarray(host_devices(device, devices, a.nbytes), dtype=np.int32)
    rc = lib.b2l_loo_host_mgpu_f64(a.ctypes.data, S, N, ss, sn, M, CUTOFFMIN,
                                   _native.FLAG_WAIC_ONLY if waic_only else 0, gk, *[o.ctypes.data for o in outs],
                                   stats.ctypes.data, devs.ctypes.data, int(devs.size), int(chunk_obs))
    _native.check(rc)
    return {"elpd_i": outs[0], "pareto_k": outs[1], "lppd_i": outs[2], "var_i": outs[3],
            "lppdw_i": outs[4], "stats": StatsRecord(stats), "M": M, "good_k": gk, "n_samples": S}


# --------------------------------------------------------------------------------- device (torch)
def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("pyloo_b200 needs a CUDA device and has no CPU fallback")
    return torch


def _stream_ptr(torch, device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _workspace(torch, lib, S, N, M, obs_fastest, device):
    need = ctypes.c_size_t(0)
    _native.check(lib.b2l_workspace_bytes(S, N, M, 1 if obs_fastest else 0, ctypes.byref(need)))
    return torch.empty(max(int(need.value), 256), dtype=torch.uint8, device=device)


def workspace_for(S: int, N: int, reff: float, obs_fastest: bool, device):
    """Device workspace (uint8 tensor) sized by the library for repeated calls on one problem shape."""
    torch = _torch()
    return _workspace(torch, _native.load(), S, N, tail_length(S, reff), obs_fastest, device)


def profile(enable: bool) -> None:
    """Per-kernel CUDA-event timing inside the library (benchmarks only; see include/psisloo_b200.h)."""
    _native.check(_native.load().b2l_profile(1 if enable else 0))


def profile_read() -> dict:
    """{kind: (milliseconds, launches)} accumulated since the last read (synchronises the events)."""
    ms = np.zeros(len(_native.PROF_KINDS))
    cnt = np.zeros(len(_native.PROF_KINDS), dtype=np.int64)
    _native.check(_native.load().b2l_profile_read(ms.ctypes.data, cnt.ctypes.data))
    return {k: (float(m), int(c)) for k, m, c in zip(_native.PROF_KINDS, ms, cnt)}


def handover_reasons(reset: bool = True) -> dict:
    """Diagnostics: observations handed from the split path to the general kernel, by reason."""
    names = {1: "nan_inf", 2: "range", 3: "threshold_retries", 4: "key_runs", 5: "order_check",
             6: "gpd_quantile", 7: "gpd_factor_overflow", 8: "gpd_product", 9: "gpd_profile", 10: "body_cancellation"}
    out = np.zeros(16, dtype=np.uint64)
    _native.check(_native.load().b2l_handover_reasons(out.ctypes.data, 1 if reset else 0))
    return {names[i]: int(out[i]) for i in names if out[i]}


def split_launch_info(S: int, M: int, mode: str = "psislw", n_rows: int = 1 << 30) -> dict:
    """Launch shape of the split (stream + tail kernel) path; needs a GPU (occupancy queries)."""
    info = np.zeros(16, dtype=np.int32)
    _native.check(_native.load().b2l_split_launch_info(S, M, 0 if mode == "psislw" else 1, n_rows, info.ctypes.data))
    keys = ("eligible", "stream_threads", "draws_per_thread", "tail_regs_per_lane", "candidate_cap", "q0",
            "row_buffers", "fused_apply", "stream_grid", "tail_grid", "stream_ctas_per_sm", "tail_ctas_per_sm",
            "stream_smem_bytes", "tail_smem_bytes", "obs_per_round", "stream_block")
    return {k: int(v) for k, v in zip(keys, info)}


def tile_shape_info(S: int, M: int) -> dict:
    """Plan of the cluster kernel that serves ``loo`` on the ``(chain, draw, obs)`` layout for ``S`` draws and a tail
    of ``M``: pure arithmetic in the library, no GPU needed.  ``eligible`` = 0: the shape takes the panel route."""
    info = np.zeros(16, dtype=np.int32)
    _native.check(_native.load().b2l_tile_shape_info(S, M, info.ctypes.data))
    keys = ("eligible", "obs_per_tile", "cluster_size", "n_chunks", "chunk_len", "draws_per_cta", "boxes_per_cta",
            "draws_per_box", "rank_tight", "rank_loose", "smem_bytes", "tail_regs_per_lane", "candidate_cap")
    return {k: int(v) for k, v in zip(keys, info)}


def psislw_cuda(lw, reff: float = 1.0, *, out=None, want_diag: bool = False, workspace=None):
    """PSIS on a device-resident float64 tensor ``(N, S)`` (any unit-stride layout); asynchronous
    on the current stream.  Returns ``(lw_out, k[, diag])`` tensors."""
    torch = _torch()
    lib = _native.load()
    if lw.dtype != torch.float64 or lw.dim() != 2 or not lw.is_cuda:
        raise ValueError("expected a 2-D float64 CUDA tensor")
    N, S = lw.shape
    M = tail_length(S, reff)
    _check_tail(S, M)
    sn, ss = lw.stride()
    if not (ss == 1 or sn == 1 or N == 1 or S == 1):
        lw = lw.contiguous()
        sn, ss = lw.stride()
    if out is None:
        out = torch.empty((N, S), dtype=torch.float64, device=lw.device)
    osn, oss = out.stride()
    k = torch.empty(N, dtype=torch.float64, device=lw.device)
    diag = torch.zeros((N, _native.DIAG_STRIDE), dtype=torch.float64, device=lw.device) if want_diag else None
    obs_fast = not (ss == 1 or S == 1) or not (oss == 1 or S == 1)
    ws = workspace if workspace is not None else _workspace(torch, lib, S, N, M, obs_fast, lw.device)
    with torch.cuda.device(lw.device):
        rc = lib.b2l_psislw_dev_f64(lw.data_ptr(), S, N, ss, sn, M, CUTOFFMIN, out.data_ptr(), oss, osn,
                                    k.data_ptr(), diag.data_ptr() if want_diag else None,
                                    ws.data_ptr(), ws.numel(), _stream_ptr(torch, lw.device))
    _native.check(rc)
    return (out, k, diag) if want_diag else (out, k)


def loo_cuda(ll_sn, reff: float = 1.0, *, want_diag: bool = False, workspace=None, counters=None,
             waic_only: bool = False, want_tail_idx: bool = False):
    """Fused pointwise PSIS-LOO + WAIC on a device-resident float64 ``(S, N)`` tensor (obs-fastest
    ArviZ layout, or a transposed view of a row-contiguous ``(N, S)`` tensor).  Asynchronous.
    Returns dict of device tensors ``elpd_i, pareto_k, lppd_i, var_i, lppdw_i, counters[, diag]``."""
    torch = _torch()
    lib = _native.load()
    if ll_sn.dtype != torch.float64 or ll_sn.dim() != 2 or not ll_sn.is_cuda:
        raise ValueError("expected a 2-D float64 CUDA tensor")
    S, N = ll_sn.shape
    M = tail_length(S, reff)
    _check_tail(S, M)
    ss, sn = ll_sn.stride()
    if not (ss == 1 or sn == 1 or N == 1 or S == 1):
        ll_sn = ll_sn.contiguous()
        ss, sn = ll_sn.stride()
    dev = ll_sn.device
    outs = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(5)]
    if counters is None:
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
    diag = torch.zeros((N, _native.DIAG_STRIDE), dtype=torch.float64, device=dev) if want_diag else None
    ws = workspace if workspace is not None else _workspace(torch, lib, S, N, M, not (ss == 1 or S == 1), dev)
    tail_idx = torch.full((N, M), -1, dtype=torch.int32, device=dev) if want_tail_idx else None
    with torch.cuda.device(dev):
        rc = lib.b2l_loo_dev_ex_f64(ll_sn.data_ptr(), S, N, ss, sn, M, CUTOFFMIN,
                                    _native.FLAG_WAIC_ONLY if waic_only else 0, *[o.data_ptr() for o in outs],
                                    counters.data_ptr(), diag.data_ptr() if want_diag else None,
                                    tail_idx.data_ptr() if want_tail_idx else None,
                                    ws.data_ptr(), ws.numel(), _stream_ptr(torch, dev))
    _native.check(rc)
    res = {"elpd_i": outs[0], "pareto_k": outs[1], "lppd_i": outs[2], "var_i": outs[3],
           "lppdw_i": outs[4], "counters": counters, "M": M, "n_samples": S, "workspace": ws}
    if want_diag:
        res["diag"] = diag
    if want_tail_idx:
        res["tail_idx"] = tail_idx   # (N, M) draw indices of each tail (psis.py:139-141), padded with -1
    return res


def stats_cuda(res: dict, good_k: float | None = None, workspace=None):
    """Shard statistics record (device tensor of 32 doubles) for the outputs of :func:`loo_cuda`."""
    torch = _torch()
    lib = _native.load()
    e = res["elpd_i"]
    N = e.numel()
    gk = good_k_threshold(res["n_samples"]) if good_k is None else good_k
    ws = workspace if workspace is not None else res.get("workspace")
    if ws is None or ws.numel() < 64 * 1024:
        ws = torch.empty(64 * 1024, dtype=torch.uint8, device=e.device)
    stats = torch.empty(_native.STATS_LEN, dtype=torch.float64, device=e.device)
    with torch.cuda.device(e.device):
        rc = lib.b2l_stats_dev_f64(e.data_ptr(), res["pareto_k"].data_ptr(), res["lppd_i"].data_ptr(),
                                   res["var_i"].data_ptr(), res["lppdw_i"].data_ptr(), N, gk,
                                   res["counters"].data_ptr(), stats.data_ptr(), ws.data_ptr(), ws.numel(),
                                   _stream_ptr(torch, e.device))
    _native.check(rc)
    return stats


def row_launch_info(S: int, M: int, mode: str = "psislw") -> dict:
    """Launch shape the library picks for a row kernel (needs a GPU: queries occupancy)."""
    lib = _native.load()
    vals = [ctypes.c_int32(0) for _ in range(5)]
    _native.check(lib.b2l_row_launch_info(S, M, 0 if mode == "psislw" else 1, *[ctypes.byref(v) for v in vals]))
    keys = ("grid", "block", "smem_bytes", "ctas_per_sm", "nbuf")
    return {k: int(v.value) for k, v in zip(keys, vals)}


# --------------------------------------------------------------------------------- SIS / TIS / e_loo
_IS_METHODS = {"sis": _native.IS_SIS, "tis": _native.IS_TIS}
_CHUNK_BYTES = 1 << 30  # host arrays are staged through the GPU in slabs of about this size


def _method_code(method) -> int:
    name = getattr(method, "value", method)
    try:
        return _IS_METHODS[str(name).lower()]
    except KeyError:
        raise ValueError(f"method must be 'sis' or 'tis' on this entry point, not {name!r}") from None


def islw_cuda(lw, method, *, out=None):
    """SIS / TIS normalised log weights of a device-resident float64 ``(N, S)`` tensor with contiguous
    rows (pyloo/sis.py:86-106, pyloo/tis.py:91-120).  Asynchronous.  Returns ``(lw_out, ess)``."""
    torch = _torch()
    lib = _native.load()
    if lw.dtype != torch.float64 or lw.dim() != 2 or not lw.is_cuda:
        raise ValueError("expected a 2-D float64 CUDA tensor")
    N, S = lw.shape
    if S > 1 and lw.stride(1) != 1:
        lw = lw.contiguous()
    if out is None:
        out = torch.empty((N, S), dtype=torch.float64, device=lw.device)
    ess = torch.empty(N, dtype=torch.float64, device=lw.device)
    with torch.cuda.device(lw.device):
        rc = lib.b2l_islw_dev_f64(lw.data_ptr(), S, N, lw.stride(0), _method_code(method), out.data_ptr(),
                                  out.stride(0), ess.data_ptr(), _stream_ptr(torch, lw.device))
    _native.check(rc)
    return out, ess


def loo_is_cuda(ll_sn, method, *, counters=None, workspace=None):
    """Pointwise LOO with SIS / TIS weights on a device-resident float64 ``(S, N)`` log-likelihood
    (obs-fastest ArviZ layout or a transposed row-contiguous view).  Asynchronous.
    Returns dict of device tensors ``elpd_i, ess_i, lppd_i, counters``."""
    torch = _torch()
    lib = _native.load()
    if ll_sn.dtype != torch.float64 or ll_sn.dim() != 2 or not ll_sn.is_cuda:
        raise ValueError("expected a 2-D float64 CUDA tensor")
    S, N = ll_sn.shape
    ss, sn = ll_sn.stride()
    if not (ss == 1 or sn == 1 or N == 1 or S == 1):
        ll_sn = ll_sn.contiguous()
        ss, sn = ll_sn.stride()
    dev = ll_sn.device
    outs = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(3)]
    if counters is None:
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
    if workspace is None:
        need = ctypes.c_size_t(0)
        _native.check(lib.b2l_is_workspace_bytes(S, N, 0 if (ss == 1 or S == 1) else 1, ctypes.byref(need)))
        workspace = torch.empty(max(int(need.value), 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.b2l_loo_is_dev_f64(ll_sn.data_ptr(), S, N, ss, sn, _method_code(method),
                                    *[o.data_ptr() for o in outs], counters.data_ptr(), workspace.data_ptr(),
                                    workspace.numel(), _stream_ptr(torch, dev))
    _native.check(rc)
    return {"elpd_i": outs[0], "ess_i": outs[1], "lppd_i": outs[2], "counters": counters, "n_samples": S}


def eloo_cuda(x, lw, lr=None, kind: str = "mean", tail_len: int = 20, *, workspace=None):
    """Weighted expectation of ``x`` under log weights ``lw`` with the function-specific Pareto k
    (pyloo/e_loo.py:429-463, :328-390) for device-resident float64 ``(N, S)`` tensors with contiguous rows.
    ``kind``: ``mean`` / ``variance`` / ``sd`` / ``none`` (k of the ratios only; ``x`` may be None).
    Asynchronous.  Returns ``(value or None, khat)``."""
    torch = _torch()
    lib = _native.load()
    code = _native.ELOO_TYPES[kind]

    def rows(t):
        if t is None:
            return None
        if t.dtype != torch.float64 or t.dim() != 2 or not t.is_cuda:
            raise ValueError("expected 2-D float64 CUDA tensors")
        return t if (t.shape[1] == 1 or t.stride(1) == 1) else t.contiguous()

    x, lw, lr = rows(x if code != 3 else None), rows(lw), rows(lr)
    N, S = lw.shape
    for t in (x, lr):
        if t is not None and tuple(t.shape) != (N, S):
            raise ValueError("x, log_weights and log_ratios must have the same shape")
    dev = lw.device
    value = torch.empty(N, dtype=torch.float64, device=dev) if code != 3 else None
    khat = torch.empty(N, dtype=torch.float64, device=dev)
    if workspace is None:
        need = ctypes.c_size_t(0)
        _native.check(lib.b2l_eloo_workspace_bytes(S, N, 0 if lr is None else 1, code, ctypes.byref(need)))
        workspace = torch.empty(max(int(need.value), 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.b2l_eloo_dev_f64(x.data_ptr() if x is not None else None, x.stride(0) if x is not None else 0,
                                  lw.data_ptr(), lw.stride(0), lr.data_ptr() if lr is not None else None,
                                  lr.stride(0) if lr is not None else 0, S, N, code, int(tail_len),
                                  value.data_ptr() if value is not None else None, khat.data_ptr(),
                                  workspace.data_ptr(), workspace.numel(), _stream_ptr(torch, dev))
    _native.check(rc)
    return value, khat


def eloo_quantile_cuda(x, lw, probs):
    """Weighted quantiles (pyloo/e_loo.py:466-554) of device-resident float64 ``(N, S)`` tensors with
    contiguous rows; ``probs`` is a host sequence in (0, 1).  Asynchronous.  Returns an ``(N, len(probs))`` tensor."""
    torch = _torch()
    lib = _native.load()
    for t in (x, lw):
        if t.dtype != torch.float64 or t.dim() != 2 or not t.is_cuda:
            raise ValueError("expected 2-D float64 CUDA tensors")
    if x.shape != lw.shape:
        raise ValueError("x and log_weights must have the same shape")
    x = x if (x.shape[1] == 1 or x.stride(1) == 1) else x.contiguous()
    lw = lw if (lw.shape[1] == 1 or lw.stride(1) == 1) else lw.contiguous()
    N, S = x.shape
    pr = np.ascontiguousarray(np.atleast_1d(np.asarray(probs, dtype=np.float64)))
    out = torch.empty((N, pr.size), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.b2l_eloo_quantile_dev_f64(x.data_ptr(), x.stride(0), lw.data_ptr(), lw.stride(0), S, N,
                                           pr.ctypes.data, int(pr.size), out.data_ptr(),
                                           _stream_ptr(torch, x.device))
    _native.check(rc)
    return out


def eloo_quantile_host(x_ns, lw_ns, probs, *, device=None):
    """Weighted quantiles on HOST ``(N, S)`` arrays.  Returns an ``(N, len(probs))`` array."""
    torch = _torch()
    x = np.asarray(x_ns, dtype=np.float64)
    lw = np.asarray(lw_ns, dtype=np.float64)
    N, S = lw.shape
    pr = np.atleast_1d(np.asarray(probs, dtype=np.float64))
    out = np.empty((N, pr.size), dtype=np.float64)
    dev = _dev(device)
    for i0, i1 in _slabs(N, 16 * S):
        d_x = torch.from_numpy(np.ascontiguousarray(x[i0:i1])).to(dev)
        d_lw = torch.from_numpy(np.ascontiguousarray(lw[i0:i1])).to(dev)
        out[i0:i1] = eloo_quantile_cuda(d_x, d_lw, pr).cpu().numpy()
    return out


def _slabs(N: int, row_bytes: int):
    step = max(1, min(N, _CHUNK_BYTES // max(row_bytes, 1)))
    for i0 in range(0, N, step):
        yield i0, min(N, i0 + step)


def _dev(device):
    torch = _torch()
    return torch.device("cuda", current_device() if device is None else int(device))


def islw_host(lw_ns: np.ndarray, method, *, device=None):
    """SIS / TIS on a HOST ``(N, S)`` array.  Returns ``(lw_out (N, S), ess (N,))``; input untouched."""
    torch = _torch()
    a = np.asarray(lw_ns, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("expected a 2-D array")
    N, S = a.shape
    out = np.empty((N, S), dtype=np.float64)
    ess = np.empty(N, dtype=np.float64)
    dev = _dev(device)
    for i0, i1 in _slabs(N, 16 * S):
        d_in = torch.from_numpy(np.ascontiguousarray(a[i0:i1])).to(dev)
        d_out, d_ess = islw_cuda(d_in, method)
        out[i0:i1] = d_out.cpu().numpy()
        ess[i0:i1] = d_ess.cpu().numpy()
    return out, ess


def loo_is_host(ll_sn: np.ndarray, method, *, device=None):
    """Pointwise SIS / TIS LOO on a HOST sample-major ``(S, N)`` log-likelihood.
    Returns ``dict(elpd_i, ess_i, lppd_i, n_nan_in)`` of NumPy arrays (log scale)."""
    torch = _torch()
    a = _as_strided_f64_2d(ll_sn)
    S, N = a.shape
    res = {k: np.empty(N, dtype=np.float64) for k in ("elpd_i", "ess_i", "lppd_i")}
    dev = _dev(device)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    for i0, i1 in _slabs(N, 8 * S):
        d_in = torch.from_numpy(np.ascontiguousarray(a[:, i0:i1])).to(dev)
        r = loo_is_cuda(d_in, method, counters=counters)
        for k in res:
            res[k][i0:i1] = r[k].cpu().numpy()
    res["n_nan_in"] = int(counters[0].item())
    res["n_samples"] = S
    return res


def eloo_host(x_ns, lw_ns, lr_ns=None, kind: str = "mean", tail_len: int = 20, *, device=None):
    """Weighted expectation + Pareto k on HOST ``(N, S)`` arrays.  Returns ``(value or None, khat)``."""
    torch = _torch()
    lw = np.asarray(lw_ns, dtype=np.float64)
    if lw.ndim != 2:
        raise ValueError("expected 2-D arrays")
    N, S = lw.shape
    x = None if (x_ns is None or kind == "none") else np.asarray(x_ns, dtype=np.float64)
    lr = None if lr_ns is None else np.asarray(lr_ns, dtype=np.float64)
    value = np.empty(N, dtype=np.float64) if x is not None else None
    khat = np.empty(N, dtype=np.float64)
    dev = _dev(device)

    def up(arr, i0, i1):
        return None if arr is None else torch.from_numpy(np.ascontiguousarray(arr[i0:i1])).to(dev)

    for i0, i1 in _slabs(N, 24 * S):
        v, k = eloo_cuda(up(x, i0, i1), up(lw, i0, i1), up(lr, i0, i1), kind, tail_len)
        if value is not None:
            value[i0:i1] = v.cpu().numpy()
        khat[i0:i1] = k.cpu().numpy()
    return value, khat


def psis_expectation_host(x_ns, lr_ns, reff: float = 1.0, kind: str = "mean", tail_len: int = 20, *, device=None):
    """PSIS weights of the raw log ratios ``lr_ns`` chained with the weighted expectation of ``x_ns`` under
    them, both ``(N, S)`` HOST arrays -- what ``loo_predictive_metric`` and ``loo_score`` do through
    ``psislw`` + ``e_loo`` (pyloo/loo_predictive_metric.py:208-218, pyloo/loo_score.py:227-237, :311-321).
    The smoothed ``(N, S)`` weights never leave the device.
    Returns ``(value (N,), khat (N,), pareto_k (N,))``."""
    torch = _torch()
    lr = np.asarray(lr_ns, dtype=np.float64)
    x = np.asarray(x_ns, dtype=np.float64)
    if lr.ndim != 2 or x.shape != lr.shape:
        raise ValueError("x and log ratios must be 2-D arrays of the same shape")
    N, S = lr.shape
    M = tail_length(S, reff)
    _check_tail(S, M)
    value, khat, pk = (np.empty(N, dtype=np.float64) for _ in range(3))
    dev = _dev(device)
    ws = lw = None
    for i0, i1 in _slabs(N, 40 * S):
        d_lr = torch.from_numpy(np.ascontiguousarray(lr[i0:i1])).to(dev)
        d_x = torch.from_numpy(np.ascontiguousarray(x[i0:i1])).to(dev)
        if ws is None:
            ws = workspace_for(S, i1 - i0, reff, False, dev)
            lw = torch.empty((i1 - i0, S), dtype=torch.float64, device=dev)
        d_lw, d_k = psislw_cuda(d_lr, reff, out=lw[: i1 - i0], workspace=ws)
        d_v, d_kh = eloo_cuda(d_x, d_lw, d_lr, kind, tail_len)
        value[i0:i1] = d_v.cpu().numpy()
        khat[i0:i1] = d_kh.cpu().numpy()
        pk[i0:i1] = d_k.cpu().numpy()
    return value, khat, pk


def group_loo_host(ll_sn, group_index, n_groups: int, reff: float = 1.0, method: str = "psis", *, device=None):
    """Leave-one-group-out pointwise pass on a HOST sample-major ``(S, N)`` log-likelihood
    (pyloo/loo_group.py:188-285): per-group sums of the log-likelihood on the device
    (``b2l_group_sum_dev_f64``), then the ordinary LOO pass with one "observation" per group.
    ``group_index``: int array of length N with values in ``[0, n_groups)``.
    Returns the dict of :func:`loo_host` (PSIS) or :func:`loo_is_host` (SIS / TIS) over the groups."""
    torch = _torch()
    lib = _native.load()
    a = np.asarray(ll_sn, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("expected a 2-D array")
    S, N = a.shape
    gidx = np.asarray(group_index, dtype=np.int64)
    if gidx.shape != (N,) or gidx.min() < 0 or gidx.max() >= n_groups:
        raise ValueError("group_index must hold one group number in [0, n_groups) per observation")
    G = int(n_groups)
    dev = _dev(device)
    members = torch.from_numpy(np.argsort(gidx, kind="stable").astype(np.int32)).to(dev)
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(gidx, minlength=G))]).astype(np.int32)).to(dev)
    sums = torch.empty((G, S), dtype=torch.float64, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    step = max(1, min(S, _CHUNK_BYTES // max(8 * N, 1)))
    with torch.cuda.device(dev):
        for s0 in range(0, S, step):
            s1 = min(S, s0 + step)
            slab = torch.from_numpy(np.ascontiguousarray(a[s0:s1])).to(dev)  # (s1 - s0, N), observations fastest
            rc = lib.b2l_group_sum_dev_f64(slab.data_ptr(), s1 - s0, N, N, 1, members.data_ptr(), offsets.data_ptr(),
                                           G, sums.data_ptr() + 8 * s0, S, counters.data_ptr(),
                                           _stream_ptr(torch, dev))
            _native.check(rc)
            torch.cuda.current_stream(dev).synchronize()  # the slab is released before the next upload
    n_nan = int(counters[0].item())
    if _method_name(method) == "psis":
        res = loo_cuda(sums.t(), reff)
        st = StatsRecord(stats_cuda(res).cpu().numpy())
        out = {k: res[k].cpu().numpy() for k in ("elpd_i", "pareto_k", "lppd_i")}
        out.update(stats=st, good_k=good_k_threshold(S), n_samples=S, n_nan_in=n_nan)
        return out
    res = loo_is_cuda(sums.t(), method)
    out = {k: res[k].cpu().numpy() for k in ("elpd_i", "ess_i", "lppd_i")}
    out.update(n_samples=S, n_nan_in=n_nan)
    return out


def _method_name(method) -> str:
    return str(getattr(method, "value", method)).lower()


def loo_subset_cuda(ll_sn, obs_index, reff: float = 1.0, *, waic_only: bool = False):
    """The PSIS stage of ``loo_subsample`` (pyloo/loo_subsample.py:330, :371-383) on a device-resident
    ``(S, N)`` log-likelihood: the observations ``obs_index`` (int64 tensor or sequence, repeats allowed) are
    gathered into contiguous rows on the device and go through the fused LOO pass.  Asynchronous.
    Returns the dict of :func:`loo_cuda` with one entry per element of ``obs_index``."""
    torch = _torch()
    lib = _native.load()
    if ll_sn.dtype != torch.float64 or ll_sn.dim() != 2 or not ll_sn.is_cuda:
        raise ValueError("expected a 2-D float64 CUDA tensor")
    S, N = ll_sn.shape
    idx = torch.as_tensor(obs_index, dtype=torch.int64, device=ll_sn.device).contiguous()
    if idx.dim() != 1:
        raise ValueError("obs_index must be one-dimensional")
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= N):
        raise IndexError("obs_index out of range")
    m = idx.numel()
    rows = torch.empty((m, S), dtype=torch.float64, device=ll_sn.device)
    ss, sn = ll_sn.stride()
    with torch.cuda.device(ll_sn.device):
        rc = lib.b2l_gather_rows_dev_f64(ll_sn.data_ptr(), S, N, ss, sn, idx.data_ptr(), m, rows.data_ptr(), S,
                                         _stream_ptr(torch, ll_sn.device))
    _native.check(rc)
    return loo_cuda(rows.t(), reff, waic_only=waic_only)

"""Build / load libpsisloo_b200.so and declare its C ABI (include/psisloo_b200.h) for ctypes.

There is NO CPU fallback: if the shared library is missing or no CUDA device is visible the
compute entry points raise ``RuntimeError``.
"""

from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "lib", "libpsisloo_b200.so")
SRCS = [os.path.join(_PKG, "csrc", n) for n in ("psisloo_b200.cu", "b2l_split_stream.cu", "b2l_split_tail.cu", "b2l_is.cu", "b2l_tile.cu")]
SRC = SRCS[0]
HEADERS = [
    os.path.join(_PKG, "csrc", "b2l_common.cuh"),
    os.path.join(_PKG, "csrc", "b2l_row_kernel.cuh"),
    os.path.join(_PKG, "csrc", "b2l_split.cuh"),
    os.path.join(_PKG, "csrc", "b2l_split_host.h"),
    os.path.join(_PKG, "csrc", "b2l_is_host.h"),
    os.path.join(_PKG, "csrc", "b2l_tile_host.h"),
    os.path.join(_ROOT, "include", "psisloo_b200.h"),
]

STATS_LEN = 32
FLAG_WAIC_ONLY = 1
DIAG_STRIDE = 8
E_NODEVICE = -4

EXPORTS = [
    "b2l_version", "b2l_last_error", "b2l_device_count", "b2l_workspace_bytes",
    "b2l_psislw_dev_f64", "b2l_loo_dev_f64", "b2l_stats_dev_f64", "b2l_stats_merge",
    "b2l_psislw_host_f64", "b2l_loo_host_f64", "b2l_row_launch_info", "b2l_profile", "b2l_profile_read",
    "b2l_split_launch_info", "b2l_tile_shape_info", "b2l_handover_reasons",
    "b2l_islw_dev_f64", "b2l_is_workspace_bytes", "b2l_loo_is_dev_f64", "b2l_eloo_workspace_bytes",
    "b2l_eloo_dev_f64", "b2l_eloo_quantile_dev_f64", "b2l_group_sum_dev_f64", "b2l_gather_rows_dev_f64",
    "b2l_loo_host_mgpu_f64", "b2l_psislw_host_mgpu_f64", "b2l_loo_dev_ex_f64",
]
PROF_KINDS = ("stream", "tail", "apply", "row", "transpose", "stats", "is", "eloo")
IS_SIS, IS_TIS = 1, 2
ELOO_TYPES = {"mean": 0, "variance": 1, "sd": 2, "none": 3}

_lock = threading.Lock()
_lib = None


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libpsisloo_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in [*SRCS, *HEADERS])


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into an in-tree shared library."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    objdir = os.path.join(_PKG, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
             "-Xcompiler", "-fPIC"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    # one nvcc per translation unit, in parallel (the kernel templates dominate the build time)
    procs, objs = [], []
    for src in SRCS:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([_nvcc(), *flags, "-c", "-o", obj, src],
                                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    logs = []
    for src, pr in procs:
        out, err = pr.communicate()
        logs.append(err)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}\n{err}")
    res = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


def _declare(lib) -> None:
    c = ctypes
    i32, i64, f64, vp, sz, u32 = c.c_int32, c.c_int64, c.c_double, c.c_void_p, c.c_size_t, c.c_uint32
    lib.b2l_version.restype = c.c_int
    lib.b2l_version.argtypes = []
    lib.b2l_last_error.restype = c.c_char_p
    lib.b2l_last_error.argtypes = []
    lib.b2l_device_count.restype = c.c_int
    lib.b2l_device_count.argtypes = []
    lib.b2l_workspace_bytes.restype = c.c_int
    lib.b2l_workspace_bytes.argtypes = [i64, i64, i32, i32, c.POINTER(sz)]
    lib.b2l_psislw_dev_f64.restype = c.c_int
    lib.b2l_psislw_dev_f64.argtypes = [vp, i64, i64, i64, i64, i32, f64, vp, i64, i64, vp, vp, vp, sz, vp]
    lib.b2l_loo_dev_f64.restype = c.c_int
    lib.b2l_loo_dev_f64.argtypes = [vp, i64, i64, i64, i64, i32, f64, u32, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.b2l_loo_dev_ex_f64.restype = c.c_int
    lib.b2l_loo_dev_ex_f64.argtypes = [vp, i64, i64, i64, i64, i32, f64, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.b2l_stats_dev_f64.restype = c.c_int
    lib.b2l_stats_dev_f64.argtypes = [vp, vp, vp, vp, vp, i64, f64, vp, vp, vp, sz, vp]
    lib.b2l_stats_merge.restype = c.c_int
    lib.b2l_stats_merge.argtypes = [vp, i32, vp]
    lib.b2l_psislw_host_f64.restype = c.c_int
    lib.b2l_psislw_host_f64.argtypes = [vp, i64, i64, i64, i64, i32, f64, vp, i64, i64, vp, i32, i64]
    lib.b2l_loo_host_f64.restype = c.c_int
    lib.b2l_loo_host_f64.argtypes = [vp, i64, i64, i64, i64, i32, f64, u32, f64, vp, vp, vp, vp, vp, vp, i32, i64]
    lib.b2l_loo_host_mgpu_f64.restype = c.c_int
    lib.b2l_loo_host_mgpu_f64.argtypes = [vp, i64, i64, i64, i64, i32, f64, u32, f64, vp, vp, vp, vp, vp, vp, vp, i32, i64]
    lib.b2l_psislw_host_mgpu_f64.restype = c.c_int
    lib.b2l_psislw_host_mgpu_f64.argtypes = [vp, i64, i64, i64, i64, i32, f64, vp, i64, i64, vp, vp, i32, i64]
    lib.b2l_profile.restype = c.c_int
    lib.b2l_profile.argtypes = [i32]
    lib.b2l_profile_read.restype = c.c_int
    lib.b2l_profile_read.argtypes = [vp, vp]
    lib.b2l_handover_reasons.restype = c.c_int
    lib.b2l_handover_reasons.argtypes = [vp, i32]
    lib.b2l_split_launch_info.restype = c.c_int
    lib.b2l_split_launch_info.argtypes = [i64, i32, i32, i64, vp]
    lib.b2l_tile_shape_info.restype = c.c_int
    lib.b2l_tile_shape_info.argtypes = [i64, i32, vp]
    lib.b2l_islw_dev_f64.restype = c.c_int
    lib.b2l_islw_dev_f64.argtypes = [vp, i64, i64, i64, i32, vp, i64, vp, vp]
    lib.b2l_is_workspace_bytes.restype = c.c_int
    lib.b2l_is_workspace_bytes.argtypes = [i64, i64, i32, c.POINTER(sz)]
    lib.b2l_loo_is_dev_f64.restype = c.c_int
    lib.b2l_loo_is_dev_f64.argtypes = [vp, i64, i64, i64, i64, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.b2l_eloo_workspace_bytes.restype = c.c_int
    lib.b2l_eloo_workspace_bytes.argtypes = [i64, i64, i32, i32, c.POINTER(sz)]
    lib.b2l_eloo_dev_f64.restype = c.c_int
    lib.b2l_eloo_dev_f64.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, i32, i32, vp, vp, vp, sz, vp]
    lib.b2l_eloo_quantile_dev_f64.restype = c.c_int
    lib.b2l_eloo_quantile_dev_f64.argtypes = [vp, i64, vp, i64, i64, i64, vp, i32, vp, vp]
    lib.b2l_group_sum_dev_f64.restype = c.c_int
    lib.b2l_group_sum_dev_f64.argtypes = [vp, i64, i64, i64, i64, vp, vp, i32, vp, i64, vp, vp]
    lib.b2l_gather_rows_dev_f64.restype = c.c_int
    lib.b2l_gather_rows_dev_f64.argtypes = [vp, i64, i64, i64, i64, vp, i64, vp, i64, vp]
    lib.b2l_row_launch_info.restype = c.c_int
    lib.b2l_row_launch_info.argtypes = [i64, i32, i32] + [c.POINTER(i32)] * 5


def load():
    """Return the loaded library (building it first if the sources are newer)."""
    global _lib
    with _lock:
        if _lib is None:
            if needs_build():
                build()
            lib = ctypes.CDLL(LIB_PATH)
            _declare(lib)
            _lib = lib
        return _lib


def check(rc: int) -> None:
    """Translate a C-ABI return code into a Python exception (never silently continue)."""
    if rc == 0:
        return
    msg = load().b2l_last_error().decode("utf-8", "replace")
    if rc == E_NODEVICE:
        raise RuntimeError(f"pyloo_b200 needs a CUDA device and has no CPU fallback: {msg}")
    if rc == -1:
        raise ValueError(msg)
    if rc == -2:
        raise NotImplementedError(msg)
    raise RuntimeError(f"libpsisloo_b200 error {rc}: {msg}")

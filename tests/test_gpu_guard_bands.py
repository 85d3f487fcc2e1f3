"""Out-of-bounds writes, found the plain way (the pool has no compute-sanitizer): every buffer handed to the C ABI --
outputs, diagnostics, tail indices, counters, the workspace -- is carved out of one allocation between 4 KB canary
bands; after the call every band must be untouched, the input unchanged, and a second call must give the same bits."""
import ctypes

import numpy as np
import pytest

from b2l_testutil import has_cuda

pytestmark = pytest.mark.gpu

if has_cuda():
    import torch
    from pyloo_b200 import _native, engine

BAND = 4096
CANARY = 0xA5


class Arena:
    def __init__(self, sizes):
        self.offsets, total = [], BAND
        for nbytes in sizes:
            self.offsets.append(total)
            total += (nbytes + 255) // 256 * 256 + BAND
        self.sizes = list(sizes)
        self.buf = torch.full((total,), CANARY, dtype=torch.uint8, device="cuda")

    def ptr(self, i):
        return self.buf.data_ptr() + self.offsets[i]

    def view(self, i, dtype, shape):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.buf[self.offsets[i]:self.offsets[i] + n].view(dtype).view(shape)

    def bands_intact(self):
        keep = torch.ones(self.buf.numel(), dtype=torch.bool, device="cuda")
        for off, nbytes in zip(self.offsets, self.sizes):
            keep[off:off + nbytes] = False
        return bool((self.buf[keep] == CANARY).all())


@pytest.mark.parametrize("S,N,reff", [(4000, 70, 1.0), (16000, 18, 1.0), (1000, 20, 1.0), (4000, 22, 0.2), (4000, 24, 0.1),
                                      (8000, 10, 0.25), (2000, 7, 1.0), (4000, 1, 1.0)])
def test_loo_entry_point_stays_inside_its_buffers(S, N, reff):
    lib = _native.load()
    rng = np.random.default_rng(S + N)
    ll_h = -1.4 + rng.normal(size=(S, N))
    ll_h[5 % S, 0] = np.nan                      # one column for the general kernel as well
    M = engine.tail_length(S, reff)
    need = ctypes.c_size_t(0)
    _native.check(lib.b2l_workspace_bytes(S, N, M, 1, ctypes.byref(need)))
    sizes = [S * N * 8] + [N * 8] * 5 + [4 * 8, N * _native.DIAG_STRIDE * 8, N * M * 4, int(need.value)]
    ar = Arena(sizes)
    ll = ar.view(0, torch.float64, (S, N))
    ll.copy_(torch.from_numpy(ll_h))
    ar.view(6, torch.int64, (4,)).zero_()
    results = []
    for _ in range(2):
        rc = lib.b2l_loo_dev_ex_f64(ar.ptr(0), S, N, N, 1, M, engine.CUTOFFMIN, 0, *[ar.ptr(i) for i in range(1, 6)],
                                    ar.ptr(6), ar.ptr(7), ar.ptr(8), ar.ptr(9), sizes[9], None)
        _native.check(rc)
        torch.cuda.synchronize()
        results.append([ar.view(i, torch.float64, (N,)).clone() for i in range(1, 6)])
    assert ar.bands_intact()
    assert torch.equal(torch.nan_to_num(ll, nan=7.0), torch.nan_to_num(torch.from_numpy(ll_h).cuda(), nan=7.0))
    for a, b in zip(*results):
        assert torch.equal(torch.nan_to_num(a, nan=7.0), torch.nan_to_num(b, nan=7.0))
    tail = ar.view(8, torch.int32, (N, M))
    assert int(tail.max()) < S and int(tail.min()) >= -1


@pytest.mark.parametrize("S,N,reff", [(4000, 40, 0.9), (600, 33, 1.0), (4000, 24, 0.1), (130, 5, 1.0)])
def test_psislw_entry_point_stays_inside_its_buffers(S, N, reff):
    lib = _native.load()
    rng = np.random.default_rng(S * 3 + N)
    lw_h = 1.5 * rng.normal(size=(N, S))
    M = engine.tail_length(S, reff)
    need = ctypes.c_size_t(0)
    _native.check(lib.b2l_workspace_bytes(S, N, M, 0, ctypes.byref(need)))
    sizes = [N * S * 8, N * S * 8, N * 8, N * _native.DIAG_STRIDE * 8, int(need.value)]
    ar = Arena(sizes)
    ar.view(0, torch.float64, (N, S)).copy_(torch.from_numpy(lw_h))
    rc = lib.b2l_psislw_dev_f64(ar.ptr(0), S, N, 1, S, M, engine.CUTOFFMIN, ar.ptr(1), 1, S, ar.ptr(2), ar.ptr(3),
                                ar.ptr(4), sizes[4], None)
    _native.check(rc)
    torch.cuda.synchronize()
    assert ar.bands_intact()
    assert torch.equal(ar.view(0, torch.float64, (N, S)), torch.from_numpy(lw_h).cuda())   # psis.py:78: input untouched
    out = ar.view(1, torch.float64, (N, S))
    assert bool(torch.isfinite(out).all())
    np.testing.assert_allclose(torch.logsumexp(out, dim=1).cpu().numpy(), 0.0, atol=1e-12)

"""GPU parity tests of the tile path: ``pl.loo`` on the ArviZ (chain, draw, obs) layout, i.e. the
observation-fastest ``(S, N)`` matrix read where it lies by the cluster kernel of ``csrc/b2l_tile.cu``
(2-D TMA tiles, no transposed panels) + the tail kernel, against the CPU oracle on the same seeded inputs.

Tolerances as in test_gpu_parity.py: elpd_i, lppd_i, var_i, Pareto k within 1e-10 relative (values that can
sit at zero get an absolute floor), the cutoff value and the tail count -- hence the tail index set --
bit-exact (pyloo/psis.py:135-141)."""

import os

import numpy as np
import pytest

from b2l_testutil import has_cuda

pytestmark = pytest.mark.gpu

if has_cuda():
    import torch
    from pyloo_b200 import engine

from oracle import psis_oracle as orc

RTOL = 1e-10


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol, equal_nan=True)


def same_special(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(np.isposinf(a), np.isposinf(b))
    assert np.array_equal(np.isneginf(a), np.isneginf(b))


def gpu_loo(ll_sn, reff, **kw):
    res = engine.loo_cuda(torch.from_numpy(np.ascontiguousarray(ll_sn)).cuda(), reff, **kw)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items() if k != "workspace"}


def oracle_tail(ll_sn, M):
    """(cutoff value, tail count) per observation of r = -ll, as pyloo/psis.py:134-141 defines them."""
    cut, cnt = [], []
    for col in ll_sn.T:
        z = -col - (-col).max()
        c = max(np.sort(z)[-M - 1], orc.CUTOFFMIN)
        cut.append(c)
        cnt.append(int((z > c).sum()))
    return np.array(cut), np.array(cnt)


def check_against_oracle(ll, reff, r, check_tail=True):
    with np.errstate(all="ignore"):
        pw = orc.loo_pointwise(ll, reff)
        ww = orc.waic_pointwise(ll)
    for key, ref in (("elpd_i", pw["elpd_i"]), ("pareto_k", pw["pareto_k"]), ("lppd_i", pw["lppd_i"]),
                     ("var_i", ww["var_i"]), ("lppdw_i", ww["lppd_i"])):
        same_special(r[key], ref)
        close(r[key], ref, atol=1e-13)   # lppd_i and k can sit arbitrarily close to 0
    if check_tail:
        cut, cnt = oracle_tail(ll, engine.tail_length(ll.shape[0], reff))
        assert np.array_equal(r["diag"][:, 1], cut)          # bit-exact cutoff => bit-exact tail index set
        assert np.array_equal(r["diag"][:, 2].astype(int), cnt)


@pytest.mark.parametrize("tw", ["8", "16"])
@pytest.mark.parametrize("S,N,reff", [(4000, 702, 1.0), (4000, 64, 0.9), (2000, 250, 0.7), (1024, 40, 1.0),
                                      (3000, 18, 1.0), (4096, 34, 1.0), (1500, 6, 0.5)])
def test_tile_path_vs_oracle(S, N, reff, tw, monkeypatch):
    """Ragged shapes: partial last tile (N not a multiple of the tile width), draw counts that do not divide by
    the cluster size or by 16 draw slots, the longest supported CTA share (S = 4096)."""
    monkeypatch.setenv("B2L_TILE_W", tw)
    rng = np.random.default_rng(S * 7 + N)
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.3, 2.0, size=(1, N))
    engine.handover_reasons()
    r = gpu_loo(ll, reff, want_diag=True)
    check_against_oracle(ll, reff, r)
    assert int(r["counters"][3]) == 0 and engine.handover_reasons() == {}   # nothing left the fast path


@pytest.mark.parametrize("S,N,reff", [(4000, 300, 0.5), (4000, 200, 0.3), (4000, 120, 0.2), (2000, 100, 0.3),
                                      (3000, 64, 0.16), (4096, 50, 0.15), (1100, 40, 0.25)])
def test_tile_path_long_tails_vs_oracle(S, N, reff):
    """Relative efficiencies well below 1 (what pl.loo derives from the ESS of a multi-chain posterior,
    pyloo/loo.py:204-216): M = 3 sqrt(S / reff) grows to several hundred draws, more than one per threshold bin;
    the tile path serves M + 2 <= 512 with threshold ranks near the top of the 32 bin minima."""
    M = engine.tail_length(S, reff)
    assert 230 <= M + 1 <= 511 or S < 4000
    rng = np.random.default_rng(S + N)
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.3, 2.0, size=(1, N))
    engine.profile(True)
    engine.handover_reasons()
    r = gpu_loo(ll, reff, want_diag=True)
    prof = engine.profile_read()
    engine.profile(False)
    check_against_oracle(ll, reff, r)
    assert prof["transpose"][1] == 0 and prof["stream"][1] >= 1           # the cluster kernel, no panels
    assert int(r["counters"][3]) <= max(1, N // 50), engine.handover_reasons(reset=False)


@pytest.mark.parametrize("S,N,reff", [(512, 40, 1.0), (600, 30, 0.8), (1000, 200, 1.0), (1000, 64, 0.5), (1022, 18, 1.0),
                                      (800, 50, 0.3)])
def test_tile_path_short_posteriors_vs_oracle(S, N, reff):
    """4 chains x 250 draws and the like: 64 to 128 draws per CTA of the cluster, 4 to 8 per thread."""
    rng = np.random.default_rng(3 * S + N)
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.3, 2.0, size=(1, N))
    engine.profile(True)
    engine.handover_reasons()
    r = gpu_loo(ll, reff, want_diag=True)
    prof = engine.profile_read()
    engine.profile(False)
    check_against_oracle(ll, reff, r)
    assert prof["transpose"][1] == 0 and prof["stream"][1] >= 1
    assert int(r["counters"][3]) <= max(1, N // 50), engine.handover_reasons(reset=False)


@pytest.mark.parametrize("csize", ["2", "4", "8"])
@pytest.mark.parametrize("S,N,reff", [(2000, 100, 1.0), (700, 60, 1.0)])
def test_tile_path_cluster_sizes_vs_oracle(S, N, reff, csize, monkeypatch):
    """The library picks the cluster size from S and M (2, 4 or 8 CTAs per tile); every size gives the oracle's
    numbers (sums differ in rounding only, the tail set not at all)."""
    monkeypatch.setenv("B2L_TILE_CSIZE", csize)
    rng = np.random.default_rng(S + int(csize))
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.3, 2.0, size=(1, N))
    r = gpu_loo(ll, reff, want_diag=True)
    check_against_oracle(ll, reff, r)


@pytest.mark.parametrize("S,N,reff", [(8000, 60, 1.0), (16000, 40, 1.0), (6000, 34, 0.8), (12000, 24, 1.0), (10000, 30, 0.7),
                                      (16384, 18, 1.0), (5000, 50, 1.0), (16000, 30, 0.5), (8000, 40, 0.25)])
def test_tile_path_long_posteriors_vs_oracle(S, N, reff):
    """More than 4096 draws (BASELINE configs[3] has 16 000): the draw axis is cut into 2, 3 or 4 equal chunks, every
    (tile, chunk) pair is a work unit of the cluster kernel, the column's chunks are folded by tile_merge_kernel and
    the tail kernel forms x = fl(r - max r) from the raw candidates.  One read of HBM, no transposed panels."""
    rng = np.random.default_rng(S + N)
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.3, 2.0, size=(1, N))
    engine.profile(True)
    engine.handover_reasons()
    r = gpu_loo(ll, reff, want_diag=True)
    prof = engine.profile_read()
    engine.profile(False)
    check_against_oracle(ll, reff, r)
    assert prof["transpose"][1] == 0 and prof["stream"][1] >= 1
    assert int(r["counters"][3]) <= max(1, N // 25), engine.handover_reasons(reset=False)


@pytest.mark.parametrize("chunks", ["2", "4"])
def test_tile_path_chunked_units_on_a_short_posterior(chunks, monkeypatch):
    """The chunked units forced onto S = 4000 (B2L_TILE_CHUNKS): same numbers as the oracle, the same tail index
    sets as the plain units, chains of different location as chunks (more hand-overs, same values)."""
    rng = np.random.default_rng(31)
    S, N = 4000, 120
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.3, 2.0, size=(1, N))
    ll[1000:2000, ::3] += 0.6          # one "chain" sits higher in every third column
    ll[5, 7] = np.nan
    ll[3000, 9] = -np.inf
    ll[:, 11] = -2.0
    plain = gpu_loo(ll, 1.0, want_diag=True, want_tail_idx=True)
    monkeypatch.setenv("B2L_TILE_CHUNKS", chunks)
    engine.handover_reasons()
    r = gpu_loo(ll, 1.0, want_diag=True, want_tail_idx=True)
    with np.errstate(all="ignore"):
        pw = orc.loo_pointwise(ll, 1.0)
        ww = orc.waic_pointwise(ll)
    for key, ref in (("elpd_i", pw["elpd_i"]), ("pareto_k", pw["pareto_k"]), ("lppd_i", pw["lppd_i"]),
                     ("var_i", ww["var_i"]), ("lppdw_i", ww["lppd_i"])):
        same_special(r[key], ref)
        close(r[key], ref, atol=1e-13)
    ok = np.isfinite(pw["pareto_k"])
    a = np.sort(plain["tail_idx"][ok], axis=1)
    b = np.sort(r["tail_idx"][ok], axis=1)
    assert np.array_equal(a, b)
    assert np.array_equal(plain["diag"][ok, 1], r["diag"][ok, 1])   # the cutoff, bit for bit


def test_host_loo_leaves_the_cluster_kernel_when_it_hands_over_most_columns():
    """Student-t(3) log-likelihoods: the normaliser of most columns is carried by a few draws, the cluster kernel hands
    them to the general kernel one by one.  The host pipeline notices (statistics record of a finished chunk) and sends
    the remaining chunks down the transposed-panel route; the values do not depend on the route.  B2L_FLAG_NO_TILE on
    the device entry point does the same on request."""
    rng = np.random.default_rng(41)
    S, N = 2000, 3072
    ll = -1.4 + rng.standard_t(3, size=(S, N))
    engine.profile(True)
    r = engine.loo_host(ll, 1.0, chunk_obs=256, device=0)
    prof = engine.profile_read()
    engine.profile(False)
    assert prof["stream"][1] >= 3 and prof["transpose"][1] >= 1       # both routes ran
    with np.errstate(all="ignore"):
        pw = orc.loo_pointwise(ll, 1.0)
    for key in ("elpd_i", "pareto_k", "lppd_i"):
        same_special(r[key], pw[key])
        close(r[key], pw[key], atol=1e-13)
    assert r["stats"].n == N


@pytest.mark.parametrize("S,reff", [(512, 1.0), (1000, 0.5), (2000, 0.25), (4000, 1.0), (8000, 0.5), (16000, 1.0)])
def test_tile_path_very_few_observations(S, reff):
    """Fewer observations than a tile holds, fewer work units than clusters, partial last tiles -- on every kind of
    plan (clusters of 2 / 4 / 8 CTAs, chunked units, long tails): the launch completes and matches the oracle."""
    rng = np.random.default_rng(S)
    for N in (2, 6, 10, 18, 34):
        ll = -1.4 + rng.normal(size=(S, N))
        r = gpu_loo(ll, reff)
        pw = orc.loo_pointwise(ll[:, :6], reff)
        close(r["elpd_i"][:6], pw["elpd_i"][:6])
        close(r["pareto_k"][:6], pw["pareto_k"][:6], atol=1e-13)
        assert np.isfinite(r["elpd_i"]).all()


def test_tile_path_runs_the_cluster_kernel():
    """The eligible shapes really take the tile kernel (per-kernel timers: no transpose launch)."""
    rng = np.random.default_rng(3)
    ll = torch.from_numpy(-1.4 + rng.normal(size=(2048, 96))).cuda()
    engine.profile(True)
    engine.loo_cuda(ll, 1.0)
    torch.cuda.synchronize()
    prof = engine.profile_read()
    engine.profile(False)
    assert prof["transpose"][1] == 0 and prof["stream"][1] >= 1 and prof["tail"][1] >= 1


def test_tile_path_special_columns_are_handed_over():
    """NaN / +-inf / constant / very wide columns leave the fast path for the general kernel, which reads the
    column where it lies (strided): same values as the oracle, counters as pyloo/loo.py:218-227."""
    rng = np.random.default_rng(21)
    S, N = 2000, 48
    ll = -1.4 + rng.normal(size=(S, N))
    ll[3, 2] = np.nan         # NaN -> -1e10 (loo.py:227)
    ll[5, 4] = -np.inf        # loo keeps it (elpd NaN, k inf)
    ll[7, 6] = np.inf
    ll[:, 8] = -2.5           # constant column -> k = inf
    ll[10, 9] = 1e10
    ll[11, 10] = -1e10
    ll[:, 20] *= 400.0        # range of ll far beyond 600: the exp(ll - min ll) sum would overflow
    ll[100, 33] = -900.0      # one far outlier below
    r = gpu_loo(ll, 1.0, want_diag=True)
    with np.errstate(all="ignore"):
        pw = orc.loo_pointwise(ll, 1.0)
        ww = orc.waic_pointwise(ll)
    for key, ref in (("elpd_i", pw["elpd_i"]), ("pareto_k", pw["pareto_k"]), ("lppd_i", pw["lppd_i"]),
                     ("var_i", ww["var_i"]), ("lppdw_i", ww["lppd_i"])):
        same_special(r[key], ref)
        # column 9 holds a +1e10 draw: the reference's own lw + ll cancels to ~1e-6 there (see test_loo_special_values)
        close(r[key], ref, rtol=1e-8 if key == "elpd_i" else RTOL, atol=1e-13)
    assert r["counters"][0] == 1 and r["counters"][1] == 1 and r["counters"][2] == 1
    assert int(r["counters"][3]) >= 6


def test_tile_path_heavy_tails():
    """Student-t log-likelihoods: many k > 0.7, short tails, wide columns -- fast path or hand-over, same numbers."""
    rng = np.random.default_rng(22)
    S, N = 4000, 128
    ll = -np.abs(rng.standard_t(2.5, size=(S, N))) * 3.0
    r = gpu_loo(ll, 1.0, want_diag=True)
    check_against_oracle(ll, 1.0, r)
    assert np.array_equal(r["pareto_k"] > 0.7, orc.loo_pointwise(ll, 1.0)["pareto_k"] > 0.7)


@pytest.mark.parametrize("kind", ["t3", "ties", "two_chains", "scaled_t5", "few_values"])
@pytest.mark.parametrize("seed", [1, 2])
def test_tile_path_awkward_columns_vs_oracle(kind, seed):
    """Random shapes with the distributions that stress the fast path: heavy tails whose raw tail carries nearly the
    whole normaliser (the all-draw total minus the raw tail cancels -- such columns must be handed over before the
    rounding of the total shows in elpd_i), exact ties at the cutoff, chains with different locations, columns with
    a handful of distinct values."""
    rng = np.random.default_rng(100 * seed + len(kind))
    S = int(rng.integers(1024, 4097))
    N = 2 * int(rng.integers(100, 300))
    reff = float(rng.choice([1.0, 0.9, 0.8, 0.72]))
    z = rng.normal(size=(S, N))
    if kind == "t3":
        z = rng.standard_t(3, size=(S, N))
    elif kind == "ties":
        z = np.round(z, 2)
    elif kind == "two_chains":
        z[: S // 2] += 0.8
        z[S // 2:] *= 1.7
    elif kind == "scaled_t5":
        z = rng.standard_t(5, size=(S, N)) * rng.uniform(0.2, 4.0, size=(1, N))
    elif kind == "few_values":
        z = np.round(rng.standard_t(4, size=(S, N)), 1)
        z[:, ::7] = np.floor(z[:, ::7])
    ll = -1.4 + z
    r = gpu_loo(ll, reff, want_diag=True)
    check_against_oracle(ll, reff, r)


def test_tile_path_autocorrelated_chains():
    """Four chains with different locations and strong autocorrelation: the CTAs of a cluster see different
    distributions (their draw segments are different chains); thresholds are medians over the CTAs."""
    rng = np.random.default_rng(23)
    S, N = 4000, 64
    z = rng.normal(size=(S, N))
    for s in range(1, S):
        z[s] = 0.9 * z[s - 1] + np.sqrt(1 - 0.81) * z[s]
    ll = -1.4 + z
    ll[1000:2000] += 0.5
    ll[3000:] -= 0.7
    r = gpu_loo(ll, 1.0, want_diag=True)
    check_against_oracle(ll, 1.0, r)


def test_tile_path_rounds_and_positions_do_not_matter(monkeypatch):
    """Batch invariance (pyloo/tests/base_tests/test_loo_i.py:41-58): many short rounds, one long round, the other
    tile width and a shifted column window give the same bits for every observation."""
    rng = np.random.default_rng(24)
    ll = -1.4 + rng.normal(size=(2000, 1000))
    keys = ("elpd_i", "pareto_k", "lppd_i", "var_i", "lppdw_i")
    base = gpu_loo(ll, 1.0)
    monkeypatch.setenv("B2L_TILE_ROUND", "64")
    short = gpu_loo(ll, 1.0)
    monkeypatch.delenv("B2L_TILE_ROUND")
    monkeypatch.setenv("B2L_TILE_W", "16")
    wide = gpu_loo(ll, 1.0)
    monkeypatch.delenv("B2L_TILE_W")
    window = gpu_loo(ll[:, 6:], 1.0)
    again = gpu_loo(ll, 1.0)
    for key in keys:
        assert np.array_equal(base[key], short[key])
        assert np.array_equal(base[key], wide[key])
        assert np.array_equal(base[key][6:], window[key])
        assert np.array_equal(base[key], again[key])


def test_tile_path_full_size_slab_properties():
    """BASELINE configs[2] shape on a 75 776-observation slab (S = 4000): no hand-overs on i.i.d. data, totals of
    the statistics record equal the NumPy reductions, a strided subset equals the oracle."""
    torch.manual_seed(5)
    S, N = 4000, 75_776
    ll = torch.randn(S, N, dtype=torch.float64, device="cuda") - 1.4
    res = engine.loo_cuda(ll, 1.0)
    st = engine.StatsRecord(engine.stats_cuda(res).cpu().numpy())
    torch.cuda.synchronize()
    assert int(res["counters"][3]) == 0 and st.n == N
    e = res["elpd_i"].cpu().numpy()
    close(st.elpd_sum, e.sum(), 1e-12)
    close(st.lppd_sum, res["lppd_i"].cpu().numpy().sum(), 1e-12)
    idx = np.arange(0, N, 1511)
    sub = ll[:, torch.from_numpy(idx).cuda()].cpu().numpy()
    pw = orc.loo_pointwise(sub, 1.0)
    close(e[idx], pw["elpd_i"])
    close(res["pareto_k"].cpu().numpy()[idx], pw["pareto_k"], atol=1e-13)
    close(res["lppd_i"].cpu().numpy()[idx], pw["lppd_i"], atol=1e-13)


# ------------------------------------------------------------------ several devices behind the host entry points
def _device_lists():
    n = torch.cuda.device_count() if has_cuda() else 1
    lists = [[0, 0], [0, 0, 0]]            # shards share a device: the shard logic without needing two GPUs
    if n >= 2:
        lists += [[0, 1], list(range(n))]
    return lists


def test_host_loo_over_device_shards_equals_one_device():
    """``pl.loo`` cuts the observation axis into one shard per device (b2l_loo_host_mgpu_f64): pointwise results are
    the same bits as the one-device call, the merged statistics record agrees to rounding (Chan merge in shard order)."""
    rng = np.random.default_rng(31)
    ll = -1.4 + rng.normal(size=(2000, 3000))
    one = engine.loo_host(ll, 1.0, device=0)
    for devs in _device_lists():
        many = engine.loo_host(ll, 1.0, devices=devs)
        for key in ("elpd_i", "pareto_k", "lppd_i", "var_i", "lppdw_i"):
            assert np.array_equal(one[key], many[key]), (devs, key)
        assert many["stats"].n == 3000
        close(many["stats"].elpd_sum, one["stats"].elpd_sum, 1e-13)
        close(many["stats"].elpd_m2, one["stats"].elpd_m2, 1e-11)
        close(many["stats"].waic_m2, one["stats"].waic_m2, 1e-11)
        assert many["stats"].k_gt_good == one["stats"].k_gt_good


def test_host_psislw_over_device_shards_and_memory_kinds():
    """Row shards of ``pl.psislw`` over a device list; pageable (bounce buffers), driver-staged and pinned callers give
    the same bits; the calling thread's current device is left alone."""
    rng = np.random.default_rng(32)
    x = rng.normal(size=(900, 1000))
    lw0, k0 = engine.psislw_host(x, 0.9, device=0)
    before = torch.cuda.current_device()
    for devs in _device_lists():
        lw, k = engine.psislw_host(x, 0.9, devices=devs)
        assert np.array_equal(lw, lw0) and np.array_equal(k, k0), devs
    assert torch.cuda.current_device() == before
    pinned = torch.empty((900, 1000), dtype=torch.float64, pin_memory=True)
    pinned.copy_(torch.from_numpy(x))
    lw_p, k_p = engine.psislw_host(pinned.numpy(), 0.9, device=0)
    assert np.array_equal(lw_p, lw0) and np.array_equal(k_p, k0)
    os.environ["B2L_HOST_BOUNCE"] = "0"
    try:
        lw_s, k_s = engine.psislw_host(x, 0.9, device=0)
        ll = -1.4 + rng.normal(size=(1200, 700))
        a = engine.loo_host(ll, 1.0, device=0)
    finally:
        del os.environ["B2L_HOST_BOUNCE"]
    b = engine.loo_host(ll, 1.0, device=0)
    assert np.array_equal(lw_s, lw0) and np.array_equal(k_s, k0)
    for key in ("elpd_i", "pareto_k", "lppd_i", "var_i"):
        assert np.array_equal(a[key], b[key])
    # column shards of an observation-fastest host matrix in psislw, chunks smaller than a shard
    xt = np.ascontiguousarray(x.T).T
    lw_t, k_t = engine.psislw_host(xt, 0.9, devices=[0, 0], chunk_obs=200)
    assert np.array_equal(lw_t, lw0) and np.array_equal(k_t, k0)


def test_host_device_selection_rules(monkeypatch):
    """LOCAL_RANK (torchrun) beats B2L_DEVICE; explicit arguments beat both; small inputs stay on one GPU."""
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    monkeypatch.delenv("B2L_DEVICE", raising=False)
    monkeypatch.delenv("B2L_DEVICES", raising=False)
    assert engine.host_devices(nbytes=1 << 20) == [0]
    assert engine.host_devices(nbytes=1 << 40) == list(range(torch.cuda.device_count()))
    monkeypatch.setenv("B2L_DEVICES", "0,0")
    assert engine.host_devices(nbytes=10) == [0, 0]
    monkeypatch.setenv("B2L_DEVICE", "3")
    assert engine.host_devices(nbytes=1 << 40) == [3]
    monkeypatch.setenv("LOCAL_RANK", "1")
    assert engine.host_devices(nbytes=1 << 40) == [1] and engine.current_device() == 1
    assert engine.host_devices(device=2) == [2] and engine.host_devices(devices=(0, 2)) == [0, 2]


# ------------------------------------------------------------------ the tail index sets themselves
def _oracle_tail_sets(x_ns, M):
    """Per row of log ratios: the draws with x > cutoff, exactly as pyloo/psis.py:134-139 selects them."""
    sets = []
    for row in x_ns:
        z = row - row.max()
        c = max(z[np.argsort(z)[-M - 1]], orc.CUTOFFMIN)
        sets.append(set(np.where(z > c)[0].tolist()))
    return sets


@pytest.mark.parametrize("name,path", [("cfg2_normal_s4000.npz", "tile"), ("cfg5_student_t_s8000.npz", "rows"),
                                       ("cfg2_normal_s4000.npz", "rows"), ("cfg2_normal_s4000.npz", "general")])
def test_tail_index_sets_are_bit_exact(name, path, monkeypatch):
    """north_star: "the selected tail indices are bit-exact".  The nullable tail_idx output of b2l_loo_dev_ex_f64
    (N x M draw indices, -1 padded) equals, as a set, np.where(x > x_cutoff) of the reference's selection on the
    golden inputs of BASELINE configs[1] and [4] -- through the tile kernel (matrix in the (S, N) layout), the
    split row path and the general kernel."""
    from b2l_testutil import golden

    g = golden(name)
    x = g["x"]                                   # (N, S) raw log ratios; loo sees ll = -x
    reff = float(g["reff"])
    M = engine.tail_length(x.shape[1], reff)
    if path == "general":
        monkeypatch.setenv("B2L_FORCE_LEGACY", "1")
    t = torch.from_numpy(np.ascontiguousarray(-x.T) if path == "tile" else np.ascontiguousarray(-x)).cuda()
    res = engine.loo_cuda(t if path == "tile" else t.t(), reff, want_tail_idx=True)
    torch.cuda.synchronize()
    idx = res["tail_idx"].cpu().numpy()
    assert idx.shape == (x.shape[0], M)
    want = _oracle_tail_sets(x, M)
    for i, row in enumerate(idx):
        got = row[row >= 0]
        assert len(got) == len(set(got.tolist()))            # no draw twice
        assert set(got.tolist()) == want[i], i


def test_host_loo_with_an_odd_number_of_observations_takes_the_tile_path():
    """ArviZ data has any N.  The host entry pads the device pitch of its chunks to an even number of doubles (the
    2-D TMA tiles need a 16-byte pitch), so an odd N -- and odd chunks -- still run the tile kernel."""
    rng = np.random.default_rng(41)
    S, N = 2000, 701
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.5, 1.5, size=(1, N))
    engine.profile(True)
    r = engine.loo_host(ll, 1.0, device=0, chunk_obs=333)       # chunks of 333, 333, 35 observations
    prof = engine.profile_read()
    engine.profile(False)
    assert prof["transpose"][1] == 0 and prof["stream"][1] == 3 and prof["tail"][1] == 3
    pw = orc.loo_pointwise(ll, 1.0)
    ww = orc.waic_pointwise(ll)
    close(r["elpd_i"], pw["elpd_i"])
    close(r["pareto_k"], pw["pareto_k"], atol=1e-13)
    close(r["lppd_i"], pw["lppd_i"], atol=1e-13)
    close(r["var_i"], ww["var_i"])
    pinned = torch.empty((S, N), dtype=torch.float64, pin_memory=True)
    pinned.copy_(torch.from_numpy(ll))
    r2 = engine.loo_host(pinned.numpy(), 1.0, device=0, chunk_obs=333)
    for key in ("elpd_i", "pareto_k", "lppd_i", "var_i"):
        assert np.array_equal(r[key], r2[key])

"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs and against the committed golden vectors produced by the real reference.

Tolerances (BASELINE.json north_star): Pareto k, elpd_i, lppd_i, var_i within 1e-10 relative;
tail selection bit-exact (cutoff value and tail count identical); identical inf/NaN pattern and
k > 0.7 flags.  Log weights: 1e-10 relative with a 1e-12 absolute floor (a log weight can be
arbitrarily close to 0, where "relative" has no meaning)."""

import numpy as np
import pytest

from b2l_testutil import golden, has_cuda

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


def gpu_psislw(x, reff, diag=False):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out = engine.psislw_cuda(t, reff, want_diag=diag)
    torch.cuda.synchronize()
    return tuple(o.cpu().numpy() for o in out)


def gpu_loo(ll_sn, reff, **kw):
    res = engine.loo_cuda(torch.from_numpy(np.ascontiguousarray(ll_sn)).cuda(), reff, **kw)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items() if k != "workspace"}


def oracle_tail(x, M):
    """(cutoff value, tail count) per row, exactly as pyloo/psis.py:134-141 defines them."""
    cut, cnt = [], []
    for row in x:
        z = row - row.max()
        c = max(np.sort(z)[-M - 1], orc.CUTOFFMIN)
        cut.append(c)
        cnt.append(int((z > c).sum()))
    return np.array(cut), np.array(cnt)


# ------------------------------------------------------------------ golden vectors (real reference)
def test_golden_cfg1_create_model():
    g = golden("cfg1_create_model.npz")
    for tag, reff in (("r10", 1.0), ("r07", 0.7)):
        r = gpu_loo(g["ll_sn"], reff)
        close(r["elpd_i"], g[f"elpd_i_{tag}"])
        close(r["pareto_k"], g[f"k_{tag}"])
        close(r["lppd_i"], g[f"lppd_i_{tag}"])
        close(r["var_i"], g["var_i"])
        lw, k = gpu_psislw(-g["ll_sn"].T, reff)
        close(lw, g[f"lw_{tag}"], atol=1e-12)
        close(k, g[f"k_{tag}"])


def test_golden_cfg2_normal_s4000():
    g = golden("cfg2_normal_s4000.npz")
    lw, k, diag = gpu_psislw(g["x"], float(g["reff"]), diag=True)
    close(k, g["k"])
    close(lw[:8], g["lw"], atol=1e-12)
    cut, cnt = oracle_tail(g["x"], 200)
    assert np.array_equal(diag[:, 1], cut)           # bit-exact cutoff => bit-exact tail index set
    assert np.array_equal(diag[:, 2].astype(int), cnt)
    close(np.log(np.exp(lw).sum(axis=1)), 0.0, rtol=0, atol=1e-12)  # weights normalise (test_psis.py:25)


@pytest.mark.parametrize("name", ["cfg3_loo_s4000.npz", "cfg4_loo_s16000.npz"])
def test_golden_cfg34_loo(name):
    g = golden(name)
    r = gpu_loo(g["ll_sn"], float(g["reff"]))
    close(r["elpd_i"], g["elpd_i"])
    close(r["pareto_k"], g["k"])
    close(r["lppd_i"], g["lppd_i"])
    close(r["var_i"], g["var_i"])
    close(r["lppdw_i"], g["lppd_i"])


def test_golden_cfg5_student_t():
    g = golden("cfg5_student_t_s8000.npz")
    lw, k = gpu_psislw(g["x"], float(g["reff"]))
    close(k, g["k"])
    same_special(k, g["k"])
    assert np.array_equal(k > 0.7, g["k"] > 0.7)      # identical k > 0.7 flags
    close(lw[:4], g["lw"], atol=1e-12)
    close(lw.max(axis=1), g["lw_max"], atol=1e-12)
    close(lw.min(axis=1), g["lw_min"], atol=1e-12)
    r = gpu_loo(np.ascontiguousarray(-g["x"].T), float(g["reff"]))
    close(r["elpd_i"], g["elpd_i"])
    close(r["lppd_i"], g["lppd_i"])
    close(r["pareto_k"], g["k"])


@pytest.mark.parametrize("name", ["short4", "const100", "len8", "ties", "nan", "big", "clamp", "s33"])
def test_golden_edge_cases(name):
    g = golden("edge_cases.npz")
    x = np.atleast_2d(g[f"{name}_x"])
    lw, k = gpu_psislw(x, 1.0)
    ref_lw, ref_k = np.atleast_2d(g[f"{name}_lw"]), np.atleast_1d(g[f"{name}_k"])
    close(k, ref_k)
    same_special(k, ref_k)
    same_special(lw, ref_lw)
    if name == "ties":  # tie order inside the tail is unspecified in the reference (App. D)
        close(np.sort(lw, axis=-1), np.sort(ref_lw, axis=-1), atol=1e-12)
    else:
        close(lw, ref_lw, atol=1e-12)


# ------------------------------------------------------------------ seeded inputs vs the oracle
@pytest.mark.parametrize("S,N,reff,scale", [(4000, 1024, 0.9, 1.0), (1000, 256, 1.0, 3.0), (2000, 128, 0.5, 0.2),
                                            (600, 64, 1.0, 1.0), (513, 33, 1.0, 1.0), (1001, 40, 0.8, 1.0),
                                            (16000, 24, 1.0, 1.3), (25000, 4, 1.0, 1.0)])
def test_psislw_vs_oracle(S, N, reff, scale):
    rng = np.random.default_rng(S + N)
    x = scale * rng.normal(size=(N, S))
    lw, k, diag = gpu_psislw(x, reff, diag=True)
    ref_lw, ref_k = orc.psislw(x, reff)
    close(k, ref_k, atol=1e-13)
    close(lw, ref_lw, atol=1e-12)
    M = orc.tail_length(S, reff)
    cut, cnt = oracle_tail(x, M)
    assert np.array_equal(diag[:, 1], cut)
    assert np.array_equal(diag[:, 2].astype(int), cnt)


@pytest.mark.parametrize("S,N,reff", [(4000, 96, 0.12), (4000, 64, 0.08), (6000, 48, 0.1), (16000, 20, 0.3)])
def test_tails_beyond_510_draws(S, N, reff):
    """r_eff far below 1: M = 3 sqrt(S / r_eff) passes the 510 draws two 512-key sorts cover; the tail kernel's
    1024-key build (TL = 32) serves M + 2 <= 800 on the row route (psislw, and loo on shapes the cluster kernel does
    not take)."""
    M = orc.tail_length(S, reff)
    assert 510 < M <= 798
    rng = np.random.default_rng(S + N)
    x = 1.3 * rng.normal(size=(N, S))
    engine.handover_reasons()
    lw, k, diag = gpu_psislw(x, reff, diag=True)
    hand = engine.handover_reasons()
    ref_lw, ref_k = orc.psislw(x, reff)
    close(k, ref_k, atol=1e-13)
    close(lw, ref_lw, atol=1e-12)
    cut, cnt = oracle_tail(x, M)
    assert np.array_equal(diag[:, 1], cut) and np.array_equal(diag[:, 2].astype(int), cnt)
    assert sum(hand.values()) <= max(2, N // 10), hand          # the split path did the work, not the general kernel
    r = gpu_loo(np.ascontiguousarray(-x.T), reff)
    pw = orc.loo_pointwise(-x.T, reff)
    close(r["elpd_i"], pw["elpd_i"])
    close(r["pareto_k"], pw["pareto_k"], atol=1e-13)
    close(r["lppd_i"], pw["lppd_i"])


def test_psislw_heavy_tail_stress_cfg5_shape():
    rng = np.random.default_rng(55)
    x = rng.standard_t(1.5, size=(384, 8000))
    lw, k = gpu_psislw(x, 1.0)
    with np.errstate(all="ignore"):
        ref_lw, ref_k = orc.psislw(x, 1.0)
    close(k, ref_k)
    same_special(k, ref_k)
    assert np.array_equal(k > 0.7, ref_k > 0.7)
    close(lw, ref_lw, atol=1e-12)


def test_psislw_special_rows():
    rng = np.random.default_rng(7)
    x = rng.normal(size=(10, 600))
    x[1, 17] = np.nan            # NaN row -> all-NaN weights, k = inf (App. D)
    x[2, 5] = np.inf             # +inf -> inf - inf = NaN
    x[3, 9] = -np.inf            # zero-weight draw: fine
    x[4, :] = 1.0                # constant row -> -log S, k = inf (test_psis.py:121-125)
    x[5, :300] = -np.inf
    x[6, :] = -np.inf            # all -inf -> NaN
    x[7] = np.round(x[7], 1)     # ties
    x[8] *= 500.0                # cutoffmin clamp
    lw, k = gpu_psislw(x, 1.0)
    with np.errstate(all="ignore"):
        ref_lw, ref_k = orc.psislw(x, 1.0)
    same_special(k, ref_k)
    same_special(lw, ref_lw)
    close(k, ref_k)
    close(np.sort(lw, axis=-1), np.sort(ref_lw, axis=-1), atol=1e-12)
    close(lw[4], -np.log(600.0), rtol=1e-12)


@pytest.mark.parametrize("S,N,reff", [(4000, 700, 1.0), (2000, 33, 0.7), (8000, 150, 1.0), (1000, 1, 1.0)])
def test_loo_vs_oracle_obs_fastest(S, N, reff):
    rng = np.random.default_rng(S * 3 + N)
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.3, 2.0, size=(1, N))
    r = gpu_loo(ll, reff)
    pw = orc.loo_pointwise(ll, reff)
    ww = orc.waic_pointwise(ll)
    close(r["elpd_i"], pw["elpd_i"])
    close(r["pareto_k"], pw["pareto_k"], atol=1e-13)   # k can sit arbitrarily close to 0
    close(r["lppd_i"], pw["lppd_i"])
    close(r["var_i"], ww["var_i"])
    close(r["lppdw_i"], ww["lppd_i"])


def test_loo_special_values():
    rng = np.random.default_rng(11)
    ll = -1.4 + rng.normal(size=(1000, 12))
    ll[3, 2] = np.nan        # NaN -> -1e10 (loo.py:227)
    ll[5, 4] = -np.inf       # loo keeps it (elpd NaN, k inf); waic -> -1e10
    ll[7, 6] = np.inf        # loo: NaN; waic -> +1e10
    ll[:, 8] = -2.5          # constant column -> k = inf (test_loo.py:89-97)
    ll[10, 9] = 1e10
    ll[11, 10] = -1e10
    r = gpu_loo(ll, 1.0)
    with np.errstate(all="ignore"):
        pw = orc.loo_pointwise(ll, 1.0)
        ww = orc.waic_pointwise(ll)
    for key_g, ref in (("elpd_i", pw["elpd_i"]), ("pareto_k", pw["pareto_k"]), ("lppd_i", pw["lppd_i"]),
                       ("var_i", ww["var_i"]), ("lppdw_i", ww["lppd_i"])):
        same_special(r[key_g], ref)
        # obs 9 holds a +1e10 draw: the reference forms lw + ll = fl(fl(-1e10 - max - lse) + 1e10), which
        # cancels to ~1e-6 absolute, i.e. its own elpd_i carries ~1e-9 of rounding noise there
        # (pyloo/loo.py:289).  The kernel's closed form has no such cancellation; compare at 1e-8.
        close(r[key_g], ref, rtol=1e-8 if key_g == "elpd_i" else RTOL)
    assert r["counters"][0] == 1 and r["counters"][1] == 1 and r["counters"][2] == 1


def test_waic_only_flag_skips_psis_but_keeps_waic_outputs():
    rng = np.random.default_rng(12)
    ll = -1.0 + rng.normal(size=(1500, 50))
    full = gpu_loo(ll, 1.0)
    wo = gpu_loo(ll, 1.0, waic_only=True)
    # the WAIC-only launch runs the general row kernel, the full one the split path: same numbers
    # to rounding (different but equivalent shifts inside the log-sum-exp), not the same bits
    close(full["lppd_i"], wo["lppd_i"], rtol=1e-13)
    close(full["var_i"], wo["var_i"], rtol=1e-13)
    assert np.all(np.isinf(wo["pareto_k"]))


@pytest.mark.parametrize("S,N", [(3, 3), (7, 33), (63, 31), (64, 40), (1000, 70), (4000, 257), (16000, 9)])
def test_waic_column_kernel_against_oracle(S, N):
    """WAIC alone on the (chain, draw, obs) layout runs the one-pass column kernel (online logsumexp + chunked
    Chan variance): against the oracle for ragged S / N, an outlier as the first draw, NaN and +-inf columns."""
    rng = np.random.default_rng(S * 1000 + N)
    ll = -1.4 + rng.normal(size=(S, N)) * rng.uniform(0.1, 3.0, size=N)
    ll[0, 0] = 40.0                                   # the variance shift starts on an outlier
    if S >= 7:
        ll[3, 1] = np.nan
        ll[5, 2 % N] = np.inf
        ll[6, min(N - 1, 4)] = -np.inf
    if N > 6:
        ll[:, 5] = -np.inf                            # loo-policy lppd_i is NaN, the WAIC one is -1e10
        ll[:, 6] = 0.25                               # constant column: variance exactly 0
    wo = gpu_loo(ll, 1.0, waic_only=True)
    with np.errstate(all="ignore"):
        ww = orc.waic_pointwise(ll)
        lppd_loo = np.array([orc.logsumexp_row(np.where(np.isnan(c), -1e10, c), b_inv=S) for c in ll.T])
    close(wo["lppdw_i"], ww["lppd_i"])
    close(wo["var_i"], ww["var_i"], rtol=1e-10, atol=1e-18)
    same_special(wo["lppd_i"], lppd_loo)
    close(wo["lppd_i"], lppd_loo)
    assert np.all(np.isinf(wo["pareto_k"])) and np.all(np.isnan(wo["elpd_i"]))
    c = wo["counters"]
    assert c[0] == int(np.isnan(ll).sum()) and c[1] == int(np.isposinf(ll).sum()) and c[2] == int(np.isneginf(ll).sum())


def test_layouts_agree_bitwise():
    """Rows layout, obs-fastest layout and odd-S / unaligned (non-TMA) path agree: an observation's result
    must not depend on the tile it lands in (batch-invariance property,
    pyloo/tests/base_tests/test_loo_i.py:41-58).  The obs-fastest layout runs the tile kernel (2-D TMA tiles
    of the matrix where it lies): the tail index set and hence Pareto k are the same bits as the row path's,
    the normalising sums are accumulated in another order (1e-12)."""
    rng = np.random.default_rng(13)
    ll_ns = np.ascontiguousarray(-1.4 + rng.normal(size=(1300, 2000)))      # rows contiguous
    a = gpu_loo(ll_ns.T, 1.0)                                               # stride_s == 1
    b = gpu_loo(np.ascontiguousarray(ll_ns.T), 1.0)                         # stride_n == 1
    assert np.array_equal(a["pareto_k"], b["pareto_k"])
    for key in ("elpd_i", "lppd_i", "var_i"):
        close(a[key], b[key], rtol=1e-12)
    # inside the tile path: a column's result does not depend on the tile / round / position it lands in
    c = gpu_loo(np.ascontiguousarray(ll_ns[22:1300].T), 1.0)
    for key in ("elpd_i", "pareto_k", "lppd_i", "var_i"):
        assert np.array_equal(b[key][22:], c[key])
    sub = gpu_loo(np.ascontiguousarray(ll_ns[37:38].T), 1.0)                # one observation alone (general kernel)
    assert a["pareto_k"][37] == sub["pareto_k"][0]
    for key in ("elpd_i", "lppd_i", "var_i"):
        close(a[key][37], sub[key][0], rtol=1e-13)
    # unaligned rows (base pointer off by 8 bytes) -> cooperative-load path
    big = torch.from_numpy(ll_ns).cuda()
    flat = torch.empty(ll_ns.size + 1, dtype=torch.float64, device="cuda")
    flat[1:] = big.reshape(-1)
    shifted = flat[1:].view(1300, 2000)
    out1, k1 = engine.psislw_cuda(big, 1.0)
    out2, k2 = engine.psislw_cuda(shifted, 1.0)
    torch.cuda.synchronize()
    # (general row kernel vs split path: equal to rounding, the tail index set is the same)
    assert torch.allclose(k1, k2, rtol=1e-11, atol=1e-13) and torch.allclose(out1, out2, rtol=1e-12, atol=1e-12)
    # obs-fastest psislw in and out
    xt = big.t().contiguous()                                                # (S, N)
    out3, k3 = engine.psislw_cuda(xt.t(), 1.0, out=torch.empty_like(xt).t())
    torch.cuda.synchronize()
    assert torch.equal(k1, k3) and torch.equal(out1, out3.contiguous())


def test_stats_record_matches_numpy_and_merges():
    rng = np.random.default_rng(14)
    ll = -1.4 + rng.normal(size=(2000, 3001))
    t = torch.from_numpy(ll).cuda()
    res = engine.loo_cuda(t, 1.0)
    st = engine.StatsRecord(engine.stats_cuda(res).cpu().numpy())
    e = res["elpd_i"].cpu().numpy()
    w = (res["lppdw_i"] - res["var_i"]).cpu().numpy()
    assert st.n == 3001
    close(st.elpd_sum, e.sum(), 1e-12)
    close(st.elpd_m2 / st.n, np.var(e), 1e-11)
    close(st.lppd_sum, res["lppd_i"].cpu().numpy().sum(), 1e-12)
    close(st.p_waic_sum, res["var_i"].cpu().numpy().sum(), 1e-12)
    close(st.waic_m2 / st.n, np.var(w), 1e-11)
    assert st.k_gt_good == int((res["pareto_k"].cpu().numpy() > engine.good_k_threshold(2000)).sum())
    # two shards merged == one shard (multi-GPU invariance of se)
    parts = []
    for sl in (slice(0, 1200), slice(1200, 3001)):
        rr = engine.loo_cuda(t[:, sl], 1.0)
        parts.append(engine.stats_cuda(rr).cpu().numpy())
    merged = engine.stats_merge(parts)
    close(merged.elpd_sum, st.elpd_sum, 1e-13)
    close(merged.elpd_m2, st.elpd_m2, 1e-11)
    close(merged.waic_m2, st.waic_m2, 1e-11)
    assert merged.n == st.n


def test_host_entry_points_chunked_equal_device_path():
    rng = np.random.default_rng(15)
    x = rng.normal(size=(3000, 1000))
    lw_h, k_h = engine.psislw_host(x, 0.9, chunk_obs=700)
    lw_d, k_d = gpu_psislw(x, 0.9)
    assert np.array_equal(lw_h, lw_d) and np.array_equal(k_h, k_d)
    lw_t, k_t = engine.psislw_host(np.ascontiguousarray(x.T).T, 0.9, chunk_obs=999)  # obs-fastest host view
    assert np.array_equal(lw_t, lw_d) and np.array_equal(k_t, k_d)
    ll = -1.4 + rng.normal(size=(1000, 2500))
    r = engine.loo_host(ll, 1.0, chunk_obs=640)
    d = gpu_loo(ll, 1.0)
    for key in ("elpd_i", "pareto_k", "lppd_i", "var_i", "lppdw_i"):
        assert np.array_equal(r[key], d[key])
    assert r["stats"].n == 2500
    close(r["stats"].elpd_sum, d["elpd_i"].sum(), 1e-12)


def test_full_size_properties_cfg2():
    """BASELINE configs[1] at full size (S = 4000, N = 100 000): size-independent properties."""
    torch.manual_seed(1)
    N, S = 100_000, 4000
    x = torch.randn(N, S, dtype=torch.float64, device="cuda")
    out, k = engine.psislw_cuda(x, 0.9)
    torch.cuda.synchronize()
    lse = torch.logsumexp(out, dim=1)
    assert float(lse.abs().max()) < 1e-11                         # weights normalise
    assert bool(torch.isfinite(k).all()) and float(k.max()) < 0.7 and float(k.min()) > -0.5
    assert float(out.max()) <= 0.0                                # truncated at the max raw weight
    # untouched body: lw_out - lw is constant over the S - M smallest draws of each row
    idx = torch.arange(0, N, 997, device="cuda")
    xs, order = torch.sort(x[idx], dim=1)
    body = torch.gather(out[idx], 1, order)[:, : S - 201] - xs[:, : S - 201]
    assert float((body - body[:, :1]).abs().max()) < 1e-12
    # parity on a strided subset against the oracle
    rows = x[idx[:48]].cpu().numpy()
    ref_lw, ref_k = orc.psislw(rows, 0.9)
    close(k[idx[:48]].cpu().numpy(), ref_k)
    close(out[idx[:48]].cpu().numpy(), ref_lw, atol=1e-12)
    # shifting a row by a constant leaves the normalised weights and k unchanged (to rounding)
    out2, k2 = engine.psislw_cuda(x[:4096] + 3.0, 0.9)
    torch.cuda.synchronize()
    assert float((k2 - k[:4096]).abs().max()) < 1e-9


def test_packed_and_full_key_candidate_paths_agree_bitwise(monkeypatch):
    """The 31-bit packed candidate sort (fast path) and the full 64-bit (key, index) sort give the
    same bits; forcing the latter exercises the collision-fallback code."""
    rng = np.random.default_rng(21)
    x = np.ascontiguousarray(rng.normal(size=(600, 4000)))
    monkeypatch.setenv("B2L_SPLIT", "0")   # both runs on the general row kernel
    fast = gpu_psislw(x, 0.9)
    monkeypatch.setenv("B2L_FORCE_LEGACY", "1")
    slow = gpu_psislw(x, 0.9)
    monkeypatch.delenv("B2L_FORCE_LEGACY")
    monkeypatch.delenv("B2L_SPLIT")
    assert np.array_equal(fast[0], slow[0]) and np.array_equal(fast[1], slow[1])


def test_quantisation_collisions_fall_back_to_exact_order():
    """Distinct doubles closer than the 31-bit quantisation step inside the tail: detected, exact order kept."""
    rng = np.random.default_rng(22)
    x = rng.normal(size=(64, 4000))
    top = np.argsort(x, axis=1)[:, -50:]
    for i in range(64):                      # squeeze the 50 largest draws into a 1e-13-wide cluster
        x[i, top[i]] = 3.0 + 1e-15 * rng.permutation(50) * (i + 1)
    lw, k, diag = gpu_psislw(x, 0.9, diag=True)
    with np.errstate(all="ignore"):
        ref_lw, ref_k = orc.psislw(x, 0.9)
    close(k, ref_k)
    same_special(k, ref_k)
    close(lw, ref_lw, atol=1e-12)
    cut, cnt = oracle_tail(x, 200)
    assert np.array_equal(diag[:, 1], cut) and np.array_equal(diag[:, 2].astype(int), cnt)


# ------------------------------------------------------------------ split path (stream + tail kernels)
def _ar1(rng, n, s, rho):
    e = rng.normal(size=(n, s))
    x = np.empty_like(e)
    x[:, 0] = e[:, 0]
    for t in range(1, s):
        x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho * rho) * e[:, t]
    return x


@pytest.mark.parametrize("kind", ["normal", "student_t", "ar1", "duplicates", "trend", "lognormal"])
def test_split_path_vs_oracle_and_general_kernel(kind, monkeypatch):
    """The split path (register-resident stream kernel + warp-per-observation tail kernel) against the
    oracle and against the general row kernel on inputs that stress the threshold guess: heavy tails,
    strong autocorrelation (clustered tail draws -> retries), repeated draws (exact ties, also at the
    cutoff -> hand-over to the general kernel) and a trend along the chain."""
    rng = np.random.default_rng(31)
    N, S = 192, 4000
    if kind == "normal":
        x = rng.normal(size=(N, S))
    elif kind == "student_t":
        x = rng.standard_t(1.5, size=(N, S))
    elif kind == "ar1":
        x = _ar1(rng, N, S, 0.97)
    elif kind == "duplicates":   # Metropolis-like chains: each draw repeated a random number of times
        base = rng.normal(size=(N, S))
        idx = np.sort(rng.integers(0, S // 3, size=(N, S)), axis=1)
        x = np.take_along_axis(base, idx, axis=1)
    elif kind == "trend":
        x = rng.normal(size=(N, S)) + np.linspace(0, 4, S)[None, :]
    else:
        x = np.exp(rng.normal(size=(N, S)))
    lw, k, diag = gpu_psislw(x, 0.9, diag=True)
    with np.errstate(all="ignore"):
        ref_lw, ref_k = orc.psislw(x, 0.9)
    close(k, ref_k, atol=1e-13)
    same_special(k, ref_k)
    if kind == "duplicates":   # tie order inside the tail is unspecified in the reference
        close(np.sort(lw, axis=-1), np.sort(ref_lw, axis=-1), atol=1e-12)
    else:
        close(lw, ref_lw, atol=1e-12)
    cut, cnt = oracle_tail(x, 200)
    assert np.array_equal(diag[:, 1], cut) and np.array_equal(diag[:, 2].astype(int), cnt)
    monkeypatch.setenv("B2L_SPLIT", "0")
    lw_g, k_g, diag_g = gpu_psislw(x, 0.9, diag=True)
    monkeypatch.delenv("B2L_SPLIT")
    close(k, k_g, rtol=1e-11, atol=1e-13)
    close(lw, lw_g, rtol=1e-12, atol=1e-12)
    assert np.array_equal(diag[:, 1], diag_g[:, 1]) and np.array_equal(diag[:, 2], diag_g[:, 2])
    # LOO mode of the same path
    r = gpu_loo(np.ascontiguousarray(-x.T), 0.9)
    with np.errstate(all="ignore"):
        pw = orc.loo_pointwise(-x.T, 0.9)
    close(r["elpd_i"], pw["elpd_i"])
    close(r["pareto_k"], pw["pareto_k"], atol=1e-13)
    close(r["lppd_i"], pw["lppd_i"])


def test_split_path_is_batch_invariant():
    """An observation's bits do not depend on which stream/tail batch or CTA it lands in."""
    rng = np.random.default_rng(32)
    x = np.ascontiguousarray(rng.normal(size=(9000, 1000)))       # > one 8192-row round? (S = 1000: 16384 rows)
    lw, k = gpu_psislw(x, 1.0)
    lw2, k2 = gpu_psislw(x[4321:4400], 1.0)
    assert np.array_equal(lw[4321:4400], lw2) and np.array_equal(k[4321:4400], k2)


def test_heavy_tailed_rows_stay_on_the_split_path():
    """BASELINE configs[4] shape: Student-t(1.5) log-ratios, S = 8000.  Almost every row has its cutoff
    clamped at log(DBL_MIN) (psis.py:136) and k > 0.7; the split path must handle them itself (hand-overs to
    the general kernel are counted) and agree with the oracle."""
    rng = np.random.default_rng(56)
    x = rng.standard_t(1.5, size=(256, 8000))
    # observation rows contiguous: the (S, N) view of an (N, S) matrix takes the row route (stream + tail kernels)
    res = engine.loo_cuda(torch.from_numpy(np.ascontiguousarray(-x)).cuda().T, 1.0)
    torch.cuda.synchronize()
    r = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items() if k != "workspace"}
    with np.errstate(all="ignore"):
        pw = orc.loo_pointwise(-x.T, 1.0)
    close(r["pareto_k"], pw["pareto_k"], atol=1e-13)
    same_special(r["pareto_k"], pw["pareto_k"])
    close(r["elpd_i"], pw["elpd_i"])
    close(r["lppd_i"], pw["lppd_i"])
    assert (pw["pareto_k"] > 0.7).mean() > 0.9
    assert int(r["counters"][3]) <= 26          # <= 10 % of the rows handed over


@pytest.mark.parametrize("S", [64, 130, 1022, 2050, 3000, 4002, 4094, 4096, 4098, 6002, 12288])
def test_split_path_shape_boundaries(S):
    """Draw counts around the stream kernel's thread x register shapes (partial last slots, pads, the
    switch to the next shape) and piecewise apply transfers."""
    rng = np.random.default_rng(S)
    N = 37
    x = rng.normal(size=(N, S)) * 1.7
    lw, k, diag = gpu_psislw(x, 1.0, diag=True)
    ref_lw, ref_k = orc.psislw(x, 1.0)
    close(k, ref_k, atol=1e-13)
    close(lw, ref_lw, atol=1e-12)
    M = orc.tail_length(S, 1.0)
    cut, cnt = oracle_tail(x, M)
    assert np.array_equal(diag[:, 1], cut) and np.array_equal(diag[:, 2].astype(int), cnt)
    r = gpu_loo(np.ascontiguousarray(-x.T), 1.0)
    pw = orc.loo_pointwise(-x.T, 1.0)
    close(r["elpd_i"], pw["elpd_i"])
    close(r["pareto_k"], pw["pareto_k"], atol=1e-13)
    close(r["lppd_i"], pw["lppd_i"])


def test_split_path_padded_rows_and_many_rounds(monkeypatch):
    """Row stride larger than S (aligned padding) on input and output, and more than three rounds of the
    stream -> tail -> apply pipeline (scratch slots alternate, the last apply stage runs alone)."""
    rng = np.random.default_rng(77)
    N, S, pad = 700, 2000, 6
    x = rng.normal(size=(N, S))
    buf = torch.zeros((N, S + pad), dtype=torch.float64, device="cuda")
    buf[:, :S] = torch.from_numpy(x).cuda()
    out = torch.full((N, S + pad), 7.0, dtype=torch.float64, device="cuda")
    monkeypatch.setenv("B2L_BATCH", "96")          # 8 rounds
    lw, k = engine.psislw_cuda(buf[:, :S], 1.0, out=out[:, :S])
    torch.cuda.synchronize()
    monkeypatch.delenv("B2L_BATCH")
    ref_lw, ref_k = orc.psislw(x, 1.0)
    close(k.cpu().numpy(), ref_k, atol=1e-13)
    close(out[:, :S].cpu().numpy(), ref_lw, atol=1e-12)
    assert bool((out[:, S:] == 7.0).all())          # nothing written past the rows
    assert bool((buf[:, :S].cpu() == torch.from_numpy(x)).all())   # input untouched (psis.py:78)


def test_split_path_is_deterministic_across_runs_and_round_sizes(monkeypatch):
    """Same bits on every run and for every round size: the order in which the stream kernel's warps emit
    candidates depends on atomics, but nothing downstream may depend on it (exact re-ranking of equal sort
    keys, order-independent fixed-point sum for the candidates outside the tail)."""
    torch.manual_seed(5)
    x = torch.randn(20000, 4000, dtype=torch.float64, device="cuda")
    out1, k1 = engine.psislw_cuda(x, 0.9)
    out2, k2 = engine.psislw_cuda(x, 0.9)
    monkeypatch.setenv("B2L_BATCH", "3000")
    out3, k3 = engine.psislw_cuda(x, 0.9)
    monkeypatch.delenv("B2L_BATCH")
    torch.cuda.synchronize()
    assert torch.equal(k1, k2) and torch.equal(out1, out2)
    assert torch.equal(k1, k3) and torch.equal(out1, out3)
    r1 = engine.loo_cuda(x.t(), 1.0)
    r2 = engine.loo_cuda(x.t(), 1.0)
    torch.cuda.synchronize()
    for key in ("elpd_i", "pareto_k", "lppd_i", "var_i"):
        assert torch.equal(r1[key], r2[key])

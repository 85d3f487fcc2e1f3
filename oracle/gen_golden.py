"""Generate tests/golden/*.npz from the REAL reference (build container only).

    python oracle/gen_golden.py

Imports the reference's own ``_psislw`` / ``_gpdfit`` / ``_gpinv`` / ``_logsumexp`` /
``make_ufunc`` from ``/root/reference`` (see ``oracle/_refload.py``), runs them on seeded
synthetic inputs shaped like BASELINE.json's configs and on the edge cases the reference's tests
exercise (pyloo/tests/base_tests/test_psis.py:61-125, test_loo.py:89-171), and stores inputs and
outputs.  The reference cannot travel to the GPU box, the vectors can.  TEST INFRASTRUCTURE.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import _refload  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def create_model_loglik(seed=10):
    """Replay the legacy RNG stream of pyloo/tests/helpers.py:26-49 up to ``log_likelihood['y']``
    without ArviZ."""
    np.random.seed(seed)
    c, d, j = 4, 500, 8
    np.random.randn(c, d)          # mu
    np.random.randn(c, d)          # tau
    np.random.randn(c, d, j)       # eta
    np.random.randn(c, d, j)       # theta
    np.random.randn(c, d, j)       # posterior_predictive y
    np.random.randn(c, d)          # energy
    np.random.randn(c, d)          # diverging
    np.random.randn(c, d)          # max_depth
    return np.random.randn(c, d, j)


def student_t(rng, df, size):
    return rng.standard_t(df, size=size)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    psis, utils = _refload.load_reference()
    rng = np.random.default_rng(20261018)

    # ---- cfg 1: create_model(seed=10), loo pieces for reff 1.0 / 0.7, all via reference code
    ll = create_model_loglik(10)
    assert abs(ll[0, 0, 0] - 0.55759091) < 1e-8
    ll_sn = ll.reshape(-1, ll.shape[-1])  # (S, N) sample-major, C-order == stack(chain, draw)
    out = {"ll_sn": ll_sn}
    for tag, reff in (("r10", 1.0), ("r07", 0.7)):
        elpd_i, k, lppd_i = _refload.reference_loo_arrays(ll_sn, reff)
        lw, k2 = _refload.reference_psislw_batch(-ll_sn.T, reff)
        assert np.array_equal(k, k2)
        out[f"elpd_i_{tag}"] = elpd_i
        out[f"k_{tag}"] = k
        out[f"lppd_i_{tag}"] = lppd_i
        out[f"lw_{tag}"] = lw
    out["var_i"] = ll_sn.T.var(axis=-1)  # waic.py:145 semantics (ddof 0) on the stacked view
    np.savez_compressed(os.path.join(GOLDEN, "cfg1_create_model.npz"), **out)

    # ---- cfg 2 subset: normal log-ratios, S = 4000, reff = 0.9 (M = 200)
    x = rng.normal(size=(32, 4000))
    lw, k = _refload.reference_psislw_batch(x, 0.9)
    np.savez_compressed(os.path.join(GOLDEN, "cfg2_normal_s4000.npz"), x=x, lw=lw[:8], k=k,
                        lse_check=np.array([utils._logsumexp(r) for r in lw]), reff=0.9)

    # ---- cfg 3 subset: ll = -1.4 + z, layout (S, N), reff = 1.0 (M = 190): loo + waic pieces
    ll3 = -1.4 + rng.normal(size=(4000, 24))
    elpd_i, k, lppd_i = _refload.reference_loo_arrays(ll3, 1.0)
    np.savez_compressed(os.path.join(GOLDEN, "cfg3_loo_s4000.npz"), ll_sn=ll3, elpd_i=elpd_i,
                        k=k, lppd_i=lppd_i, var_i=ll3.var(axis=0), reff=1.0)

    # ---- cfg 4 subset: S = 16000 (M = 380), wider model
    ll4 = -1.4 - 0.3 + 1.3 * rng.normal(size=(16000, 6))
    elpd_i, k, lppd_i = _refload.reference_loo_arrays(ll4, 1.0)
    np.savez_compressed(os.path.join(GOLDEN, "cfg4_loo_s16000.npz"), ll_sn=ll4, elpd_i=elpd_i,
                        k=k, lppd_i=lppd_i, var_i=ll4.var(axis=0), reff=1.0)

    # ---- cfg 5 subset: Student-t(1.5) log-ratios, S = 8000, reff = 1.0 (M = 269)
    x5 = student_t(rng, 1.5, (24, 8000))
    lw5, k5 = _refload.reference_psislw_batch(x5, 1.0)
    elpd5, k5b, lppd5 = _refload.reference_loo_arrays(np.ascontiguousarray(-x5.T), 1.0)
    assert np.array_equal(k5, k5b, equal_nan=True)
    np.savez_compressed(os.path.join(GOLDEN, "cfg5_student_t_s8000.npz"), x=x5, lw=lw5[:4], k=k5,
                        lw_max=lw5.max(axis=1), lw_min=lw5.min(axis=1),
                        elpd_i=elpd5, lppd_i=lppd5, reff=1.0)

    # ---- edge cases (test_psis.py:95-125, test_loo.py:89-171)
    edge = {}
    edge["short4_x"] = np.array([1.0, 1.1, 1.2, 1.3])                       # tail <= 4 -> k = inf
    edge["const100_x"] = np.ones(100)                                       # constant -> -log n
    edge["len8_x"] = rng.normal(size=(5, 8))                                # rows of 8 -> k = inf
    ties = np.round(rng.normal(size=(6, 500)), 1)                            # heavy ties
    edge["ties_x"] = ties
    nanrow = rng.normal(size=(3, 400)); nanrow[1, 17] = np.nan               # NaN propagates
    edge["nan_x"] = nanrow
    big = rng.normal(size=(4, 600)); big[0, 5] = 1e10; big[1, 7] = -1e10; big[2, :3] = 1e10
    edge["big_x"] = big
    clamp = rng.normal(size=(3, 1000)) * 400.0                                # cutoffmin clamp
    edge["clamp_x"] = clamp
    small = rng.normal(size=(7, 33))
    edge["s33_x"] = small
    for name in ("short4", "const100", "len8", "ties", "nan", "big", "clamp", "s33"):
        xin = edge[f"{name}_x"]
        with np.errstate(all="ignore"):
            lw_e, k_e = _refload.reference_psislw_batch(xin, 1.0)
        edge[f"{name}_lw"] = lw_e
        edge[f"{name}_k"] = k_e
    np.savez_compressed(os.path.join(GOLDEN, "edge_cases.npz"), **edge)

    # ---- _gpdfit / _gpinv known answers
    gp = {}
    for i, n in enumerate((5, 17, 135, 200, 380)):
        t = np.sort(rng.pareto(2.0, size=n))
        k, s = psis._gpdfit(t)
        gp[f"t{i}"] = t
        gp[f"ks{i}"] = np.array([k, s])
    probs = [np.array([0.1, 0.5, 0.9]), np.array([0.0, 0.5, 1.0]), np.array([-0.1, 0.5, 1.1])]
    rows = []
    for pi, p in enumerate(probs):
        for kappa in (-1, -0.5, 0, 0.5, 1):
            for sigma in (0, 1, 2):
                with np.errstate(all="ignore"):
                    rows.append(np.concatenate([[pi, kappa, sigma], psis._gpinv(p, kappa, sigma)]))
    gp["gpinv_rows"] = np.array(rows)
    gp["gpinv_probs"] = np.array(probs)
    np.savez_compressed(os.path.join(GOLDEN, "gpd_known_answers.npz"), **gp)

    total = sum(os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN))
    print(f"golden vectors written to {GOLDEN} ({total / 1e6:.2f} MB); numpy {np.__version__}")


if __name__ == "__main__":
    main()

"""Generate tests/golden/compare.npz from the REAL reference ``loo_compare`` (build container only).

    python oracle/gen_golden_compare.py

``pyloo/compare.py`` is loaded from ``/root/reference`` by ``oracle/_refload.py`` with placeholder siblings for
``loo`` / ``waic`` / ``loo_kfold`` / ``loo_subsample`` (they need ArviZ / PyMC; ``loo_compare`` only calls them for
InferenceData inputs, pyloo/compare.py:416-448) and with the reference's own ``elpd.py`` (pure pandas).  The models
enter as precomputed, pointwise ``ELPDData`` (pyloo/compare.py:338-391): seeded synthetic ``loo_i`` / ``waic_i`` for
K = 2, 3 and 4 models, all three scales, all three weighting methods (``stacking`` :477-536, ``pseudo-bma`` :580-596,
``bb-pseudo-bma`` :539-577 with a fixed integer seed).  Stored: the inputs and the reference's whole result frame
(rank order, elpd, p, elpd_diff, weight, se, dse).  TEST INFRASTRUCTURE.
"""

from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import _refload  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
SCALES = {"log": 1.0, "negative_log": -1.0, "deviance": -2.0}
METHODS = ("stacking", "pseudo-bma", "bb-pseudo-bma")


class _Values:
    """What ``elpds[name]["loo_i"].values`` needs (the reference stores an xarray.DataArray there)."""

    def __init__(self, v):
        self.values = np.asarray(v, dtype=float)


def pointwise_models(rng, n_obs, n_models):
    """Pointwise log-scale elpds of K nested-ish models on the same observations (model k is a little worse and
    noisier), plus a p_ic column: enough structure for non-trivial stacking weights."""
    y = rng.normal(size=n_obs)
    out = []
    for k in range(n_models):
        mu = 0.25 * k * np.sin(np.arange(n_obs) * 0.37 + k)
        sd = 1.0 + 0.15 * k
        e = -0.5 * np.log(2 * np.pi * sd**2) - 0.5 * ((y - mu) / sd) ** 2 + 0.01 * rng.normal(size=n_obs)
        out.append(e)
    return out


def build_elpd(cls, ic, scale, e_log, n_samples=4000):
    sv = SCALES[scale]
    ic_i = sv * e_log
    n = ic_i.size
    total = float(ic_i.sum())
    se = float((n * np.var(ic_i)) ** 0.5)
    p = float(0.01 * n + 0.3)
    if ic == "loo":
        return cls([total, se, p, n_samples, n, False, _Values(ic_i), scale],
                   index=["elpd_loo", "se", "p_loo", "n_samples", "n_data_points", "warning", "loo_i", "scale"])
    return cls([total, se, p, n_samples, n, False, _Values(ic_i), scale],
               index=["elpd_waic", "se", "p_waic", "n_samples", "n_data_points", "warning", "waic_i", "scale"])


def main():
    placeholder = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("placeholder: not on the tested path"))  # noqa: E731
    mods = _refload.load_reference_modules(
        ("rcparams", "elpd", "compare"),
        stubs={"loo": {"loo": placeholder}, "waic": {"waic": placeholder}, "loo_kfold": {"loo_kfold": placeholder},
               "loo_subsample": {"loo_subsample": placeholder}})
    compare, elpd_mod = mods["compare"], mods["elpd"]
    rng = np.random.default_rng(20261020)
    out = {}
    cases = []
    for n_models, n_obs in ((2, 200), (3, 1000), (4, 5000)):
        e_logs = pointwise_models(rng, n_obs, n_models)
        tag = f"k{n_models}"
        for k, e in enumerate(e_logs):
            out[f"{tag}_e{k}"] = e
        for ic in ("loo", "waic"):
            for scale in SCALES:
                for method in METHODS:
                    models = {f"m{k}": build_elpd(elpd_mod.ELPDData, ic, scale, e) for k, e in enumerate(e_logs)}
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        df = compare.loo_compare(models, ic=ic, method=method, scale=scale, seed=7, b_samples=200)
                    key = f"{tag}_{ic}_{scale}_{method}"
                    cases.append(key)
                    out[key + "_order"] = np.array([int(name[1:]) for name in df.index])
                    for colname in (f"elpd_{ic}", f"p_{ic}", "elpd_diff", "weight", "se", "dse"):
                        out[key + "_" + colname.replace(f"_{ic}", "")] = df[colname].to_numpy(dtype=float)
    out["cases"] = np.array(cases)
    out["pandas_version"] = np.array(pd.__version__)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "compare.npz"), **out)
    print(f"wrote {len(cases)} cases to tests/golden/compare.npz")


if __name__ == "__main__":
    main()

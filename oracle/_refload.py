"""Load the REAL reference numerics from /root/reference -- TEST INFRASTRUCTURE ONLY.

Works only in the build container (``/root/reference`` does not exist on the GPU box).
``import pyloo`` itself fails here (xarray / arviz / pymc are absent), so the two pure-NumPy
modules on the hot path are loaded by file path with stub ``xarray`` / ``arviz`` modules, as
described in SURVEY.md Appendix B.  Nothing is copied: the functions execute from where they lie.

Used by ``oracle/gen_golden.py`` (fixture generation) and by ``tests/test_oracle_vs_reference.py``
(skipped automatically when the reference tree is absent).
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PYLOO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pyloo", "psis.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    for key, val in attrs.items():
        setattr(mod, key, val)
    return mod


def load_reference_modules(names=("utils", "psis"), stubs=None):
    """Load ``pyloo/<name>.py`` for each name (in order) from the reference tree under stub ``xarray`` /
    ``arviz`` modules and return ``{name: module}``.  Besides ``utils`` and ``psis`` this works for the
    NumPy-only numerics of ``sis``, ``tis`` and ``e_loo`` (their xarray drivers are not callable).
    ``stubs``: ``{submodule: {attribute: object}}`` registered as ``pyloo.<submodule>`` first -- placeholders for
    siblings that need ArviZ / PyMC and are imported but not called on the path under test (``compare.py``
    imports ``loo``, ``waic``, ``loo_kfold``, ``loo_subsample`` and only calls them for InferenceData inputs)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    stubs = stubs or {}
    keys = ["xarray", "arviz", "arviz.data", "pyloo"] + [f"pyloo.{n}" for n in list(names) + list(stubs)]
    saved = {k: sys.modules.get(k) for k in keys}
    try:
        class _DataArray:  # placeholder type for isinstance checks only
            pass

        class _InferenceData:
            pass

        def _apply_ufunc(func, *arrays, kwargs=None, **_ignored):
            return func(*arrays, **(kwargs or {}))

        if "xarray" not in sys.modules:
            sys.modules["xarray"] = _stub("xarray", DataArray=_DataArray, apply_ufunc=_apply_ufunc)
        if "arviz" not in sys.modules:
            az = _stub("arviz", InferenceData=_InferenceData)
            az.__path__ = []  # a package, so that ``from arviz.data import InferenceData`` resolves
            az.data = _stub("arviz.data", InferenceData=_InferenceData)
            sys.modules["arviz"] = az
            sys.modules["arviz.data"] = az.data
        pkg = types.ModuleType("pyloo")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pyloo")]
        sys.modules["pyloo"] = pkg
        for short, attrs in stubs.items():
            sys.modules[f"pyloo.{short}"] = _stub(f"pyloo.{short}", **attrs)
        mods = {}
        for short in names:
            spec = importlib.util.spec_from_file_location(
                f"pyloo.{short}", os.path.join(REFERENCE_ROOT, "pyloo", f"{short}.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"pyloo.{short}"] = mod
            spec.loader.exec_module(mod)
            mods[short] = mod
        return mods
    finally:
        for key, val in saved.items():
            if val is None:
                sys.modules.pop(key, None)
            else:
                sys.modules[key] = val


def load_reference():
    """Return ``(psis_module, utils_module)`` holding the reference's own functions."""
    mods = load_reference_modules(("utils", "psis"))
    return mods["psis"], mods["utils"]


def reference_psislw_batch(lw_ns, reff):
    """Drive the reference ``_psislw`` with the reference ``make_ufunc`` exactly as
    pyloo/psis.py:89-106 configures it.  ``lw_ns`` has samples on the last axis."""
    import numpy as np

    psis, utils = load_reference()
    work = np.array(lw_ns, dtype=np.float64, copy=True)  # psis.py:78 deepcopy
    n_samples = work.shape[-1]
    cutoff_ind = -int(np.ceil(min(n_samples / 5.0, 3 * (n_samples / reff) ** 0.5))) - 1
    cutoffmin = np.log(np.finfo(float).tiny)
    out = np.empty_like(work), np.empty(work.shape[:-1])
    ufunc = utils.make_ufunc(psis._psislw, n_dims=1, n_output=2, ravel=False, check_shape=False)
    lw, k = ufunc(work, cutoff_ind=cutoff_ind, cutoffmin=cutoffmin, out=out)
    return lw, k


def reference_loo_arrays(ll_sn, reff):
    """Reference numerics of the PSIS branch of ``loo`` on a sample-major (S, N...) array:
    pyloo/loo.py:189 (strided stack view), :227, :286-289, :319-337.  Host glue restated,
    numerics executed by the reference's own ``_psislw`` / ``_logsumexp`` / ``make_ufunc``."""
    import numpy as np

    psis, utils = load_reference()
    ll = np.moveaxis(np.asarray(ll_sn, dtype=np.float64), 0, -1)  # strided view like .stack()
    if np.any(np.isnan(ll)):
        ll = np.where(np.isnan(ll), -1e10, ll)
    n_samples = ll.shape[-1]
    lw, k = reference_psislw_batch(-ll, reff)
    lw += ll
    lse = utils.make_ufunc(utils._logsumexp, n_dims=1, ravel=False)
    elpd_i = lse(lw)
    lppd_i = lse(ll, b_inv=n_samples)
    return elpd_i, k, lppd_i

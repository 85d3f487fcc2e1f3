"""Data ingress for the hot path (reference: pyloo/utils.py:21-79 ``to_inference_data``,
:257-302 ``get_log_likelihood``; pyloo/loo.py:189 ``stack(__sample__=("chain","draw"))``).

ArviZ / xarray objects are used when those packages are importable.  Neither is required: the
``Lite*`` classes below are a minimal NumPy-backed stand-in with the attributes the path touches
(``.log_likelihood`` / ``.posterior`` groups, ``.data_vars``, ``.dims``, ``.values``), so the
engine -- and its tests -- run in an image without ArviZ.
"""

from __future__ import annotations

import warnings

import numpy as np

__all__ = ["LiteDataArray", "LiteDataset", "InferenceDataLite", "from_dict", "to_inference_data",
           "get_log_likelihood", "sample_major", "wrap_like", "is_dataarray_like"]

SAMPLE_DIM = "__sample__"


class LiteDataArray:
    """Array with named dimensions (the slice of ``xarray.DataArray`` the hot path relies on)."""

    def __init__(self, values, dims, name=None, coords=None):
        self.values = np.asarray(values)
        self.dims = tuple(dims)
        if len(self.dims) != self.values.ndim:
            raise ValueError("dims must match the array rank")
        self.name = name
        self.coords = dict(coords or {})

    shape = property(lambda self: self.values.shape)
    ndim = property(lambda self: self.values.ndim)
    dtype = property(lambda self: self.values.dtype)
    sizes = property(lambda self: dict(zip(self.dims, self.values.shape)))

    def __len__(self):
        return self.values.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __getattr__(self, item):
        # xarray exposes dimensions as attributes: the reference tests ``hasattr(x, "__sample__")``
        # (pyloo/psis.py:79, pyloo/base.py:114)
        dims = object.__getattribute__(self, "dims")
        if item in dims:
            return np.arange(object.__getattribute__(self, "values").shape[dims.index(item)])
        raise AttributeError(item)

    def __neg__(self):
        return LiteDataArray(-self.values, self.dims, self.name, self.coords)

    def __getitem__(self, key):
        vals = self.values[key]
        if np.ndim(vals) == 0:
            return vals
        if isinstance(key, (int, np.integer)):
            return LiteDataArray(vals, self.dims[1:], self.name)
        return LiteDataArray(vals, self.dims[-np.ndim(vals):], self.name)

    def rename(self, name):
        return LiteDataArray(self.values, self.dims, name, self.coords)

    def copy(self):
        return LiteDataArray(self.values.copy(), self.dims, self.name, self.coords)

    def transpose(self, *dims):
        if Ellipsis in dims:
            head = [d for d in dims if d is not Ellipsis]
            dims = tuple(head) + tuple(d for d in self.dims if d not in head)
        perm = [self.dims.index(d) for d in dims]
        return LiteDataArray(self.values.transpose(perm), dims, self.name, self.coords)

    def stack(self, **kw):
        """``stack(__sample__=("chain","draw"))``: the stacked dims move last, like xarray."""
        (new, old), = kw.items()
        keep = [d for d in self.dims if d not in old]
        arr = self.transpose(*keep, *old).values
        n_new = int(np.prod(arr.shape[len(keep):]))
        return LiteDataArray(arr.reshape(*arr.shape[:len(keep)], n_new), (*keep, new), self.name)

    def isel(self, indexers):
        idx = tuple(indexers.get(d, slice(None)) for d in self.dims)
        dims = tuple(d for d in self.dims if not isinstance(indexers.get(d, slice(None)), (int, np.integer)))
        return LiteDataArray(self.values[idx], dims, self.name)

    def sum(self, dim=None):
        if dim is None:
            return self.values.sum()
        ax = self.dims.index(dim)
        return LiteDataArray(self.values.sum(axis=ax), self.dims[:ax] + self.dims[ax + 1:], self.name)

    def __repr__(self):
        return f"LiteDataArray(name={self.name!r}, dims={self.dims}, shape={self.shape})"


class LiteDataset:
    """Named collection of :class:`LiteDataArray` (one InferenceData group)."""

    def __init__(self, arrays):
        self._arrays = dict(arrays)

    @property
    def data_vars(self):
        return self._arrays

    def __getitem__(self, key):
        return self._arrays[key]

    def __contains__(self, key):
        return key in self._arrays

    def __getattr__(self, item):
        arrays = object.__getattribute__(self, "_arrays")
        if item in arrays:
            return arrays[item]
        if item in ("chain", "draw") and arrays:
            first = next(iter(arrays.values()))
            return getattr(first, item)
        raise AttributeError(item)

    def keys(self):
        return self._arrays.keys()


class InferenceDataLite:
    """Groups as attributes, like ``arviz.InferenceData``."""

    def __init__(self, **groups):
        self._groups = list(groups)
        for name, ds in groups.items():
            setattr(self, name, ds)

    def groups(self):
        return list(self._groups)


def from_dict(posterior=None, log_likelihood=None, dims=None, **other):
    """Build an :class:`InferenceDataLite` from ``{var: array(chain, draw, *shape)}`` dictionaries
    (the subset of ``arviz.from_dict`` the reference's NumPy-only fixtures use,
    pyloo/tests/helpers.py:64-84)."""
    dims = dims or {}

    def group(block, sampled=True):
        out = {}
        for var, arr in block.items():
            arr = np.asarray(arr)
            if sampled:
                extra = dims.get(var) or [f"{var}_dim_{i}" for i in range(arr.ndim - 2)]
                out[var] = LiteDataArray(arr, ("chain", "draw", *extra), name=var)
            else:  # observed_data / constant_data carry no chain / draw dimensions
                out[var] = LiteDataArray(arr, dims.get(var) or [f"{var}_dim_{i}" for i in range(arr.ndim)], name=var)
        return LiteDataset(out)

    groups = {}
    if posterior is not None:
        groups["posterior"] = group(posterior)
    if log_likelihood is not None:
        groups["log_likelihood"] = group(log_likelihood)
    for name, block in other.items():
        if isinstance(block, dict):
            groups[name] = group(block, sampled=name not in ("observed_data", "constant_data"))
    return InferenceDataLite(**groups)


def _arviz():
    try:
        import arviz  # type: ignore

        return arviz
    except Exception:  # pragma: no cover - arviz is absent in the build image
        return None


def to_inference_data(obj):
    """pyloo/utils.py:21-79: pass InferenceData through, refuse lists/tuples and ragged dicts, otherwise
    defer to ``arviz.convert_to_inference_data`` (when ArviZ is installed)."""
    if isinstance(obj, InferenceDataLite) or hasattr(obj, "log_likelihood") or hasattr(obj, "posterior"):
        return obj
    az = _arviz()
    if az is not None and isinstance(obj, az.InferenceData):
        return obj
    if isinstance(obj, (list, tuple)):
        raise ValueError("Lists and tuples cannot be converted to InferenceData directly")
    if isinstance(obj, dict) and not all(isinstance(v, (np.ndarray, list)) for v in obj.values()):
        raise ValueError("Dictionary values must be array-like")
    if az is not None:
        try:
            return az.convert_to_inference_data(obj)
        except Exception as err:
            raise ValueError(f"Can only convert ArviZ-supported objects to InferenceData, not "
                             f"{obj.__class__.__name__}") from err
    if isinstance(obj, dict):
        return from_dict(posterior=obj)       # arviz puts a bare dict into the posterior group
    if isinstance(obj, np.ndarray):
        return from_dict(posterior={"x": obj})
    raise ValueError(f"Can only convert InferenceData-like objects, dict or numpy array to "
                     f"InferenceData, not {obj.__class__.__name__}")


def get_log_likelihood(idata, var_name=None, single_var=True):
    """pyloo/utils.py:257-302 (same errors: ``TypeError`` when absent / ambiguous / unknown name)."""
    if (not hasattr(idata, "log_likelihood") and hasattr(idata, "sample_stats")
            and hasattr(idata.sample_stats, "log_likelihood")):
        warnings.warn("Storing the log_likelihood in sample_stats groups has been deprecated",
                      DeprecationWarning, stacklevel=2)
        return idata.sample_stats.log_likelihood
    if not hasattr(idata, "log_likelihood"):
        raise TypeError("log likelihood not found in inference data object")
    group = idata.log_likelihood
    if var_name is None:
        names = list(group.data_vars)
        if len(names) > 1:
            if single_var:
                raise TypeError(f"Found several log likelihood arrays {names}, var_name cannot be None")
            return group[names]
        return group[names[0]]
    try:
        return group[var_name]
    except KeyError as err:
        raise TypeError(f"No log likelihood data named {var_name} found") from err


def is_dataarray_like(obj) -> bool:
    return hasattr(obj, "dims") and hasattr(obj, "values") and not isinstance(obj, np.ndarray)


def sample_major(da):
    """Log-likelihood with ``chain`` / ``draw`` dims -> ``(values_sn, obs_dims, obs_shape, stacked)``.

    ``values_sn`` is the ``(S, N)`` sample-major float64 matrix whose row order equals
    ``stack(__sample__=("chain","draw"))`` (chain outer, draw inner; pyloo/loo.py:189).  For the ArviZ
    layout ``(chain, draw, obs...)`` it is a reshape *view* -- no copy of the big array."""
    dims = tuple(da.dims)
    if "chain" not in dims or "draw" not in dims:
        raise ValueError("log likelihood must have chain and draw dimensions")
    obs_dims = tuple(d for d in dims if d not in ("chain", "draw"))
    vals = np.asarray(da.values)
    perm = [dims.index("chain"), dims.index("draw")] + [dims.index(d) for d in obs_dims]
    if perm != list(range(len(dims))):
        vals = vals.transpose(perm)
    obs_shape = vals.shape[2:]
    S = vals.shape[0] * vals.shape[1]
    N = int(np.prod(obs_shape)) if obs_shape else 1
    if vals.dtype != np.float64:
        vals = vals.astype(np.float64)
    mat = vals.reshape(S, N)  # view when C-contiguous, copy otherwise (host glue)
    return mat, obs_dims, obs_shape


def wrap_like(template, values, dims, name):
    """Return ``values`` wrapped in the same array family as ``template`` (xarray -> xarray,
    Lite -> Lite), named ``name`` (pyloo/psis.py:107-110)."""
    if isinstance(template, LiteDataArray):
        return LiteDataArray(values, dims, name=name)
    try:
        import xarray as xr  # type: ignore

        if isinstance(template, xr.DataArray):
            coords = {d: template.coords[d] for d in dims if d in template.coords}
            try:
                return xr.DataArray(values, dims=dims, coords=coords, name=name)
            except Exception:
                return xr.DataArray(values, dims=dims, name=name)
    except ImportError:  # pragma: no cover
        pass
    return LiteDataArray(values, dims, name=name)

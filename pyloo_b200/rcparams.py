"""The three validated defaults the hot path reads (reference: pyloo/rcparams.py:30-34,122).

``stats.ic_pointwise`` and ``stats.ic_scale`` are consulted by :func:`loo` / :func:`waic`
(pyloo/loo.py:181,193; pyloo/waic.py:93,99).  Keys cannot be added or removed.
"""

from collections.abc import MutableMapping

_SCALES = ("deviance", "log", "negative_log")


def _as_bool(value):
    if isinstance(value, bool):
        return value
    raise ValueError(f"Value must be True or False, not {value}")


def _as_scale(value):
    if isinstance(value, str) and value.lower() in _SCALES:
        return value.lower()
    raise ValueError(f"Scale must be one of {set(_SCALES)}, not {value}")


def _as_backend(value):
    if isinstance(value, str) and value.lower() == "matplotlib":
        return "matplotlib"
    raise ValueError(f"Backend must be one of {{'matplotlib'}}, not {value}")


_SPEC = {
    "stats.ic_pointwise": (False, _as_bool),
    "stats.ic_scale": ("log", _as_scale),
    "plot.backend": ("matplotlib", _as_backend),
}


class RcParams(MutableMapping):
    """Validated, fixed-key mapping of defaults."""

    def __init__(self, **overrides):
        self._store = {key: default for key, (default, _) in _SPEC.items()}
        for key, val in overrides.items():
            self[key] = val

    def __setitem__(self, key, val):
        if key not in _SPEC:
            raise KeyError(f"{key} is not a valid rc parameter (see rcParams.keys() for a list of valid parameters)")
        try:
            self._store[key] = _SPEC[key][1](val)
        except ValueError as err:
            raise ValueError(f"Key {key}: {err}") from err

    def __getitem__(self, key):
        return self._store[key]

    def _frozen(self, *_args, **_kwargs):
        raise TypeError("RcParams keys cannot be deleted")

    __delitem__ = clear = pop = popitem = _frozen

    def setdefault(self, key, default=None):
        raise TypeError("Defaults in RcParams are handled on object initialization.")

    def __iter__(self):
        return iter(sorted(self._store))

    def __len__(self):
        return len(self._store)

    def copy(self):
        return dict(self._store)

    def __repr__(self):
        return f"RcParams({self._store})"

    def __str__(self):
        return "\n".join(f"{k:<22}: {v}" for k, v in sorted(self._store.items()))


rcParams = RcParams()

"""ELPD result record -- the output contract of ``loo`` / ``waic`` (reference: pyloo/elpd.py:100-498).

A ``pandas.Series`` subclass whose index carries the same row names, in the same order, as the
reference (pyloo/loo.py:554-577, :599-624; pyloo/waic.py:163-207) and whose ``str()`` is the same
report: estimate table, warning line and the Pareto-k table with bins ``(-inf, good_k]``,
``(good_k, 1]``, ``(1, inf)`` (pyloo/elpd.py:300-330).  Only the kinds on the hot path (loo, waic)
are rendered here.
"""

from copy import copy as _shallow, deepcopy as _deep

import numpy as np
import pandas as pd

_HEAD = "\nComputed from {n_samples} posterior samples and {n_points} observations log-likelihood matrix.\n"
_HEAD_LOGO = "\nComputed from {n_samples} posterior samples and {n_groups} groups log-likelihood matrix.\n"
_TABLE = {
    "loo": ("elpd_loo", "p_loo", "looic"),
    "waic": ("elpd_waic", "p_waic", None),
    "logo": ("elpd_logo", "p_logo", "logoic"),  # pyloo/elpd.py:74-81, :165-222
}
_K_TABLE = (
    "\n------\n\nPareto k diagnostic values:\n"
    "                         Count   Pct.\n"
    "(-Inf, {gk:.2f}]   (good)      {c0:d}   {p0:.1f}%\n"
    "   ({gk:.2f}, 1]   (bad)         {c1:d}    {p1:.1f}%\n"
    "   (1, Inf)   (very bad)    {c2:d}    {p2:.1f}%"
)
_WARN = "\n\nThere has been a warning during the calculation. Please check the results."


def _k_counts(k_values, good_k):
    """Bin counts with numpy.histogram edge semantics (pyloo/elpd.py:305-311, :501-505)."""
    vals = np.asarray(getattr(k_values, "values", k_values), dtype=float).ravel()
    counts, _ = np.histogram(vals, bins=np.asarray([-np.inf, good_k, 1, np.inf]))
    return counts


class ELPDData(pd.Series):
    """Series of ELPD estimates and diagnostics with a printable report."""

    @property
    def _constructor(self):
        return ELPDData

    def _kind(self):
        kind = str(self.index[0]).split("_")[1]
        if kind not in _TABLE:
            raise ValueError("Invalid ELPDData object")
        return kind

    def __str__(self):
        kind = self._kind()
        est, pen, ic = _TABLE[kind]
        if kind == "logo":
            out = _HEAD_LOGO.format(n_samples=self["n_samples"], n_groups=self["n_groups"])
            out += "\n         Estimate       SE\n"
            out += f"{est}   {self[est]:<8.2f}    {self['se']:<.2f}\n"
            out += f"{pen}       {self[pen]:<8.2f}    {self.get('p_logo_se', float('nan')):<.2f}\n"
            out += f"{ic}      {self['logoic']:<8.2f}    {self['logoic_se']:<.2f}"
            if self["warning"]:
                out += _WARN
            if "pareto_k" in self and self.get("good_k") is not None:
                good_k = self["good_k"]
                counts = _k_counts(self["pareto_k"], good_k)
                if counts[1] == 0 and counts[2] == 0:
                    out += (f"\n\nAll Pareto k estimates are good (k < {good_k:.1f})."
                            "\nSee help('pareto-k-diagnostic') for details.")
                else:
                    pct = counts / np.sum(counts) * 100
                    out += _K_TABLE.format(gk=good_k, c0=int(counts[0]), c1=int(counts[1]), c2=int(counts[2]),
                                           p0=pct[0], p1=pct[1], p2=pct[2])
            return out
        out = _HEAD.format(n_samples=self["n_samples"], n_points=self["n_data_points"])
        out += "\n         Estimate       SE\n"
        out += f"{est}   {self[est]:<8.2f}    {self['se']:<.2f}\n"
        if kind == "loo":
            out += f"{pen}       {self[pen]:<8.2f}    {self['p_loo_se']:<.2f}\n"
            out += f"{ic}      {self['looic']:<8.2f}    {self['looic_se']:<.2f}"
        else:
            out += f"{pen}       {self[pen]:<8.2f}    -"
        if self["warning"]:
            out += _WARN
        if kind == "loo":
            good_k = self["good_k"] if "good_k" in self else None
            if "pareto_k" in self and good_k is not None:
                counts = _k_counts(self["pareto_k"], good_k)
                if counts[1] == 0 and counts[2] == 0:
                    out += (f"\n\nAll Pareto k estimates are good (k < {good_k:.1f})."
                            "\nSee help('pareto-k-diagnostic') for details.")
                else:
                    pct = counts / np.sum(counts) * 100
                    out += _K_TABLE.format(gk=good_k, c0=int(counts[0]), c1=int(counts[1]), c2=int(counts[2]),
                                           p0=pct[0], p1=pct[1], p2=pct[2])
            elif self["warning"]:
                out += ("\n\nSome Pareto k diagnostic values are high (k > 0.7), indicating that the"
                        " importance sampling approximation is unreliable. Consider using moment matching"
                        " or exact LOO for more accurate estimates. Use pointwise=True to see detailed"
                        " diagnostics.")
            else:
                out += "\n\nAll Pareto k estimates are good (k < 0.7).\nSee help('pareto-k-diagnostic') for details."
        return out

    __repr__ = __str__

    def copy(self, deep=True):
        dup = pd.Series.copy(self)
        fn = _deep if deep else _shallow
        for key in dup.keys():
            dup[key] = fn(dup[key])
        return ELPDData(dup)

    @property
    def n_samples(self):
        return self["n_samples"]

    @property
    def n_data_points(self):
        return self["n_data_points"]

    @property
    def n_groups(self):
        return self["n_groups"]

    @property
    def warning(self):
        return self["warning"]

    @property
    def method(self):
        return getattr(self, "_method", "psis")

    @method.setter
    def method(self, value):
        object.__setattr__(self, "_method", value)

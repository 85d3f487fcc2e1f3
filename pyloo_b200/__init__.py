"""pyloo_b200 -- B200-native PSIS-LOO engine behind pyloo's call signatures.

Drop-in for ``pl.psislw`` / ``pl.sislw`` / ``pl.tislw``, ``pl.compute_importance_weights``,
``pl.loo(method="psis" | "sis" | "tis")``, ``pl.waic``, ``pl.loo_compare``, ``pl.e_loo``,
``pl.loo_predictive_metric`` and ``pl.loo_score`` (reference: jordandeklerk/pyloo).  The numerics run in hand-written
sm_100a CUDA kernels behind a C ABI (``include/psisloo_b200.h``); there is no CPU fallback.
"""

from .rcparams import rcParams  # noqa: F401
from .elpd import ELPDData  # noqa: F401
from .base import ISMethod, compute_importance_weights  # noqa: F401
from .psis import psislw  # noqa: F401
from .sis import sislw  # noqa: F401
from .tis import tislw  # noqa: F401
from .e_loo import ExpectationResult, compute_pareto_k, e_loo, k_hat  # noqa: F401
from .loo import loo  # noqa: F401
from .loo_group import loo_group  # noqa: F401
from .loo_predictive_metric import loo_predictive_metric  # noqa: F401
from .loo_score import LooScoreResult, loo_score  # noqa: F401
from .waic import waic  # noqa: F401
from .compare import loo_compare  # noqa: F401
from .data import InferenceDataLite, LiteDataArray, from_dict  # noqa: F401

__version__ = "0.1.0"

__all__ = [
    "psislw", "sislw", "tislw", "e_loo", "ExpectationResult", "compute_pareto_k", "k_hat",
    "compute_importance_weights", "ISMethod", "loo", "loo_group", "loo_predictive_metric", "loo_score", "LooScoreResult", "waic", "loo_compare", "ELPDData",
    "rcParams", "InferenceDataLite", "LiteDataArray", "from_dict",
]

"""Estimators with the constructor / fit / predict API of ``sparselm.model``
(reference: src/sparselm/model/__init__.py:26-43), solved by the B200 engine.

The convex estimators are implemented; the MIQP estimators need a mixed-integer
solver and are declared out of scope: their names exist and raise at ``fit``.
"""

from ._adaptive_lasso import (
    AdaptiveGroupLasso,
    AdaptiveLasso,
    AdaptiveOverlapGroupLasso,
    AdaptiveRidgedGroupLasso,
    AdaptiveSparseGroupLasso,
)
from ._lasso import (
    GroupLasso,
    Lasso,
    OverlapGroupLasso,
    RidgedGroupLasso,
    SparseGroupLasso,
)
from ._ols import OrdinaryLeastSquares
from ._miqp import (
    L1L0,
    L2L0,
    BestSubsetSelection,
    RegularizedL0,
    RidgedBestSubsetSelection,
)

__all__ = [
    "OrdinaryLeastSquares",
    "Lasso",
    "BestSubsetSelection",
    "RidgedBestSubsetSelection",
    "RegularizedL0",
    "L1L0",
    "L2L0",
    "GroupLasso",
    "OverlapGroupLasso",
    "SparseGroupLasso",
    "RidgedGroupLasso",
    "AdaptiveLasso",
    "AdaptiveGroupLasso",
    "AdaptiveOverlapGroupLasso",
    "AdaptiveSparseGroupLasso",
    "AdaptiveRidgedGroupLasso",
]

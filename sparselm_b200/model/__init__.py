"""Estimators with the constructor / fit / predict API of ``sparselm.model``
(reference: src/sparselm/model/__init__.py:26-43 lists the same public names), solved by the B200 engine.

Three families:
  * convex, on the engine's hot path (penalised least squares, one batch of proximal-gradient columns);
  * adaptive (iteratively re-weighted) variants of the convex ones, same path, several passes;
  * mixed-integer (MIQP) estimators: they need a mixed-integer solver and are out of scope --
    the names exist so that imports keep working, and raise at ``fit`` (no CPU fallback).
``OrdinaryLeastSquares`` is the unpenalised member (conjugate gradients on the Gram).
"""

from ._adaptive_lasso import AdaptiveGroupLasso, AdaptiveLasso, AdaptiveOverlapGroupLasso  # noqa: I001
from ._adaptive_lasso import AdaptiveRidgedGroupLasso, AdaptiveSparseGroupLasso
from ._lasso import GroupLasso, Lasso, OverlapGroupLasso, RidgedGroupLasso, SparseGroupLasso
from ._miqp import L1L0, L2L0, BestSubsetSelection, RegularizedL0, RidgedBestSubsetSelection
from ._ols import OrdinaryLeastSquares

CONVEX = ("Lasso", "GroupLasso", "OverlapGroupLasso", "SparseGroupLasso", "RidgedGroupLasso")
ADAPTIVE = tuple("Adaptive" + name for name in CONVEX)
MIQP = ("BestSubsetSelection", "RidgedBestSubsetSelection", "RegularizedL0", "L1L0", "L2L0")

# same order as the reference's __all__ (OLS, Lasso, the MIQP block, group estimators, adaptive estimators)
__all__ = ["OrdinaryLeastSquares", CONVEX[0], *MIQP, *CONVEX[1:], *ADAPTIVE]

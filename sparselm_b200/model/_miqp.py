"""MIQP estimators of the reference (src/sparselm/model/_miqp/): out of scope.

They need a mixed-integer solver (Gurobi/SCIP); BASELINE.json's north_star
declares them out of scope with no CPU fallback.  The names are kept so that
``from sparselm_b200.model import L2L0`` does not break imports; using them
raises.
"""

from __future__ import annotations

from sklearn.base import BaseEstimator, RegressorMixin


class _OutOfScopeMIQP(RegressorMixin, BaseEstimator):
    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def fit(self, X, y, sample_weight=None):
        raise NotImplementedError(
            f"{type(self).__name__} is a mixed-integer (MIQP) estimator: out of scope of the "
            "B200 engine (convex estimators only) and there is no CPU fallback"
        )


class BestSubsetSelection(_OutOfScopeMIQP):
    pass


class RidgedBestSubsetSelection(_OutOfScopeMIQP):
    pass


class RegularizedL0(_OutOfScopeMIQP):
    pass


class L1L0(_OutOfScopeMIQP):
    pass


class L2L0(_OutOfScopeMIQP):
    pass

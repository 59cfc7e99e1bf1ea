"""Piece-wise (stepwise) fitting: a chain of estimators, each fitting its own slice of the
feature matrix to the residual left by the ones before it (reference: src/sparselm/stepwise.py).

Every step is an engine-backed estimator or a GridSearchCV / LineSearchCV over one, so each
link of the chain is a device solve; the residual chaining itself is host arithmetic on
``n`` numbers per step.
"""

from __future__ import annotations

import numpy as np
from sklearn.base import RegressorMixin
from sklearn.utils._param_validation import InvalidParameterError
from sklearn.utils.metaestimators import _BaseComposition
from sklearn.utils.validation import _check_sample_weight, check_is_fitted, validate_data

__all__ = ["StepwiseEstimator"]


def _inner(estimator):
    """The regressor behind a step: a searcher (GridSearchCV / LineSearchCV) wraps one."""
    return estimator.estimator if hasattr(estimator, "estimator") else estimator


def _fitted(estimator):
    check_is_fitted(estimator)
    model = estimator.best_estimator_ if hasattr(estimator, "best_estimator_") else estimator
    if not hasattr(model, "coef_"):
        raise ValueError(f"Estimator {estimator} is not a valid linear model!")
    return model


def _scopes_partition_features(scopes):
    """True iff the scopes are disjoint and together are exactly 0 .. n_features-1
    (reference stepwise.py:19-21)."""
    flat = [int(i) for scope in scopes for i in scope]
    return sorted(flat) == list(range(len(flat)))


class StepwiseEstimator(_BaseComposition, RegressorMixin):
    """Composite estimator for stepwise fitting (reference stepwise.py:44-237).

    Args:
        steps (list[(str, estimator)]): named estimators, fitted in order; step ``i`` fits
            ``X[:, estimator_feature_indices[i]]`` to the residual of steps ``< i``.  A step may
            be an estimator of ``sparselm_b200.model`` or a GridSearchCV / LineSearchCV over
            one; it may not be another StepwiseEstimator, and only the first step may fit an
            intercept.
        estimator_feature_indices (tuple[tuple[int]]): the feature indices of every step;
            disjoint, and together ``0 .. n_features-1``.  Group labels etc. of a step refer to
            its own slice.
    """

    def __init__(self, steps, estimator_feature_indices):
        self.steps = steps
        self.estimator_feature_indices = estimator_feature_indices

    def get_params(self, deep=True):
        """Parameters of the composite and (deep) of every step, ``<step>__<param>``."""
        return self._get_params("steps", deep=deep)

    def set_params(self, **params):
        """Set parameters of the steps (``<step>__<param>=value``) or replace whole steps."""
        self._set_params("steps", **params)
        return self

    def _check_structure(self):
        scopes = self.estimator_feature_indices
        if not _scopes_partition_features(scopes):
            raise InvalidParameterError(
                f"Given feature indices: {scopes} are not continuous and non-overlapping series starting from 0!")
        if any(isinstance(est, StepwiseEstimator) for _, est in self.steps):
            raise InvalidParameterError("StepwiseEstimator should not be nested with another StepwiseEstimator!")
        if any(getattr(_inner(est), "fit_intercept", False) for _, est in self.steps[1:]):
            raise InvalidParameterError("Only the first estimator in steps is allowed to fit intercept!")
        if len(self.steps) != len(scopes):
            raise InvalidParameterError("One scope of feature indices is needed per step!")

    def fit(self, X, y, sample_weight=None, *args, **kwargs):
        """Fit the chain (reference stepwise.py:148-237): step i sees the residual of the
        steps before it; extra arguments go to every step's ``fit``."""
        self._check_structure()
        scopes = [[int(i) for i in scope] for scope in self.estimator_feature_indices]
        # the number of features is fixed by the scopes, so X is checked against it (a
        # mismatch is a ValueError, stepwise.py:195-205)
        self.n_features_in_ = sum(len(s) for s in scopes)
        X, y = validate_data(self, X, y, accept_sparse=False, ensure_2d=True, y_numeric=True, multi_output=True,
                             reset=False, dtype=np.float64)
        if sample_weight is not None:
            sample_weight = _check_sample_weight(sample_weight, X, dtype=X.dtype)

        residual = np.array(y, dtype=np.float64, copy=True)
        coef = np.full(X.shape[1], np.nan)
        for (_, est), scope in zip(self.steps, scopes):
            Xs = X[:, scope]
            est.fit(Xs, residual, *args, sample_weight=sample_weight, **kwargs)
            coef[scope] = np.array(_fitted(est).coef_, copy=True)
            residual = residual - est.predict(Xs)
        self.coef_ = coef
        first = self.steps[0][1]
        self.intercept_ = _fitted(first).intercept_ if _inner(first).fit_intercept else 0.0
        return self

    def predict(self, X):
        check_is_fitted(self, "coef_")
        X = validate_data(self, X, accept_sparse=False, reset=False, dtype=np.float64)
        return X @ self.coef_ + self.intercept_

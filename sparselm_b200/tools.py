"""Helpers around a fit (reference: src/sparselm/tools.py): box-constrained re-fit of
selected coefficients and the r2 -> CV error conversion.  Host-side glue: the fits they wrap
are the engine-backed estimators."""

from __future__ import annotations

import warnings
from functools import wraps
from numbers import Real

import numpy as np

__all__ = ["constrain_coefficients", "r2_score_to_cv_error"]


def _bound(value, n, default):
    if value is None:
        return np.full(n, default)
    if isinstance(value, Real):
        return np.full(n, float(value))
    return np.asarray(value, dtype=float)


def constrain_coefficients(indices, high=None, low=None):
    """Decorator keeping selected coefficients of a fit inside ``[low, high]``
    (reference tools.py:14-103).

    ``fit_method(X, y, *args, **kwargs) -> coefs`` is run once; coefficients listed in
    ``indices`` that left their range are pinned to the violated bound -- their columns'
    contribution at that value is moved into the target and the columns are zeroed -- and the
    fit is run once more.  A RuntimeWarning tells when the second fit pushed other listed
    coefficients out of range.

    Args:
        indices (array-like[int]): coefficients to constrain.
        high, low (float | array-like | None): bounds, scalar or one per index.
    """
    idx = np.asarray(indices, dtype=int)
    hi = _bound(high, len(idx), np.inf)
    lo = _bound(low, len(idx), -np.inf)

    def decorate(fit_method):
        @wraps(fit_method)
        def constrained(X, y, *args, **kwargs):
            coefs = fit_method(X, y, *args, **kwargs)
            over, under = coefs[idx] > hi, coefs[idx] < lo
            if over.any() or under.any():
                pinned = np.concatenate([idx[over], idx[under]])
                values = np.concatenate([hi[over], lo[under]])
                Xc = np.array(X, dtype=float, copy=True)
                yc = np.array(y, dtype=float, copy=True) - Xc[:, pinned] @ values
                Xc[:, pinned] = 0.0
                coefs = fit_method(Xc, yc, *args, **kwargs)
                coefs[pinned] = values
            if (coefs[idx] > hi).any() or (coefs[idx] < lo).any():
                warnings.warn(
                    "Running the constrained fit has resulted in new out of range coefficients that were not so "
                    "in the unconstrained fit.\nDouble check the sensibility of the bounds you provided!",
                    RuntimeWarning)
            return coefs

        return constrained

    return decorate


def r2_score_to_cv_error(score, y, y_pred, weights=None):
    """CV error from an r2 score (reference tools.py:106-131):
    ``sqrt((1 - score) * sum_i w_i (y_i - y_pred_i)^2 / sum_i w_i)``."""
    y = np.asarray(y, dtype=float)
    w = np.ones(len(y)) if weights is None else np.asarray(weights, dtype=float)
    if len(w) != len(y):
        raise ValueError("Weights given but not the same length as sample.")
    if np.any(w < 0) or np.allclose(w, 0):
        raise ValueError("Weights can not be negative or all zero.")
    msd = float((w * (y - np.asarray(y_pred, dtype=float)) ** 2).sum() / w.sum())
    return np.sqrt((1 - score) * msd)

"""Hyper-parameter validation with the reference's error contract
(reference: src/sparselm/_utils/validation.py:9-59, pinned by tests/test_lasso.py:203-260)."""

from __future__ import annotations

import numpy as np


def _check_groups(groups, n_features: int) -> None:
    """groups must be None, or a 1-D list/ndarray with one label per feature."""
    if groups is None:
        return
    if not isinstance(groups, (list, np.ndarray)):
        raise TypeError("groups must be a list or ndarray")
    arr = np.asarray(groups).astype(int)
    if arr.ndim != 1:
        raise ValueError("groups must be a 1D array")
    if len(arr) != n_features:
        raise ValueError(f"groups must be the same length as the number of features {n_features}")


def _check_group_weights(group_weights, n_groups: int) -> None:
    """group_weights must be None, or a list/ndarray with one weight per group."""
    if group_weights is None:
        return
    if not isinstance(group_weights, (list, np.ndarray)):
        raise TypeError("group_weights must be a list or ndarray")
    arr = np.asarray(group_weights)
    if len(arr) != n_groups:
        raise ValueError(
            f"group_weights must be the same length as the number of groups {len(arr)} != {n_groups}"
        )

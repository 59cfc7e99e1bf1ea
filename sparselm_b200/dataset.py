"""Synthetic grouped regression problems (reference: src/sparselm/dataset.py) -- the
generator behind the group-structured test and benchmark designs."""

from __future__ import annotations

import warnings
from numbers import Integral

import numpy as np
from sklearn.datasets import make_regression
from sklearn.utils import check_random_state

__all__ = ["make_group_regression"]


def make_group_regression(n_samples=100, n_groups=20, n_features_per_group=10, n_informative_groups=5,
                          frac_informative_in_group=1.0, bias=0.0, effective_rank=None, tail_strength=0.5,
                          noise=0.0, shuffle=True, coef=False, random_state=None):
    """Random regression problem whose informative features sit in the first
    ``n_informative_groups`` groups (reference dataset.py:15-139).

    The design, target and ground-truth coefficients come from
    ``sklearn.datasets.make_regression`` (same generator, same draw order as the reference, so
    a seed gives the same problem); features with a coefficient above ``noise`` are dealt to
    the informative groups, ``round(frac_informative_in_group * size)`` per group, the rest
    fill the groups up; with ``shuffle`` the columns (and their labels) are permuted.

    Returns:
        (X, y, groups) or (X, y, groups, coefs) if ``coef``.
    """
    rng = check_random_state(random_state)
    if isinstance(n_features_per_group, Integral):
        sizes = [int(n_features_per_group)] * n_groups
    else:
        sizes = [int(s) for s in n_features_per_group]
        if len(sizes) != n_groups:
            raise ValueError("If passing a sequence of n_features_per_group, the length must be equal to n_groups.")
    n_inf = [round(frac_informative_in_group * sizes[g]) for g in range(n_informative_groups)]
    if any(k < 1 for k in n_inf):
        warnings.warn("The number of features and fraction of informative features per group resulted in "
                      "informative groups having no informative features.", UserWarning)

    X, y, w = make_regression(n_samples=n_samples, n_features=sum(sizes), n_informative=sum(n_inf), bias=bias,
                              effective_rank=effective_rank, tail_strength=tail_strength, noise=noise,
                              shuffle=shuffle, coef=True, random_state=rng)

    strong = list(np.flatnonzero(w > noise))
    weak = list(np.flatnonzero(w <= noise))
    groups = np.zeros(sum(sizes), dtype=int)
    for g, size in enumerate(sizes):
        k = n_inf[g] if g < n_informative_groups else 0
        members, strong = strong[:k], strong[k:]
        fill, weak = weak[:size - k], weak[size - k:]
        groups[members + fill] = g

    if shuffle:
        order = np.arange(sum(sizes))
        rng.shuffle(order)
        X[:, :] = X[:, order]
        groups, w = groups[order], w[order]
    return (X, y, groups, w) if coef else (X, y, groups)

"""Lasso-family estimators (reference: src/sparselm/model/_lasso.py).

NOTE on the objective: the reference *code* minimises
``1/(2n) ||X b - y||^2 + penalty`` (_lasso.py:120) and uses ``group_weights =
ones`` by default (_lasso.py:233-235); its docstrings say otherwise.  The code is
what is reproduced here.
"""

from __future__ import annotations

import warnings
from numbers import Real

import numpy as np
from sklearn.utils._param_validation import Interval
from sklearn.utils.validation import check_scalar

from .._utils.validation import _check_group_weights, _check_groups
from ._base import EngineRegressor, ProblemSpec, group_structure

_NONNEG = [Interval(Real, 0.0, None, closed="left")]


def _arr_key(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    return (a.shape, a.dtype.str, a.tobytes())


class Lasso(EngineRegressor):
    r"""Lasso: ``1/(2n) ||X b - y||_2^2 + alpha ||b||_1`` (reference _lasso.py:34-121).

    Args:
        alpha (float): regularisation strength, >= 0.
        fit_intercept, copy_X, warm_start, solver, solver_options: see EngineRegressor.
    """

    _parameter_constraints: dict = {**EngineRegressor._parameter_constraints, "alpha": _NONNEG}

    def __init__(self, alpha=1.0, fit_intercept=False, copy_X=True, warm_start=False, solver=None,
                 solver_options=None):
        super().__init__(fit_intercept=fit_intercept, copy_X=copy_X, warm_start=warm_start, solver=solver,
                         solver_options=solver_options)
        self.alpha = alpha

    def _problem_spec(self, n_features):
        return ProblemSpec(p=n_features, pe=n_features, lam1=float(self.alpha),
                           key=("Lasso", n_features, bool(self.fit_intercept)))


class GroupLasso(Lasso):
    r"""Group Lasso: ``1/(2n)||Xb-y||^2 + alpha sum_g w_g ||b_g||_2`` (_lasso.py:124-275).

    Args:
        groups (list | ndarray | None): group label of every feature; None makes every
            feature its own group (and warns, as the reference does).
        alpha (float): regularisation strength.
        group_weights (ndarray | None): one weight per group in sorted-label order;
            default ones.
        standardize (bool): penalise ``||X_g b_g||`` instead of ``||b_g||``.
    """

    _parameter_constraints: dict = {
        **Lasso._parameter_constraints,
        "groups": "no_validation",
        "group_weights": "no_validation",
        "standardize": ["boolean"],
    }

    def __init__(self, groups=None, alpha=1.0, group_weights=None, standardize=False, fit_intercept=False,
                 copy_X=True, warm_start=False, solver=None, solver_options=None):
        self.groups = groups
        self.standardize = standardize
        self.group_weights = group_weights
        super().__init__(alpha=alpha, fit_intercept=fit_intercept, copy_X=copy_X, warm_start=warm_start,
                         solver=solver, solver_options=solver_options)

    # -- validation (reference _lasso.py:208-222) --
    def _n_groups(self, n_features):
        if self.groups is None:
            return n_features
        cache = self._cached_groups(n_features)
        if cache is not None:
            return cache[4]
        return len(np.unique(self.groups))

    def _validate_hyperparams(self, X, y):
        super()._validate_hyperparams(X, y)
        if self.groups is None:
            warnings.warn(
                "groups has not been supplied such that the problem reduces to a simple Lasso. "
                "You should consider using that instead.",
                UserWarning,
            )
        _check_groups(self.groups, X.shape[1])
        _check_group_weights(self.group_weights, self._n_groups(X.shape[1]))

    # -- problem description --
    def _check_standardize(self, separable=True):
        """standardize=True (group norms ||X_g b_g||, reference _lasso.py:249-252) is solved in
        per-group whitened variables when the penalty is a function of the group norms alone.  The
        l1 term of the sparse-group estimators is not separable in those variables: they go through
        the method of multipliers of sparselm_b200/split.py (ProblemSpec.split)."""
        return bool(self.standardize)

    def _cached_groups(self, n_features):
        """The memoised group structure if it still describes `self.groups`: same object AND same
        content (a list or array mutated in place between fits must not reuse the old structure)."""
        cache = self.__dict__.get("_group_cache")
        if cache is None or cache[0] is not self.groups or cache[1] != n_features:
            return None
        if self.__dict__.get("_groups_frozen"):
            return cache  # inside one grid-search plan: `groups` cannot have been mutated since the last check
        if self.groups is not None and hash(np.asarray(self.groups).tobytes()) != cache[6]:
            return None
        return cache

    def _group_spec(self, n_features):
        """(col_perm, gptr, gw) for the current groups (memoised on the identity + a content
        fingerprint of `groups`: a grid search re-describes the same structure per candidate)."""
        cache = self._cached_groups(n_features)
        if cache is None:
            groups = np.arange(n_features) if self.groups is None else np.asarray(self.groups)
            col_perm, gptr, n_groups = group_structure(groups, n_features)
            fp = None if self.groups is None else hash(np.asarray(self.groups).tobytes())
            cache = (self.groups, n_features, col_perm, gptr, n_groups, _arr_key(self.groups), fp)
            self.__dict__["_group_cache"] = cache
        _, _, col_perm, gptr, n_groups, _, _ = cache
        gw = np.ones(n_groups) if self.group_weights is None else np.asarray(self.group_weights, dtype=float)
        return col_perm, gptr, gw

    def _structure_key(self, name, n_features):
        cache = self._cached_groups(n_features)
        gkey = cache[5] if cache is not None else _arr_key(self.groups)
        return (name, n_features, bool(self.fit_intercept), bool(self.standardize), gkey,
                _arr_key(self.group_weights))

    def _problem_spec(self, n_features):
        std = self._check_standardize()
        col_perm, gptr, gw = self._group_spec(n_features)
        return ProblemSpec(p=n_features, pe=n_features, lam1=0.0, col_perm=col_perm, gptr=gptr, gw=gw,
                           w2=float(self.alpha) * gw, standardize=std,
                           key=self._structure_key("GroupLasso", n_features))


class OverlapGroupLasso(GroupLasso):
    r"""Overlap group Lasso by column duplication (_lasso.py:279-502).

    Args:
        group_list (list[list[int]] | None): for every feature the ids of the groups it
            belongs to.  Features are duplicated once per group, a plain group Lasso is
            solved on the expanded design and duplicated coefficients are summed back.
    """

    _parameter_constraints: dict = {
        **{k: v for k, v in GroupLasso._parameter_constraints.items() if k != "groups"},
        "group_list": "no_validation",
    }

    def __init__(self, group_list=None, alpha=1.0, group_weights=None, standardize=False, fit_intercept=False,
                 copy_X=True, warm_start=False, solver=None, solver_options=None, **kwargs):
        self.group_list = group_list
        super().__init__(groups=None, alpha=alpha, group_weights=group_weights, standardize=standardize,
                         fit_intercept=fit_intercept, copy_X=copy_X, warm_start=warm_start, solver=solver,
                         solver_options=solver_options)

    def _n_groups(self, n_features):
        if self.group_list is None:
            return n_features
        return len(np.unique([gid for grp in self.group_list for gid in grp]))

    def _validate_hyperparams(self, X, y):
        # skips the GroupLasso `groups` validation (reference _lasso.py:371-391)
        EngineRegressor._validate_hyperparams(self, X, y)
        if self.group_list is not None:
            if len(self.group_list) != X.shape[1]:
                raise ValueError("The length of the group list must be the same as the number of features.")
        else:
            warnings.warn(
                "No group list has been supplied such that the problem reduces to a simple Lasso. "
                "You should consider using that instead.",
                UserWarning,
            )
        _check_group_weights(self.group_weights, self._n_groups(X.shape[1]))

    def _expansion(self, n_features):
        """beta_indices, gptr of the duplicated-column problem (reference _lasso.py:440-459)."""
        group_list = [[i] for i in range(n_features)] if self.group_list is None else self.group_list
        memberships = [np.unique(np.asarray(grp).reshape(-1)) for grp in group_list]
        ids = np.unique(np.concatenate(memberships)) if len(memberships) else np.zeros(0, dtype=int)
        feat = np.concatenate([np.full(len(m), j) for j, m in enumerate(memberships)])
        gid = np.searchsorted(ids, np.concatenate(memberships))
        order = np.lexsort((feat, gid))  # by group, then ascending feature (as the reference's scan)
        ext_idx = feat[order].astype(np.int32)
        counts = np.bincount(gid, minlength=len(ids))
        gptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        return ext_idx, gptr, len(ids)

    def _structure_key(self, name, n_features):
        gl = None if self.group_list is None else tuple(tuple(np.asarray(g).reshape(-1).tolist()) for g in self.group_list)
        return (name, n_features, bool(self.fit_intercept), bool(self.standardize), gl, _arr_key(self.group_weights))

    def _problem_spec(self, n_features):
        std = self._check_standardize()
        ext_idx, gptr, n_groups = self._expansion(n_features)
        gw = np.ones(n_groups) if self.group_weights is None else np.asarray(self.group_weights, dtype=float)
        return ProblemSpec(p=n_features, pe=len(ext_idx), lam1=0.0, ext_idx=ext_idx, gptr=gptr, gw=gw,
                           w2=float(self.alpha) * gw, standardize=std,
                           key=self._structure_key("OverlapGroupLasso", n_features))


class SparseGroupLasso(GroupLasso):
    r"""Sparse group Lasso: ``lambda1 ||b||_1 + lambda2 sum_g w_g ||b_g||``,
    ``lambda1 = l1_ratio*alpha``, ``lambda2 = (1-l1_ratio)*alpha`` (_lasso.py:505-639)."""

    _parameter_constraints: dict = {**GroupLasso._parameter_constraints, "l1_ratio": "no_validation"}

    def __init__(self, groups=None, l1_ratio=0.5, alpha=1.0, group_weights=None, standardize=False,
                 fit_intercept=False, copy_X=True, warm_start=False, solver=None, solver_options=None):
        super().__init__(groups=groups, alpha=alpha, group_weights=group_weights, standardize=standardize,
                         fit_intercept=fit_intercept, copy_X=copy_X, warm_start=warm_start, solver=solver,
                         solver_options=solver_options)
        self.l1_ratio = l1_ratio

    def _validate_hyperparams(self, X, y):
        super()._validate_hyperparams(X, y)
        check_scalar(self.l1_ratio, "l1_ratio", float, min_val=0, max_val=1)  # _lasso.py:597
        if self.l1_ratio == 0.0:
            warnings.warn("It is more efficient to use GroupLasso directly than SparseGroupLasso with l1_ratio=0",
                          UserWarning)
        if self.l1_ratio == 1.0:
            warnings.warn("It is more efficient to use Lasso directly than SparseGroupLasso with l1_ratio=1",
                          UserWarning)

    def _lambdas(self):
        return self.l1_ratio * self.alpha, (1 - self.l1_ratio) * self.alpha  # _lasso.py:621-624

    def _problem_spec(self, n_features):
        std = self._check_standardize(separable=False)
        col_perm, gptr, gw = self._group_spec(n_features)
        lam1, lam2 = self._lambdas()
        return ProblemSpec(p=n_features, pe=n_features, lam1=float(lam1), col_perm=col_perm, gptr=gptr, gw=gw,
                           w2=float(lam2) * gw, split=std, key=self._structure_key("SparseGroupLasso", n_features))


class RidgedGroupLasso(GroupLasso):
    r"""Ridged group Lasso: ``alpha sum_g w_g ||b_g|| + 1/2 sum_g delta_g ||b_g||^2``
    (_lasso.py:642-811); ``delta`` has length 1 (broadcast) or n_groups."""

    _parameter_constraints: dict = {**GroupLasso._parameter_constraints,
                                    "delta": ["array-like", Interval(Real, 0.0, None, closed="left")]}

    def __init__(self, groups=None, alpha=1.0, delta=(1.0,), group_weights=None, standardize=False,
                 fit_intercept=False, copy_X=True, warm_start=False, solver=None, solver_options=None):
        super().__init__(groups=groups, alpha=alpha, group_weights=group_weights, standardize=standardize,
                         fit_intercept=fit_intercept, copy_X=copy_X, warm_start=warm_start, solver=solver,
                         solver_options=solver_options)
        self.delta = delta

    def _validate_hyperparams(self, X, y):
        super()._validate_hyperparams(X, y)
        n_groups = self._n_groups(X.shape[1])
        if np.any(np.asarray(self.delta, dtype=float) < 0):
            raise ValueError("delta must be non-negative")
        if len(self.delta) != n_groups and len(self.delta) != 1:  # _lasso.py:750-753
            raise ValueError(f"delta must be an array of length 1 or equal to the number of groups {n_groups}.")

    def _delta_vector(self, n_groups):
        d = np.asarray(self.delta, dtype=float)
        return d * np.ones(n_groups) if len(d) != n_groups else d  # _lasso.py:762-764

    def _ridge_fields(self, std, n_groups):
        """ProblemSpec fields of the ridge: a prox-side weight normally; with standardize the
        norm is ||sqrtm(X_g^T X_g + sqrt(delta_g) I) b_g|| (_lasso.py:779-786), the whitening then
        depends on delta and the ridge is folded into the whitened Gram."""
        dl = self._delta_vector(n_groups)
        if std:
            return dict(d2=None, std_delta=dl), (_arr_key(dl),)
        return dict(d2=dl), ()

    def _problem_spec(self, n_features):
        std = self._check_standardize()
        col_perm, gptr, gw = self._group_spec(n_features)
        ridge, rkey = self._ridge_fields(std, len(gw))
        return ProblemSpec(p=n_features, pe=n_features, lam1=0.0, col_perm=col_perm, gptr=gptr, gw=gw,
                           w2=float(self.alpha) * gw, standardize=std, **ridge,
                           key=self._structure_key("RidgedGroupLasso", n_features) + rkey)

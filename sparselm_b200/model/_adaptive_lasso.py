"""Adaptive (iteratively re-weighted) estimators
(reference: src/sparselm/model/_adaptive_lasso.py).

Every fit runs up to ``max_iter`` weighted solves; after each solve the penalty
weights are recomputed from the solution and the loop stops when the weights move
by less than ``tol`` in l2 norm (reference _adaptive_lasso.py:206-232).  The
quirks of the reference are kept on purpose:

* the default update already contains alpha, ``u(x) = alpha / (|x| + eps)``
  (:181), and the caller multiplies by alpha (or lambda) again (:204, :372-374,
  :721-726);
* the first pass of the group variants uses ``alpha * ones`` -- no
  ``group_weights`` (:347-351, :658-667);
* the returned coefficients are those of the last executed solve.
"""

from __future__ import annotations

import warnings
from numbers import Integral, Real

from sklearn.utils._param_validation import Interval

from ._base import ProblemSpec
from ._lasso import GroupLasso, Lasso, OverlapGroupLasso, RidgedGroupLasso, SparseGroupLasso

_ADAPTIVE_CONSTRAINTS = {
    "tol": [Interval(Real, 0.0, 1.0, closed="both")],
    "max_iter": [Interval(Integral, 0, None, closed="left")],
    "eps": [Interval(Real, 0.0, 1.0, closed="both")],
    "update_function": [callable, None],
}


def _warn_max_iter(est):
    if est.max_iter == 1:  # reference :142-147
        warnings.warn(
            "max_iter is set to 1. It should ideally be set > 1, otherwise consider "
            "using a non-adaptive Regressor",
            UserWarning,
        )


def _adaptive(est, a1, a2):
    return dict(a1=a1, a2=a2, alpha=float(est.alpha), eps=float(est.eps), tol=float(est.tol),
                max_iter=int(est.max_iter), update_function=est.update_function)


def _fn_key(est):
    return (float(est.eps), float(est.tol), int(est.max_iter),
            None if est.update_function is None else id(est.update_function))


class AdaptiveLasso(Lasso):
    r"""Adaptive Lasso: ``|| w o b ||_1`` with ``w <- alpha * alpha / (|b| + eps)``
    (reference :45-232).

    Args:
        alpha (float): regularisation strength.
        max_iter (int): maximum number of re-weighting passes (solves).
        eps (float): offset in the weight update.
        tol (float): stop when the weights change by less than tol (l2 norm).
        update_function (callable | None): ``f(beta, eps) -> weights``; default
            ``alpha / (abs(beta) + eps)``.  A user function is evaluated on the host
            between passes.
    """

    _parameter_constraints: dict = {**Lasso._parameter_constraints, **_ADAPTIVE_CONSTRAINTS}

    def __init__(self, alpha=1.0, max_iter=3, eps=1e-6, tol=1e-10, update_function=None, fit_intercept=False,
                 copy_X=True, warm_start=True, solver=None, solver_options=None, **kwargs):
        Lasso.__init__(self, alpha=alpha, fit_intercept=fit_intercept, copy_X=copy_X, warm_start=warm_start,
                       solver=solver, solver_options=solver_options)
        self.tol = tol
        self.max_iter = max_iter
        self.eps = eps
        self.update_function = update_function

    def _validate_hyperparams(self, X, y):
        super()._validate_hyperparams(X, y)
        _warn_max_iter(self)

    def _problem_spec(self, n_features):
        a = float(self.alpha)
        return ProblemSpec(p=n_features, pe=n_features, lam1=a, adaptive=_adaptive(self, a, None),
                           key=("AdaptiveLasso", n_features, bool(self.fit_intercept), _fn_key(self)))


class AdaptiveGroupLasso(AdaptiveLasso, GroupLasso):
    r"""Adaptive group Lasso: ``sum_g v_g ||b_g||``, ``v <- (alpha w_g) alpha / (||b_g|| + eps)``
    (reference :235-374)."""

    _parameter_constraints: dict = {**GroupLasso._parameter_constraints, **_ADAPTIVE_CONSTRAINTS}

    def __init__(self, groups=None, alpha=1.0, group_weights=None, max_iter=3, eps=1e-6, tol=1e-10,
                 update_function=None, standardize=False, fit_intercept=False, copy_X=True, warm_start=True,
                 solver=None, solver_options=None, **kwargs):
        GroupLasso.__init__(self, groups=groups, alpha=alpha, group_weights=group_weights, standardize=standardize,
                            fit_intercept=fit_intercept, copy_X=copy_X, warm_start=warm_start, solver=solver,
                            solver_options=solver_options)
        self.tol = tol
        self.max_iter = max_iter
        self.eps = eps
        self.update_function = update_function

    def _validate_hyperparams(self, X, y):
        GroupLasso._validate_hyperparams(self, X, y)
        _warn_max_iter(self)

    def _problem_spec(self, n_features):
        std = self._check_standardize()
        col_perm, gptr, gw = self._group_spec(n_features)
        a = float(self.alpha)
        return ProblemSpec(p=n_features, pe=n_features, lam1=0.0, col_perm=col_perm, gptr=gptr, gw=gw,
                           w2=a + 0.0 * gw, adaptive=_adaptive(self, None, a), standardize=std,
                           key=self._structure_key("AdaptiveGroupLasso", n_features) + (_fn_key(self),))


class AdaptiveOverlapGroupLasso(OverlapGroupLasso, AdaptiveGroupLasso):
    r"""Adaptive overlap group Lasso (reference :377-524): AdaptiveGroupLasso on the
    duplicated-column problem, coefficients summed back."""

    _parameter_constraints: dict = {**OverlapGroupLasso._parameter_constraints, **_ADAPTIVE_CONSTRAINTS}

    def __init__(self, group_list=None, alpha=1.0, group_weights=None, max_iter=3, eps=1e-6, tol=1e-10,
                 update_function=None, standardize=False, fit_intercept=False, copy_X=True, warm_start=True,
                 solver=None, solver_options=None):
        OverlapGroupLasso.__init__(self, group_list=group_list, alpha=alpha, group_weights=group_weights,
                                   standardize=standardize, fit_intercept=fit_intercept, copy_X=copy_X,
                                   warm_start=warm_start, solver=solver, solver_options=solver_options)
        self.tol = tol
        self.max_iter = max_iter
        self.eps = eps
        self.update_function = update_function

    def _validate_hyperparams(self, X, y):
        # the reference's MRO skips the max_iter == 1 warning here (_lasso.py:374)
        OverlapGroupLasso._validate_hyperparams(self, X, y)

    def _problem_spec(self, n_features):
        std = self._check_standardize()
        ext_idx, gptr, n_groups = self._expansion(n_features)
        import numpy as np

        gw = np.ones(n_groups) if self.group_weights is None else np.asarray(self.group_weights, dtype=float)
        a = float(self.alpha)
        return ProblemSpec(p=n_features, pe=len(ext_idx), lam1=0.0, ext_idx=ext_idx, gptr=gptr, gw=gw,
                           w2=a + 0.0 * gw, adaptive=_adaptive(self, None, a), standardize=std,
                           key=self._structure_key("AdaptiveOverlapGroupLasso", n_features) + (_fn_key(self),))


class AdaptiveSparseGroupLasso(AdaptiveLasso, SparseGroupLasso):
    r"""Adaptive sparse group Lasso (reference :527-726): ``||w o b||_1 + sum_g v_g ||b_g||``,
    ``w <- lambda1 alpha/(|b|+eps)``, ``v <- (lambda2 w_g) alpha/(||b_g||+eps)``."""

    _parameter_constraints: dict = {**SparseGroupLasso._parameter_constraints, **_ADAPTIVE_CONSTRAINTS}

    def __init__(self, groups=None, l1_ratio=0.5, alpha=1.0, group_weights=None, max_iter=3, eps=1e-6,
                 tol=1e-10, update_function=None, standardize=False, fit_intercept=False, copy_X=True,
                 warm_start=True, solver=None, solver_options=None):
        SparseGroupLasso.__init__(self, groups=groups, l1_ratio=l1_ratio, alpha=alpha, group_weights=group_weights,
                                  standardize=standardize, fit_intercept=fit_intercept, copy_X=copy_X,
                                  warm_start=warm_start, solver=solver, solver_options=solver_options)
        self.tol = tol
        self.max_iter = max_iter
        self.eps = eps
        self.update_function = update_function

    def _validate_hyperparams(self, X, y):
        SparseGroupLasso._validate_hyperparams(self, X, y)
        _warn_max_iter(self)

    def _problem_spec(self, n_features):
        std = self._check_standardize(separable=False)
        col_perm, gptr, gw = self._group_spec(n_features)
        lam1, lam2 = (float(v) for v in self._lambdas())
        return ProblemSpec(p=n_features, pe=n_features, lam1=lam1, col_perm=col_perm, gptr=gptr, gw=gw,
                           w2=lam2 + 0.0 * gw, adaptive=_adaptive(self, lam1, lam2), split=std,
                           key=self._structure_key("AdaptiveSparseGroupLasso", n_features) + (_fn_key(self),))


class AdaptiveRidgedGroupLasso(AdaptiveGroupLasso, RidgedGroupLasso):
    r"""Adaptive ridged group Lasso (reference :729-860): adaptive group penalty plus the
    un-reweighted ridge ``1/2 sum_g delta_g ||b_g||^2``."""

    _parameter_constraints: dict = {**RidgedGroupLasso._parameter_constraints, **_ADAPTIVE_CONSTRAINTS}

    def __init__(self, groups=None, alpha=1.0, delta=(1.0,), group_weights=None, max_iter=3, eps=1e-6, tol=1e-10,
                 update_function=None, standardize=False, fit_intercept=False, copy_X=True, warm_start=True,
                 solver=None, solver_options=None):
        RidgedGroupLasso.__init__(self, groups=groups, alpha=alpha, delta=delta, group_weights=group_weights,
                                  standardize=standardize, fit_intercept=fit_intercept, copy_X=copy_X,
                                  warm_start=warm_start, solver=solver, solver_options=solver_options)
        self.tol = tol
        self.max_iter = max_iter
        self.eps = eps
        self.update_function = update_function

    def _validate_hyperparams(self, X, y):
        RidgedGroupLasso._validate_hyperparams(self, X, y)
        _warn_max_iter(self)

    def _problem_spec(self, n_features):
        std = self._check_standardize()
        col_perm, gptr, gw = self._group_spec(n_features)
        a = float(self.alpha)
        ridge, rkey = self._ridge_fields(std, len(gw))
        return ProblemSpec(p=n_features, pe=n_features, lam1=0.0, col_perm=col_perm, gptr=gptr, gw=gw,
                           w2=a + 0.0 * gw, adaptive=_adaptive(self, None, a), standardize=std, **ridge,
                           key=self._structure_key("AdaptiveRidgedGroupLasso", n_features) + rkey + (_fn_key(self),))

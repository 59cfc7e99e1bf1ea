"""Edge cases of the path: degenerate designs and targets, penalties that zero everything,
ragged folds, one-column grids, more features than rows -- each against the CPU oracle or a
closed form."""

import numpy as np
import numpy.testing as npt
import pytest

pytestmark = pytest.mark.gpu

import oracle.reference as R  # noqa: E402
from sparselm_b200.model import (  # noqa: E402
    AdaptiveGroupLasso,
    AdaptiveLasso,
    GroupLasso,
    Lasso,
    OverlapGroupLasso,
    RidgedGroupLasso,
    SparseGroupLasso,
)
from sparselm_b200.model_selection import GridSearchCV  # noqa: E402
from test_gpu_estimators import _cv_reference  # noqa: E402


def _data(seed, n, p, k=4, noise=0.3):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    w = np.zeros(p)
    w[rng.choice(p, min(k, p), replace=False)] = 2.0 * rng.standard_normal(min(k, p))
    return X, X @ w + noise * rng.standard_normal(n), rng


def _objective(name, X, y, b, icpt, alpha, groups=None, l1_ratio=0.5):
    n = len(y)
    r = y - X @ b - icpt
    if name == "Lasso":
        pen = alpha * np.abs(b).sum()
    else:
        nr = np.array([np.linalg.norm(b[groups == g]) for g in np.unique(groups)])
        pen = alpha * nr.sum() if name == "GroupLasso" else alpha * (l1_ratio * np.abs(b).sum() + (1 - l1_ratio) * nr.sum())
    return r @ r / (2 * n) + pen


@pytest.mark.parametrize("cls", [Lasso, GroupLasso, SparseGroupLasso, RidgedGroupLasso, OverlapGroupLasso,
                                 AdaptiveLasso, AdaptiveGroupLasso])
@pytest.mark.parametrize("fit_intercept", [False, True])
def test_penalty_above_alpha_max_gives_exactly_zero(cls, fit_intercept):
    X, y, rng = _data(3, 60, 24)
    y = y + 5.0
    groups = np.repeat(np.arange(6), 4)
    kw = {}
    if "Overlap" in cls.__name__:
        kw["group_list"] = [[int(g), int((g + 1) % 6)] if j % 2 else [int(g)] for j, g in enumerate(groups)]
    elif "Lasso" != cls.__name__.replace("Adaptive", ""):
        kw["groups"] = groups
    yc = y - y.mean() if fit_intercept else y
    Xc = X - X.mean(0) if fit_intercept else X
    # every penalty of the path is >= alpha * min(l1, group-l2) weighting: 10 * ||X'y||_2 / n zeroes them all
    alpha = 10.0 * np.linalg.norm(Xc.T @ yc) / len(y)
    est = cls(alpha=alpha, fit_intercept=fit_intercept, **kw).fit(X, y)
    assert est.solver_info_["status"] == 0
    assert np.all(est.coef_ == 0.0)
    assert est.intercept_ == (pytest.approx(y.mean(), rel=1e-12) if fit_intercept else 0.0)
    npt.assert_allclose(est.predict(X[:3]), y.mean() if fit_intercept else 0.0, rtol=1e-12)


@pytest.mark.parametrize("fit_intercept", [False, True])
def test_constant_and_zero_targets(fit_intercept):
    X, _, rng = _data(4, 40, 10)
    groups = np.repeat(np.arange(5), 2)
    for cls, kw in ((Lasso, {}), (SparseGroupLasso, {"groups": groups})):
        est = cls(alpha=0.1, fit_intercept=fit_intercept, **kw).fit(X, np.zeros(40))
        assert est.solver_info_["status"] == 0 and np.all(est.coef_ == 0.0) and est.intercept_ == 0.0
        if fit_intercept:  # a constant target is all intercept
            est = cls(alpha=0.1, fit_intercept=True, **kw).fit(X, np.full(40, 7.25))
            assert est.solver_info_["status"] == 0
            assert np.abs(est.coef_).max() <= 1e-12 and est.intercept_ == pytest.approx(7.25, rel=1e-12)


def test_zero_column_and_duplicated_columns():
    """An all-zero feature and an exact copy of a feature: the Gram is singular, the minimiser is
    not unique in the copies -- objective, predictions and the sum over the copies are."""
    X, y, rng = _data(6, 80, 12)
    X[:, 3] = 0.0
    X[:, 7] = X[:, 2]
    groups = np.repeat(np.arange(4), 3)
    for name, cls, kw in (("Lasso", Lasso, {}), ("GroupLasso", GroupLasso, {"groups": groups}),
                          ("SparseGroupLasso", SparseGroupLasso, {"groups": groups, "l1_ratio": 0.5})):
        for fi in (False, True):
            est = cls(alpha=0.05, fit_intercept=fi, solver_options={"tol": 1e-12}, **kw).fit(X, y)
            b_ref, i_ref = R.fit(name, X, y, alpha=0.05, fit_intercept=fi, **kw)
            assert est.solver_info_["status"] == 0
            assert est.coef_[3] == 0.0
            o_est = _objective(name, X, y, est.coef_, est.intercept_, 0.05, groups)
            o_ref = _objective(name, X, y, b_ref, i_ref, 0.05, groups)
            assert abs(o_est - o_ref) <= 1e-8 * abs(o_ref)
            npt.assert_allclose(est.predict(X), X @ b_ref + i_ref, rtol=0, atol=1e-6 * np.abs(y).max())
            if name == "Lasso":
                assert est.coef_[2] + est.coef_[7] == pytest.approx(b_ref[2] + b_ref[7], abs=1e-6 * np.abs(b_ref).max())


def test_ragged_folds_single_column_grid():
    """n = 103 in 5 folds (21, 21, 21, 20, 20 rows) and a grid of ONE alpha."""
    X, y, rng = _data(7, 103, 17)
    groups = rng.integers(0, 4, size=17)
    groups[:4] = np.arange(4)
    for fi in (False, True):
        gs = GridSearchCV(GroupLasso(groups=groups, fit_intercept=fi, solver_options={"tol": 1e-12}),
                          {"alpha": [0.07]}, cv=5).fit(X, y + 2.0 * fi)
        assert gs.batched_
        ref = _cv_reference("GroupLasso", X, y + 2.0 * fi, [0.07], 5, groups=groups, fit_intercept=fi)
        got = np.array([gs.cv_results_[f"split{i}_test_score"][0] for i in range(5)])
        npt.assert_allclose(got, ref[0], rtol=1e-8)
        assert gs.best_index_ == 0


def test_more_features_than_training_rows():
    X, y, rng = _data(8, 30, 50, k=5, noise=0.1)
    alphas = [0.05, 0.2, 0.8]
    gs = GridSearchCV(Lasso(fit_intercept=True, solver_options={"tol": 1e-12}), {"alpha": alphas}, cv=3).fit(X, y)
    ref = _cv_reference("Lasso", X, y, alphas, 3, fit_intercept=True)
    got = np.stack([gs.cv_results_[f"split{i}_test_score"] for i in range(3)], axis=1)
    npt.assert_allclose(got, ref, rtol=1e-7, atol=1e-9)
    b_ref, i_ref = R.fit("Lasso", X, y, alpha=gs.best_params_["alpha"], fit_intercept=True)
    assert np.abs(gs.best_estimator_.coef_ - b_ref).max() <= 1e-6 * np.abs(b_ref).max()


def test_two_rows_one_feature_and_wide_single_group():
    X = np.array([[1.0], [3.0]])
    y = np.array([2.0, 5.0])
    est = Lasso(alpha=0.5).fit(X, y)  # closed form: soft(x'y/n, alpha) / (x'x/n) = (8.5 - 0.5) / 5
    assert est.coef_[0] == pytest.approx(1.6, rel=1e-9)
    X, y, rng = _data(9, 50, 33)
    est = GroupLasso(groups=np.zeros(33, dtype=int), alpha=0.1, solver_options={"tol": 1e-12}).fit(X, y)
    b_ref, _ = R.fit("GroupLasso", X, y, alpha=0.1, groups=np.zeros(33, dtype=int))
    assert np.abs(est.coef_ - b_ref).max() <= 1e-6 * np.abs(b_ref).max()


def test_alpha_zero_is_unpenalised_least_squares():
    """alpha = 0 is valid in the reference (_lasso.py:77-79, interval closed at 0): the penalised
    estimators then solve (ridged) least squares.  Routed to conjugate gradients on the Gram (the
    duality-gap test of the proximal iterations degenerates without a penalty): no warning, exact
    answer, also as one candidate of a batched grid and through the adaptive loop."""
    import warnings

    from sparselm_b200.model import AdaptiveLasso, GroupLasso, Lasso, RidgedGroupLasso
    from sparselm_b200.model_selection import GridSearchCV

    rng = np.random.default_rng(41)
    n, p = 200, 30
    X = rng.standard_normal((n, p))
    y = X @ rng.standard_normal(p) + 0.5 * rng.standard_normal(n) + 2.0
    groups = np.repeat(np.arange(6), 5)
    ols = np.linalg.lstsq(X, y, rcond=None)[0]
    Xc, yc = X - X.mean(0), y - y.mean()
    ols_c = np.linalg.lstsq(Xc, yc, rcond=None)[0]
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        est = Lasso(alpha=0.0).fit(X, y)
        assert est.solver_info_["status"] == 0
        np.testing.assert_allclose(est.coef_, ols, rtol=0, atol=1e-9 * np.abs(ols).max())
        est = GroupLasso(groups=groups, alpha=0.0, fit_intercept=True).fit(X, y)
        np.testing.assert_allclose(est.coef_, ols_c, rtol=0, atol=1e-9 * np.abs(ols_c).max())
        assert est.intercept_ == pytest.approx(y.mean() - X.mean(0) @ ols_c, rel=1e-9)
        # ridge only: (X'X + n diag(delta_g)) b = X'y
        delta = np.array([0.5, 0.0, 1.0, 2.0, 0.1, 0.3])
        est = RidgedGroupLasso(groups=groups, alpha=0.0, delta=delta).fit(X, y)
        ridge = np.linalg.solve(X.T @ X + n * np.diag(np.repeat(delta, 5)), X.T @ y)
        np.testing.assert_allclose(est.coef_, ridge, rtol=0, atol=1e-9 * np.abs(ridge).max())
        ad = AdaptiveLasso(alpha=0.0).fit(X, y)
        assert ad.n_iter_ == 1  # the weights stay zero: the reference's loop stops after one solve
        np.testing.assert_allclose(ad.coef_, ols, rtol=0, atol=1e-9 * np.abs(ols).max())
        gs = GridSearchCV(Lasso(solver_options={"tol": 1e-12}), {"alpha": [0.0, 0.01, 0.1]}, cv=4).fit(X, y)
    assert gs.batched_ and (gs.solver_info_["status"] == 0).all()
    from sklearn.model_selection import KFold

    for f, (tr, te) in enumerate(KFold(4).split(X)):
        b = np.linalg.lstsq(X[tr], y[tr], rcond=None)[0]
        ref = -np.sqrt(np.mean((y[te] - X[te] @ b) ** 2))
        assert gs.cv_results_[f"split{f}_test_score"][0] == pytest.approx(ref, rel=1e-9)

"""GPU tests of the SURVEY 8f "next" rows: OrdinaryLeastSquares on the engine (conjugate
gradients on the tensor-core Gram), StepwiseEstimator chains of engine-backed estimators and
constrain_coefficients re-fits.  Mirrors the reference's tests/test_ols.py,
tests/test_stepwise.py and tests/test_tools.py; the checker is numpy's least squares (the
closed form the reference's own OLS test uses)."""

import warnings
from functools import partial

import numpy as np
import numpy.testing as npt
import pytest
from sklearn.base import clone
from sklearn.preprocessing import add_dummy_feature

pytestmark = pytest.mark.gpu

from sparselm_b200.model import Lasso, OrdinaryLeastSquares  # noqa: E402
from sparselm_b200.model_selection import GridSearchCV  # noqa: E402
from sparselm_b200.stepwise import StepwiseEstimator  # noqa: E402
from sparselm_b200.tools import constrain_coefficients  # noqa: E402


# ---- reference tests/test_ols.py ------------------------------------------------------------
def test_linear_regression():
    reg = OrdinaryLeastSquares().fit([[1], [2]], [1, 2])
    npt.assert_array_almost_equal(reg.coef_, [1])
    npt.assert_array_almost_equal(reg.intercept_, [0])
    npt.assert_array_almost_equal(reg.predict([[1], [2]]), [1, 2])
    reg = OrdinaryLeastSquares().fit([[1]], [0])  # degenerate input
    npt.assert_array_almost_equal(reg.coef_, [0])
    npt.assert_array_almost_equal(reg.intercept_, [0])
    npt.assert_array_almost_equal(reg.predict([[1]]), [0])


@pytest.mark.parametrize("fit_intercept", [True, False])
def test_linear_regression_sample_weights(fit_intercept):
    rng = np.random.default_rng(0)
    n, p = 10, 8
    X, y = rng.normal(size=(n, p)), rng.normal(size=n)
    sw = 1.0 + rng.uniform(size=n)
    reg = OrdinaryLeastSquares(fit_intercept=fit_intercept).fit(X, y, sample_weight=sw)
    assert reg.coef_.shape == (p,)
    W = np.diag(sw)
    Xa = add_dummy_feature(X) if fit_intercept else X
    ref = np.linalg.solve(Xa.T @ W @ Xa, Xa.T @ W @ y)
    if fit_intercept:
        npt.assert_allclose(reg.coef_, ref[1:], rtol=1e-7)
        npt.assert_allclose(reg.intercept_, ref[0], rtol=1e-7)
    else:
        npt.assert_allclose(reg.coef_, ref, rtol=1e-7)
        assert reg.intercept_ == 0.0


def test_fit_intercept_shapes():
    X2 = np.array([[0.38349978, 0.61650022], [0.58853682, 0.41146318]])
    X3 = np.array([[0.27677969, 0.70693172, 0.01628859], [0.08385139, 0.20692515, 0.70922346]])
    y = np.array([1, 1])
    a, b = OrdinaryLeastSquares(fit_intercept=False).fit(X2, y), OrdinaryLeastSquares(fit_intercept=True).fit(X2, y)
    c, d = OrdinaryLeastSquares(fit_intercept=False).fit(X3, y), OrdinaryLeastSquares(fit_intercept=True).fit(X3, y)
    assert a.coef_.shape == b.coef_.shape and c.coef_.shape == d.coef_.shape and a.coef_.ndim == c.coef_.ndim


@pytest.mark.parametrize("n,p", [(300, 40), (5000, 700), (60, 90)])
def test_ols_matches_lstsq(n, p):
    # over-determined: the unique minimiser; p > n: conjugate gradients from zero stay in
    # range(X^T) and converge to the minimum-norm solution, which is what lstsq returns
    rng = np.random.default_rng(n + p)
    X = rng.normal(size=(n, p)) * (1.0 + 3.0 * rng.random(p))
    y = X @ rng.normal(size=p) + 0.1 * rng.normal(size=n)
    for fit_intercept in (False, True):
        reg = OrdinaryLeastSquares(fit_intercept=fit_intercept).fit(X, y)
        Xc, yc = (X - X.mean(0), y - y.mean()) if fit_intercept else (X, y)
        ref = np.linalg.lstsq(Xc, yc, rcond=None)[0]
        npt.assert_allclose(reg.coef_, ref, rtol=1e-6, atol=1e-8 * np.abs(ref).max())
        if fit_intercept:
            assert reg.intercept_ == pytest.approx(y.mean() - X.mean(0) @ ref, rel=1e-6, abs=1e-8)
        assert reg.solver_info_["status"] == 0
        assert reg.score(X, y) > 0.99


@pytest.mark.parametrize("bad", [np.nan, np.inf])
def test_non_finite_input_raises_value_error(bad):
    # the finiteness check of a plain fit runs on the device (no host pass over X)
    rng = np.random.default_rng(2)
    X = rng.normal(size=(50, 6))
    y = rng.normal(size=50)
    Xb = X.copy()
    Xb[17, 3] = bad
    yb = y.copy()
    yb[5] = bad
    for est in (OrdinaryLeastSquares(), Lasso(alpha=0.1), OrdinaryLeastSquares(fit_intercept=True)):
        with pytest.raises(ValueError, match="NaN|infinity"):
            est.fit(Xb, y)
        with pytest.raises(ValueError, match="NaN|infinity"):
            est.fit(X, yb)
        est.fit(X, y)  # the engine is usable afterwards


def test_ols_equals_vanishing_lasso():
    rng = np.random.default_rng(5)
    X = rng.normal(size=(200, 15))
    y = X @ rng.normal(size=15) + 0.05 * rng.normal(size=200)
    ols = OrdinaryLeastSquares().fit(X, y)
    las = Lasso(alpha=1e-10, solver_options={"tol": 1e-12}).fit(X, y)
    npt.assert_allclose(ols.coef_, las.coef_, atol=1e-6)


def test_ols_in_sklearn_grid_search():
    from sklearn.model_selection import GridSearchCV as SkGrid

    rng = np.random.default_rng(6)
    X = rng.normal(size=(60, 5))
    y = X @ rng.normal(size=5) + 3.0
    gs = SkGrid(OrdinaryLeastSquares(), {"fit_intercept": [False, True]}, cv=3).fit(X, y)
    assert gs.best_params_ == {"fit_intercept": True}
    ours = GridSearchCV(OrdinaryLeastSquares(), {"fit_intercept": [False, True]}, cv=3).fit(X, y)
    assert ours.best_params_ == {"fit_intercept": True} and ours.batched_ is False


# ---- reference tests/test_tools.py -----------------------------------------------------------
@pytest.mark.parametrize("test_number", range(3))
def test_constrain_coefficients_around_engine_fit(test_number):
    rng = np.random.default_rng(test_number)
    n, p = 10, 8
    X, y = rng.normal(size=(n, p)), rng.normal(size=n)
    reg = OrdinaryLeastSquares(fit_intercept=True)
    coefs = reg.fit(X, y).coef_

    def fit(X, y, reg):
        return reg.fit(X, y).coef_

    inds = rng.choice(p, size=3, replace=False)
    low = rng.random(3) - 0.5
    high = rng.random(3) + low
    for kw in (dict(high=2, low=0), dict(high=high, low=low), dict(high=high), dict(low=low)):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            cstr = constrain_coefficients(inds, **kw)(partial(fit, reg=reg))(X, y)
        assert cstr.shape == coefs.shape
        if not any(issubclass(x.category, RuntimeWarning) for x in w):
            lo = np.broadcast_to(kw.get("low", -np.inf), (3,))
            hi = np.broadcast_to(kw.get("high", np.inf), (3,))
            assert np.all(cstr[inds] >= lo - 1e-12) and np.all(cstr[inds] <= hi + 1e-12)

            # the same bounds enforced by hand around numpy's closed form
            def lstsq_fit(X, y):
                return np.linalg.lstsq(X - X.mean(0), y - y.mean(), rcond=None)[0]

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = constrain_coefficients(inds, **kw)(lstsq_fit)(X, y)
            npt.assert_allclose(cstr, ref, rtol=1e-6, atol=1e-8)


# ---- reference tests/test_stepwise.py:82-160 (convex steps) ----------------------------------
def test_toy_composite():
    rng = np.random.default_rng(3)
    lasso1 = Lasso(fit_intercept=True, alpha=1e-6)
    lasso2 = Lasso(fit_intercept=False, alpha=1e-6)
    lasso3 = Lasso(fit_intercept=False, alpha=1e-6)
    grid = GridSearchCV(clone(lasso2), {"alpha": [1e-8, 1e-7, 1e-6]})
    scopes = [[0, 1, 8], [2, 3], [4, 5, 6, 7]]
    est1 = StepwiseEstimator([("lasso1", lasso1), ("lasso2", lasso2), ("lasso3", lasso3)], scopes)
    est2 = StepwiseEstimator([("lasso1", clone(lasso1)), ("lasso2", grid), ("lasso3", clone(lasso3))], scopes)

    w_test = rng.normal(scale=2, size=9) * 0.2
    w_test[0], w_test[-1] = 10, 0.5
    X = rng.random(size=(20, 9))
    X[:, 0] = 1
    X[:, -1] = -8 * rng.random(size=20)
    y = X @ w_test + rng.normal(scale=0.01, size=20)

    for est in (est1, est2):
        est.fit(X, y)
        assert est.intercept_ == est.steps[0][1].intercept_
        assert not np.any(np.isnan(est.coef_))
        assert not np.isclose(est.intercept_, 0)
        for (_, sub), scope in zip(est.steps, est.estimator_feature_indices):
            sub_coef = sub.best_estimator_.coef_ if hasattr(sub, "estimator") else sub.coef_
            npt.assert_array_almost_equal(sub_coef, est.coef_[scope])
        coef_1, intercept_1 = est.coef_.copy(), est.intercept_
        est.steps[0][1].fit_intercept = False
        est.fit(X, y)
        coef_2 = est.coef_.copy()
        assert np.isclose(est.intercept_, 0)
        assert abs(coef_1[0] + intercept_1 - 10) / 10 <= 0.1
        assert abs(coef_2[0] - 10) / 10 <= 0.1
        total = np.zeros(len(y))
        for (_, sub), scope in zip(est.steps, est.estimator_feature_indices):
            total += sub.predict(X[:, scope])
        npt.assert_array_almost_equal(est.predict(X), total)
        npt.assert_array_almost_equal(X @ est.coef_ + est.intercept_, total)


def test_stepwise_steps_equal_manual_residual_chain():
    rng = np.random.default_rng(8)
    X = rng.normal(size=(80, 12))
    y = X @ rng.normal(size=12) + 2.0 + 0.1 * rng.normal(size=80)
    scopes = [[0, 1, 2, 3, 4], [5, 6, 7], [8, 9, 10, 11]]
    est = StepwiseEstimator([("ols", OrdinaryLeastSquares(fit_intercept=True)), ("l1", Lasso(alpha=0.05)),
                             ("l2", Lasso(alpha=0.01))], scopes).fit(X, y)
    a = OrdinaryLeastSquares(fit_intercept=True).fit(X[:, scopes[0]], y)
    r1 = y - a.predict(X[:, scopes[0]])
    b = Lasso(alpha=0.05).fit(X[:, scopes[1]], r1)
    r2 = r1 - b.predict(X[:, scopes[1]])
    c = Lasso(alpha=0.01).fit(X[:, scopes[2]], r2)
    npt.assert_allclose(est.coef_, np.concatenate([a.coef_, b.coef_, c.coef_]), atol=1e-10)
    assert est.intercept_ == a.intercept_
    assert est.score(X, y) > 0.3

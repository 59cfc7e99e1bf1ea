"""CPU tests of the host-side widening rows: dataset generator, coefficient
constraints, r2 conversion against fixtures generated from the reference's own python
(tests/golden/make_golden_host.py), and the StepwiseEstimator parameter / error contract
(reference tests/test_stepwise.py, tests/test_tools.py, tests/test_dataset.py)."""

import json
import os
import warnings

import numpy as np
import numpy.testing as npt
import pytest
from sklearn.base import clone
from sklearn.utils._param_validation import InvalidParameterError

from sparselm_b200.dataset import make_group_regression
from sparselm_b200.model import Lasso, OrdinaryLeastSquares
from sparselm_b200.model_selection import GridSearchCV
from sparselm_b200.stepwise import StepwiseEstimator
from sparselm_b200.tools import constrain_coefficients, r2_score_to_cv_error

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_host.json")))


def _lstsq(X, y):
    return np.linalg.lstsq(X, y, rcond=None)[0]


@pytest.mark.parametrize("case", GOLD["dataset"], ids=lambda c: f"seed{c['seed']}")
def test_make_group_regression_matches_reference_fixture(case):
    X, y, groups, coefs = make_group_regression(coef=True, random_state=case["seed"], **case["kwargs"])
    npt.assert_array_equal(groups, np.array(case["groups"]))
    npt.assert_allclose(X, np.array(case["X"]), rtol=0, atol=0)
    npt.assert_allclose(y, np.array(case["y"]), rtol=0, atol=0)
    npt.assert_allclose(coefs, np.array(case["coefs"]), rtol=0, atol=0)


# reference tests/test_dataset.py:8-58
@pytest.mark.parametrize("n_informative_groups", [5, 20])
@pytest.mark.parametrize("n_features_per_group", [5, 4 * list(range(2, 7))])
@pytest.mark.parametrize("frac_informative_in_group", [1.0, 0.5])
@pytest.mark.parametrize("shuffle", [True, False])
def test_make_group_regression_structure(n_informative_groups, n_features_per_group, frac_informative_in_group,
                                         shuffle):
    X, y, groups, coefs = make_group_regression(n_informative_groups=n_informative_groups,
                                                n_features_per_group=n_features_per_group,
                                                frac_informative_in_group=frac_informative_in_group,
                                                shuffle=shuffle, coef=True)
    sizes = n_features_per_group if isinstance(n_features_per_group, list) else [n_features_per_group] * 20
    assert X.shape == (100, sum(sizes)) and y.shape == (100,) and groups.shape == (sum(sizes),)
    assert len(np.unique(groups)) == 20
    assert (coefs > 0).sum() == sum(round(frac_informative_in_group * sizes[i]) for i in range(n_informative_groups))
    npt.assert_array_almost_equal(X @ coefs, y)
    if shuffle:
        assert (np.diff(groups) == 0).sum() < 19
    assert len(make_group_regression(n_informative_groups=n_informative_groups)) == 3
    with pytest.warns(UserWarning):
        make_group_regression(frac_informative_in_group=1 / 100)
    with pytest.raises(ValueError):
        make_group_regression(n_groups=3, n_features_per_group=[2, 2])


@pytest.mark.parametrize("i", range(len(GOLD["constrain"])))
def test_constrain_coefficients_matches_reference_fixture(i):
    case = GOLD["constrain"][i]
    X, y = np.array(case["X"]), np.array(case["y"])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        coefs = constrain_coefficients(case["indices"], **case["kwargs"])(_lstsq)(X, y)
    npt.assert_allclose(coefs, np.array(case["coefs"]), rtol=1e-12, atol=1e-13)
    assert (len(w) > 0) == case["warned"]
    if case["warned"]:
        assert issubclass(w[0].category, RuntimeWarning)

    @constrain_coefficients(case["indices"], **case["kwargs"])
    def fit(X, y):
        return _lstsq(X, y)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        npt.assert_allclose(fit(X, y), coefs, rtol=0, atol=0)
    assert fit.__name__ == "fit"  # functools.wraps, as in the reference


@pytest.mark.parametrize("case", GOLD["r2_to_cv"])
def test_r2_score_to_cv_error_matches_reference_fixture(case):
    y, yp, w = np.array(case["y"]), np.array(case["y_pred"]), np.array(case["weights"])
    assert r2_score_to_cv_error(case["score"], y, yp, w) == pytest.approx(case["weighted"], rel=1e-14)
    assert r2_score_to_cv_error(case["score"], y, yp) == pytest.approx(case["unweighted"], rel=1e-14)
    with pytest.raises(ValueError):
        r2_score_to_cv_error(0.5, y, yp, w[:-1])
    with pytest.raises(ValueError):
        r2_score_to_cv_error(0.5, y, yp, -w)
    with pytest.raises(ValueError):
        r2_score_to_cv_error(0.5, y, yp, 0 * w)


# ---- StepwiseEstimator: parameter plumbing and error contract (no solve) ------------------
def test_stepwise_params_clone_and_searcher_steps():
    # reference tests/test_stepwise.py:14-79 with convex steps only (L2L0 is out of scope)
    lasso1, lasso2, lasso3 = Lasso(fit_intercept=True, alpha=1.0), Lasso(alpha=2.0), Lasso(alpha=0.1)
    steps = [("lasso1", lasso1), ("lasso2", lasso2), ("lasso3", lasso3)]
    scopes = [[0, 1, 8], [2, 3], [4, 5, 6, 7]]
    est = StepwiseEstimator(steps, scopes)
    assert est.steps[0][1].fit_intercept and not est.steps[1][1].fit_intercept
    params = est.get_params(deep=True)
    assert params["lasso1"].get_params()["alpha"] == 1.0
    assert params["lasso1__alpha"] == 1.0 and params["lasso2__alpha"] == 2.0 and params["lasso3__alpha"] == 0.1
    est.set_params(lasso2__alpha=0.5, lasso3__alpha=0.2)
    params = est.get_params(deep=True)
    assert params["lasso1__alpha"] == 1.0 and params["lasso2__alpha"] == 0.5 and params["lasso3__alpha"] == 0.2
    cloned = clone(est)
    params = cloned.get_params(deep=True)
    assert params["lasso2__alpha"] == 0.5 and params["lasso3__alpha"] == 0.2
    assert cloned.steps[1][1] is not est.steps[1][1]
    grid = GridSearchCV(lasso2, {"alpha": [0.01, 0.1, 1.0]})
    est = StepwiseEstimator([("lasso1", lasso1), ("lasso2", grid), ("lasso3", lasso3)], scopes)
    params = est.get_params(deep=True)
    assert "lasso2__alpha" not in params and params["lasso2__estimator__alpha"] == 0.5


def test_stepwise_structure_errors_come_before_any_solve():
    # reference tests/test_stepwise.py:95-118
    l1, l2, l3 = Lasso(fit_intercept=True, alpha=1e-6), Lasso(alpha=1e-6), Lasso(alpha=1e-6)
    steps = [("a", l1), ("b", l2), ("c", l3)]
    X, y = np.random.default_rng(0).random((20, 9)), np.random.default_rng(1).random(20)
    with pytest.raises(InvalidParameterError):  # scopes with a hole
        StepwiseEstimator(steps, [[0, 1], [3, 4], [5, 6, 7, 8]]).fit(X, y)
    with pytest.raises(InvalidParameterError):  # overlapping scopes
        StepwiseEstimator(steps, [[0, 1, 2], [2, 3], [4, 5, 6, 7, 8]]).fit(X, y)
    with pytest.raises(InvalidParameterError):  # a later step fitting an intercept
        StepwiseEstimator([("a", l1), ("b", Lasso(fit_intercept=True)), ("c", l3)],
                          [[0, 1, 8], [2, 3], [4, 5, 6, 7]]).fit(X, y)
    with pytest.raises(InvalidParameterError):  # also behind a searcher
        StepwiseEstimator([("a", l1), ("b", GridSearchCV(Lasso(fit_intercept=True), {"alpha": [1.0]})), ("c", l3)],
                          [[0, 1, 8], [2, 3], [4, 5, 6, 7]]).fit(X, y)
    inner = StepwiseEstimator([("a", Lasso()), ("b", Lasso())], [[0], [1]])
    with pytest.raises(InvalidParameterError):  # nesting
        StepwiseEstimator([("a", l1), ("b", inner)], [[0, 1, 2, 3, 4, 5, 6], [7, 8]]).fit(X, y)
    with pytest.raises(ValueError):  # feature count fixed by the scopes
        StepwiseEstimator(steps, [[0, 1, 8], [2, 3], [4, 5, 6, 7]]).fit(np.ones((20, 12)), y)


def test_ols_api_surface():
    ols = OrdinaryLeastSquares()
    assert ols.get_params() == dict(fit_intercept=False, copy_X=True, warm_start=False, solver=None,
                                    solver_options=None)
    assert clone(ols.set_params(fit_intercept=True)).fit_intercept is True
    assert OrdinaryLeastSquares._batchable is False
    assert GridSearchCV(ols, {"fit_intercept": [True, False]})._batch_plan(np.ones((6, 2)), np.ones(6), {}) is None

"""The cooperative few-column iterations (csrc/coop_kernels.cuh): a plain fit / refit on a
design too large for the fused shared-memory kernel (p > 160) with at most four grid
columns runs every iteration between two convergence checks inside ONE cooperative launch.
Checked against the CPU oracle and against the regular per-iteration kernels (the same
arithmetic in another parallel decomposition)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle.reference as R  # noqa: E402
from sparselm_b200.engine import get_engine  # noqa: E402
from sparselm_b200.model import (  # noqa: E402
    AdaptiveGroupLasso,
    AdaptiveLasso,
    AdaptiveOverlapGroupLasso,
    AdaptiveRidgedGroupLasso,
    AdaptiveSparseGroupLasso,
    GroupLasso,
    Lasso,
    OverlapGroupLasso,
    RidgedGroupLasso,
    SparseGroupLasso,
)
from sparselm_b200.model._base import solve_specs  # noqa: E402

ALL = [Lasso, GroupLasso, OverlapGroupLasso, SparseGroupLasso, RidgedGroupLasso, AdaptiveLasso, AdaptiveGroupLasso,
       AdaptiveSparseGroupLasso, AdaptiveOverlapGroupLasso, AdaptiveRidgedGroupLasso]


def _problem(n, p, n_groups, seed, intercept=False):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p)) + (0.5 if intercept else 0.0)
    w = np.zeros(p)
    w[rng.choice(p, max(4, p // 12), replace=False)] = 2 * rng.standard_normal(max(4, p // 12))
    y = X @ w + 0.3 * rng.standard_normal(n) + (1.5 if intercept else 0.0)
    groups = rng.integers(0, n_groups, size=p)
    groups[:n_groups] = np.arange(n_groups)
    return X, y, groups, rng


def _kwargs(cls, groups, rng, p):
    name = cls.__name__
    if name in ("Lasso", "AdaptiveLasso"):
        return {}
    kw = {}
    if "Overlap" in name:
        gids = np.unique(groups)
        kw["group_list"] = [list(rng.choice(gids, replace=False, size=rng.integers(1, 3))) for _ in range(p)]
        ng = len(gids)
    else:
        kw["groups"] = groups
        ng = len(np.unique(groups))
    kw["group_weights"] = 0.5 + rng.random(ng)
    if "Ridged" in name:
        kw["delta"] = (0.6,)
    if "Sparse" in name:
        kw["l1_ratio"] = 0.45
    return kw


@pytest.fixture()
def coop_switch():
    eng = get_engine()
    yield eng
    eng.set_option("coop", 1)


@pytest.mark.parametrize("cls", ALL)
@pytest.mark.parametrize("fit_intercept", [False, True])
def test_cooperative_fit_matches_oracle_and_regular_path(cls, fit_intercept, coop_switch):
    eng = coop_switch
    n, p = 260, 300
    X, y, groups, rng = _problem(n, p, 23, 7, fit_intercept)
    kw = _kwargs(cls, groups, rng, p)
    alpha = 0.12
    make = lambda: cls(alpha=alpha, fit_intercept=fit_intercept, solver_options={"tol": 1e-12}, **kw)  # noqa: E731
    eng.set_option("coop", 1)
    l0 = eng.launch_count()
    a = make().fit(X, y)
    launches_coop = eng.launch_count() - l0
    eng.set_option("coop", 0)
    l0 = eng.launch_count()
    b = make().fit(X, y)
    launches_reg = eng.launch_count() - l0
    assert launches_coop < launches_reg  # the cooperative path really ran (fewer, longer launches)
    scale = np.abs(b.coef_).max()
    assert np.abs(a.coef_ - b.coef_).max() <= 1e-7 * scale
    assert a.solver_info_["status"] == 0
    b_ref, i_ref = R.fit(cls.__name__, X, y, alpha=alpha, fit_intercept=fit_intercept, **kw)
    assert np.abs(a.coef_ - b_ref).max() <= 1e-6 * np.abs(b_ref).max()
    assert np.array_equal(np.abs(a.coef_) > 1e-6, np.abs(b_ref) > 1e-6)
    assert abs(a.intercept_ - i_ref) <= 1e-6 * max(1.0, abs(i_ref))


@pytest.mark.parametrize("n,p,n_groups", [(500, 1500, 150), (3000, 2100, 1), (150, 700, 40)])
def test_cooperative_sizes_and_group_shapes(n, p, n_groups, coop_switch):
    """Large / single / ragged groups, p > n; certificate from X and y."""
    eng = coop_switch
    X, y, groups, rng = _problem(n, p, n_groups, n + p)
    amax = np.abs(X.T @ y).max() / n
    for cls, kw in ((Lasso, {}), (SparseGroupLasso, {"groups": groups, "l1_ratio": 0.5})):
        if n_groups == 1 and cls is SparseGroupLasso:
            kw = {"groups": np.zeros(p, dtype=int), "l1_ratio": 0.9}  # one group of 2100 rows: not eligible
        est = cls(alpha=0.1 * amax, solver_options={"tol": 1e-11}, **kw)
        eng.set_option("coop", 1)
        a = est.fit(X, y).coef_.copy()
        info = dict(est.solver_info_)
        eng.set_option("coop", 0)
        b = est.fit(X, y).coef_.copy()
        assert info["status"] == 0
        assert np.abs(a - b).max() <= 1e-7 * max(np.abs(b).max(), 1e-300)
        assert np.count_nonzero(a) > 0


def test_cooperative_multi_column_and_frozen_columns(coop_switch):
    """Up to four columns share one launch (blockIdx.y); a column that converged at a check
    only keeps the grid barriers."""
    eng = coop_switch
    n, p = 400, 640
    X, y, groups, rng = _problem(n, p, 64, 3)
    amax = np.abs(X.T @ y).max() / n
    fd = eng.prepare(X, y, None, False, None, col_perm=SparseGroupLasso(groups=groups)._problem_spec(p).col_perm)
    specs = [SparseGroupLasso(groups=groups, alpha=a * amax)._problem_spec(p) for a in (0.9, 0.3, 0.05, 0.01)]
    outs = []
    for on in (1, 0):
        eng.set_option("coop", on)
        out = solve_specs(eng, fd, specs, use_full=True, tol=1e-11)
        outs.append((out["coef"][0].cpu().numpy()[:, :4].copy(), out["n_iter"][0, :4].copy(),
                     out["status"][0, :4].copy()))
    (ca, ia, sa), (cb, ib, sb) = outs
    assert np.all(sa == 0) and np.all(sb == 0)
    assert np.abs(ca - cb).max() <= 1e-7 * np.abs(cb).max()
    assert ia[0] < ia[3]  # the strongly penalised column converged (and froze) first


# ---- cluster mode: one thread-block cluster per remaining column, switched to mid-solve ----
@pytest.mark.parametrize("cls", [Lasso, SparseGroupLasso, AdaptiveOverlapGroupLasso, RidgedGroupLasso])
def test_cluster_mode_grid_search_matches_regular_path(cls, coop_switch):
    """A small grid (3 folds x 7 alphas) drops under the cluster threshold after the first
    columns converge; scores, coefficients and iteration counts agree with the regular kernels."""
    from sparselm_b200.model_selection import GridSearchCV

    eng = coop_switch
    n, p = 300, 420
    X, y, groups, rng = _problem(n, p, 30, 21)
    kw = _kwargs(cls, groups, rng, p)
    amax = np.abs(X.T @ y).max() / n
    grid = {"alpha": list(amax * np.logspace(-0.3, -2.2, 7))}
    res = []
    for on in (1, 0):
        eng.set_option("coop", on)
        l0 = eng.launch_count()
        gs = GridSearchCV(cls(solver_options={"tol": 1e-11}, **kw), grid, cv=3).fit(X, y)
        res.append((gs, eng.launch_count() - l0))
    (a, la), (b, lb) = res
    assert la < lb  # cluster launches replaced per-iteration launch sets
    assert a.batched_ and b.batched_
    np.testing.assert_allclose(a.cv_results_["mean_test_score"], b.cv_results_["mean_test_score"], rtol=1e-8)
    assert a.best_params_ == b.best_params_
    assert np.abs(a.best_estimator_.coef_ - b.best_estimator_.coef_).max() <= 1e-7 * np.abs(b.best_estimator_.coef_).max()
    assert np.all(a.solver_info_["status"] == 0)


def test_cluster_mode_from_the_start_with_many_columns(coop_switch):
    """More than four columns on one Gram, all under the threshold from the first check."""
    eng = coop_switch
    n, p = 500, 900
    X, y, groups, rng = _problem(n, p, 90, 5)
    amax = np.abs(X.T @ y).max() / n
    est = SparseGroupLasso(groups=groups)
    fd = eng.prepare(X, y, None, False, None, col_perm=est._problem_spec(p).col_perm)
    alphas = amax * np.logspace(-0.2, -2.5, 24)
    specs = [SparseGroupLasso(groups=groups, alpha=a)._problem_spec(p) for a in alphas]
    outs = []
    for on in (1, 0):
        eng.set_option("coop", on)
        out = solve_specs(eng, fd, specs, use_full=True, tol=1e-11)
        outs.append((out["coef"][0].cpu().numpy()[:, :24].copy(), out["status"][0, :24].copy(),
                     out["gap"][0, :24].copy()))
    (ca, sa, ga), (cb, sb, gb) = outs
    assert np.all(sa == 0) and np.all(sb == 0)
    assert np.abs(ca - cb).max() <= 1e-7 * np.abs(cb).max()
    # independent certificate from X and y (oracle) for a few columns
    from sparselm_b200.model._base import _to_original_order

    labels, G = R.group_labels(groups, p)
    for j in (0, 12, 23):
        beta = _to_original_order(ca[:, j], specs[j])
        pen = R.Penalty(labels=labels, w1=np.full(p, 0.5 * alphas[j]), w2=np.full(G, 0.5 * alphas[j]),
                        delta=np.zeros(G))
        cert = R.certificate(X, y, beta, pen)
        assert cert["gap"] <= 1e-9 * max(abs(cert["primal"]), 1e-300), (j, cert)


@pytest.mark.parametrize("cls", [GroupLasso, AdaptiveRidgedGroupLasso])
def test_cooperative_fit_standardized_and_weighted(cls, coop_switch):
    """standardize=True (whitened Gram) and sample weights go through the same cooperative kernel."""
    eng = coop_switch
    n, p = 400, 288
    X, y, _, rng = _problem(n, p, 24, 13, True)
    groups = np.repeat(np.arange(24), 12)
    sw = 2.0 + rng.random(n)  # mean far from 1: the standardized norms see weights normalised to sum n
    kw = {"groups": groups, "standardize": True}
    if "Ridged" in cls.__name__:
        kw["delta"] = (0.4,)
    fits = []
    for on in (1, 0):
        eng.set_option("coop", on)
        fits.append(cls(alpha=0.05, fit_intercept=True, solver_options={"tol": 1e-12}, **kw).fit(X, y, sample_weight=sw))
    a, b = fits
    assert a.solver_info_["status"] == 0
    assert np.abs(a.coef_ - b.coef_).max() <= 1e-7 * np.abs(b.coef_).max()
    assert abs(a.intercept_ - b.intercept_) <= 1e-8 * max(1.0, abs(b.intercept_))
    b_ref, i_ref = R.fit(cls.__name__, X, y, alpha=0.05, fit_intercept=True, sample_weight=sw, **kw)
    assert np.abs(a.coef_ - b_ref).max() <= 1e-6 * np.abs(b_ref).max()
    assert abs(a.intercept_ - i_ref) <= 1e-6 * max(1.0, abs(i_ref))

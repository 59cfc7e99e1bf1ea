"""Parity at BASELINE.json's full sizes, through size-independent properties: the oracle's
solver would need minutes per fit there, but optimality of a given coefficient vector is
cheap to certify from X and y directly (KKT residual in pure numpy, duality gap in C), and CV
scores can be recomputed from the coefficients."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import bench  # noqa: E402
import oracle.reference as R  # noqa: E402


def _cv_step(wl):
    from types import SimpleNamespace

    import torch
    from sklearn.base import clone
    from sklearn.model_selection import KFold

    from sparselm_b200.engine import get_engine
    from sparselm_b200.model_selection import batched_cv

    X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
    engine = get_engine(0)
    p = X.shape[1]
    work = clone(est)
    ests, specs = [], []
    for a in alphas:
        work.set_params(alpha=a)
        specs.append(work._problem_spec(p))
        ests.append(SimpleNamespace(fit_intercept=False))
    folds = [te for _, te in KFold(F).split(X)]
    Xd = torch.from_numpy(X).to(engine.device)
    res = batched_cv(engine, Xd, y, folds, ests, specs, dict(est._engine_options()), "neg_root_mean_squared_error")
    return res, specs, folds


def _penalty(name, spec_kw, alpha, p):
    if name == "Lasso":
        return R.Penalty(np.arange(p), np.full(p, alpha), np.zeros(p), np.zeros(p))
    labels, G = R.group_labels(spec_kw["groups"], p)
    l1r = spec_kw["l1_ratio"]
    return R.Penalty(labels, np.full(p, l1r * alpha), np.full(G, (1 - l1r) * alpha), np.zeros(G))


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_full_size_cv_grid_is_optimal_and_scores_reproduce(name):
    """C2 (Lasso, n=10k, p=2k) and C3 (SparseGroupLasso, n=20k, p=4k, 200 groups): 100 alphas x 5
    folds.  Every problem converges; for a spread of (alpha) cells the fold-0 coefficients pass
    the KKT conditions of the reference objective evaluated from X and y, their duality gap
    (oracle certificate, C) is below 1e-8 relative, and the CV score of the cell equals the
    RMSE recomputed from the coefficients."""
    wl = bench.workload(name)
    X, y, alphas, F = wl["X"], wl["y"], wl["alphas"], wl["F"]
    n, p = X.shape
    res, specs, folds = _cv_step(wl)
    assert res["n_unconverged"] == 0
    assert (res["info"]["status"] == 0).all()
    scores = res["test_scores"]
    assert scores.shape == (len(alphas), F) and np.isfinite(scores).all()
    te = folds[0]
    tr = np.setdiff1d(np.arange(n), te)
    Xt, yt = X[tr], y[tr]
    kw = {k: v for k, v in wl["oracle"].items() if k != "name"}
    supports = []
    for ci in (0, 17, 42, 71, 99):
        b_solver = res["warm"][ci].cpu().numpy()  # fold-0 solution, solver feature order
        b = np.empty(p)
        b[specs[ci].col_perm if specs[ci].col_perm is not None else np.arange(p)] = b_solver
        pen = _penalty(wl["oracle"]["name"], kw, alphas[ci], p)
        cert = R.certificate(Xt, yt, b, pen)
        assert cert["gap"] <= 1e-8 * abs(cert["primal"]), (ci, cert)  # north_star: objective within 1e-8
        # KKT residual (gradient units): an objective gap g bounds it by sqrt(2 L g), L = lambda_max(G)/n
        kkt = R.kkt_residual(Xt, yt, b, pen)
        assert kkt <= np.sqrt(2 * 2.5 * 1e-8 * abs(cert["primal"])), (ci, kkt, cert)
        rmse = np.sqrt(np.mean((y[te] - X[te] @ b) ** 2))
        assert abs(-rmse - scores[ci, 0]) <= 1e-9 * rmse
        supports.append(int(np.sum(np.abs(b) > 1e-6 * np.abs(b).max())) if np.abs(b).max() > 0 else 0)
    # alphas descend from alpha_max: the support grows along the grid
    assert supports[0] <= supports[1] <= supports[2] <= supports[3] <= supports[4] and supports[4] > supports[0]


def test_full_size_adaptive_overlap_batched_cell_equals_single_fit():
    """C4 (AdaptiveOverlapGroupLasso, n=5k, p=1.5k, 30% overlap, 3 passes): one (alpha, fold)
    cell of the batched search against a plain fit of the same estimator on that fold."""
    from sklearn.base import clone
    from sklearn.model_selection import KFold

    from sparselm_b200.model_selection import GridSearchCV

    wl = bench.workload("c4")
    X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
    gs = GridSearchCV(clone(est), {"alpha": list(alphas[[2, 11]])}, cv=F, refit=False).fit(X, y)
    assert gs.batched_ and (gs.solver_info_["status"] == 0).all()
    assert (gs.solver_info_["n_pass"] == 3).all()
    tr, te = list(KFold(F).split(X))[3]
    single = clone(est).set_params(alpha=alphas[11]).fit(X[tr], y[tr])
    assert single.n_iter_ == 3
    rmse = np.sqrt(np.mean((y[te] - single.predict(X[te])) ** 2))
    assert abs(-rmse - gs.cv_results_["split3_test_score"][1]) <= 1e-7 * rmse

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


def test_full_size_adaptive_overlap_final_pass_is_certified_by_the_oracle():
    """C4 at full size against the ORACLE's certificate (not the engine against itself): the third
    pass of three (alpha) cells on one training fold solves a group Lasso on the duplicated-column
    design X[:, beta_indices] (_lasso.py:440-461) with the weights the reference's update
    (_adaptive_lasso.py:364-374) produced from the second pass.  Those weights come out of a
    two-pass run of the same chain; the duality gap of the three-pass coefficients in the extended
    variables is then computed by the oracle (C) from X and y."""
    from sklearn.base import clone
    from sklearn.model_selection import KFold

    from sparselm_b200.engine import get_engine
    from sparselm_b200.model._base import solve_specs

    wl = bench.workload("c4")
    X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
    tr, te = list(KFold(F).split(X))[1]
    Xt, yt = X[tr], y[tr]
    p = X.shape[1]
    engine = get_engine(0)
    cells = [1, 8, 15]

    def run(passes):
        work = clone(est).set_params(max_iter=passes)
        specs = []
        for ci in cells:
            work.set_params(alpha=alphas[ci])
            specs.append(work._problem_spec(p))
        fd = engine.prepare(Xt, yt, None, False, None, col_perm=specs[0].col_perm)
        out = solve_specs(engine, fd, specs, use_full=True, **est._engine_options())
        return specs, out

    specs, out2 = run(2)
    _, out3 = run(3)
    assert (out3["status"][0, :len(cells)] == 0).all() and (out3["n_pass"][0, :len(cells)] == 3).all()
    s0 = specs[0]
    idx = np.asarray(s0.ext_idx)
    # the expansion is the reference's: same indices as the oracle's restatement of _lasso.py:440-461
    ref_idx, ref_labels, G = R.expand_overlap(est.group_list, p)
    assert np.array_equal(idx, ref_idx)
    labels = np.repeat(np.arange(G), np.diff(np.asarray(s0.gptr)))
    assert np.array_equal(labels, ref_labels)
    Xe = Xt[:, idx]
    W2 = out2["W2"].cpu().numpy()[0]      # [G, ldz]: weights after the second update = pass-3 weights
    B3 = out3["B"].cpu().numpy()[0]       # [p_ext, ldz]
    for k, ci in enumerate(cells):
        pen = R.Penalty(labels, np.zeros(len(idx)), W2[:, k].copy(), np.zeros(G))
        cert = R.certificate(Xe, yt, B3[:, k].copy(), pen)
        assert cert["gap"] <= 1e-8 * abs(cert["primal"]), (ci, cert)
        # and the folded coefficients reproduce the cell's predictions (fold-back :492-501)
        coef = R.fold_back(B3[:, k], idx, p)
        np.testing.assert_allclose(coef, out3["coef"].cpu().numpy()[0, :, k], rtol=0, atol=1e-12 * np.abs(coef).max())


def _ridged_gap_torch(X, y, beta, labels, alpha, delta):
    """Duality gap of the ridged group Lasso from X and y directly (plain torch, FP64): primal
    1/(2n)||y - Xb||^2 + alpha sum ||b_g|| + delta/2 ||b||^2, dual point s*r/n with the residual of
    the ridge-augmented system, s = min(1, 1/max_g ||g_g|| / alpha)  (SURVEY appendix A.10)."""
    import torch

    n = X.shape[0]
    G = int(labels.max().item()) + 1
    r = y - X @ beta
    g = X.T @ r / n - delta * beta
    gn = torch.sqrt(torch.zeros(G, dtype=X.dtype, device=X.device).index_add_(0, labels, g * g))
    bn = torch.sqrt(torch.zeros(G, dtype=X.dtype, device=X.device).index_add_(0, labels, beta * beta))
    rr_aug = (r @ r) + n * delta * (beta @ beta)
    P = rr_aug / (2 * n) + alpha * bn.sum()
    omega = (gn / alpha).max()
    s = 1.0 / omega if omega > 1 else torch.ones((), dtype=X.dtype, device=X.device)
    D = (2 * s * (y @ r) - s * s * rr_aug) / (2 * n)
    return float(P), float(P - D)


def test_tall_design_ridged_group_lasso():
    """C5's shape class (n >> p, RidgedGroupLasso, delta = 1): a reduced tall design against the CPU
    oracle cell by cell, and p = 8000 (the full C5 width, 100k rows) against a duality gap computed
    from X and y on the device with plain torch operations."""
    import torch
    from sklearn.model_selection import KFold

    from sparselm_b200.model import RidgedGroupLasso
    from sparselm_b200.model_selection import GridSearchCV

    # (a) n = 30000, p = 512 against the oracle
    rng = np.random.default_rng(31)
    n, p, G = 30000, 512, 64
    X = rng.standard_normal((n, p))
    w = np.zeros(p)
    w[rng.choice(p, 40, replace=False)] = 3 * rng.standard_normal(40)
    y = X @ w + 2.0 * rng.standard_normal(n)
    groups = rng.permutation(np.repeat(np.arange(G), p // G))
    amax = np.abs(X.T @ y).max() / n
    alphas = amax * np.logspace(0.3, -2, 6)
    gs = GridSearchCV(RidgedGroupLasso(groups=groups, delta=(1.0,), solver_options={"tol": 1e-12}),
                      {"alpha": list(alphas)}, cv=3).fit(X, y)
    assert gs.batched_ and (gs.solver_info_["status"] == 0).all()
    folds = list(KFold(3).split(X))
    for ci in (0, 3, 5):
        for f in (0, 2):
            tr, te = folds[f]
            b, _ = R.fit("RidgedGroupLasso", X[tr], y[tr], alpha=alphas[ci], groups=groups, delta=(1.0,))
            ref = -np.sqrt(np.mean((y[te] - X[te] @ b) ** 2))
            assert gs.cv_results_[f"split{f}_test_score"][ci] == pytest.approx(ref, rel=1e-8)
    b, _ = R.fit("RidgedGroupLasso", X, y, alpha=gs.best_params_["alpha"], groups=groups, delta=(1.0,))
    assert np.abs(gs.best_estimator_.coef_ - b).max() <= 1e-6 * np.abs(b).max()

    # (b) full width p = 8000: certificate from X and y on the device
    dev = torch.device("cuda", 0)
    n, p, G = 100000, 8000, 400
    gen = torch.Generator(device=dev).manual_seed(7)
    Xd = torch.empty((n, p), dtype=torch.float64, device=dev)
    for r0 in range(0, n, 1 << 14):
        Xd[r0:r0 + (1 << 14)].normal_(generator=gen)
    wd = torch.zeros(p, dtype=torch.float64, device=dev)
    sel = torch.randperm(p, generator=gen, device=dev)[: p // 10]
    wd[sel] = 100.0 * torch.rand(p // 10, generator=gen, device=dev, dtype=torch.float64)
    yd = Xd @ wd + 10.0 * torch.randn(n, generator=gen, device=dev, dtype=torch.float64)
    groups = np.random.default_rng(3).permutation(np.repeat(np.arange(G), p // G))
    labels = torch.from_numpy(np.unique(groups, return_inverse=True)[1].astype(np.int64)).to(dev)
    amax = float((Xd.T @ yd).abs().max().item()) / n
    tr = torch.arange(n // 2, n, device=dev)  # training rows of fold 0 of KFold(2)
    from sparselm_b200.engine import get_engine
    from sparselm_b200.model._base import solve_specs

    engine = get_engine(0)
    work = RidgedGroupLasso(groups=groups, delta=(1.0,), solver_options={"tol": 1e-10})
    cell_alphas = [amax * 10 ** e for e in (-0.5, -1.5, -2.5)]
    specs = []
    for a in cell_alphas:
        work.set_params(alpha=a)
        specs.append(work._problem_spec(p))
    fd = engine.prepare(Xd, yd.cpu().numpy(), [np.arange(0, n // 2), np.arange(n // 2, n)], False, None,
                        col_perm=specs[0].col_perm)
    out = solve_specs(engine, fd, [specs, []], tol=1e-10)
    assert (out["status"][0, :3] == 0).all()
    cp = specs[0].col_perm if specs[0].col_perm is not None else np.arange(p)
    perm = torch.from_numpy(np.asarray(cp, dtype=np.int64)).to(dev)
    Xt, yt = Xd[tr], yd[tr]
    for k, a in enumerate(cell_alphas):
        beta = torch.zeros(p, dtype=torch.float64, device=dev)
        beta[perm] = out["coef"][0, :, k]
        P, gap = _ridged_gap_torch(Xt, yt, beta, labels, a, 1.0)
        assert gap <= 1e-8 * abs(P), (a, P, gap)
    del Xd, Xt

"""GPU parity tests of the CUDA engine, kernel by kernel and end to end, through
the C ABI (ctypes).  Checker = oracle/ (CPU) and torch fp64 for plain GEMMs."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import engine_model as M  # noqa: E402
import oracle.reference as R  # noqa: E402


def _torch():
    import torch

    return torch


def _rng(seed=0):
    return np.random.default_rng(seed)


# --------------------------------------------------------------------------- #
# K1: pack + Gram build + complement + centering
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("n,p,F", [(37, 5, 1), (300, 81, 3), (1000, 250, 5), (513, 130, 4)])
def test_gram_build_matches_fp64_reference(engine, n, p, F):
    torch = _torch()
    rng = _rng(n + p)
    X = rng.standard_normal((n, p))
    y = rng.standard_normal(n)
    Xa = engine.pack(X, y)
    pa = engine.padded_cols(p)
    Xa_ref = np.zeros((n, pa))
    Xa_ref[:, :p], Xa_ref[:, p], Xa_ref[:, p + 1] = X, y, 1.0
    np.testing.assert_array_equal(Xa.cpu().numpy(), Xa_ref)

    row_ptr = np.linspace(0, n, F + 1).astype(np.int64)
    G = engine.gram_blocks(Xa, row_ptr, extra=1)
    torch.cuda.synchronize()
    for f in range(F):
        blk = Xa_ref[row_ptr[f]:row_ptr[f + 1]]
        ref = blk.T @ blk
        got = G[f].cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-12 * np.abs(ref).max())
        assert np.array_equal(got, got.T), "Gram must be exactly symmetric"
    if F > 1:
        blocks = G[:F].cpu().numpy().copy()
        engine.gram_complement(G, F, out=G[F])
        tot = blocks.sum(0)
        np.testing.assert_allclose(G[F].cpu().numpy(), tot, rtol=1e-14, atol=0)
        for f in range(F):
            np.testing.assert_allclose(G[f].cpu().numpy(), tot - blocks[f], rtol=1e-12,
                                       atol=1e-12 * np.abs(tot).max())


@pytest.mark.parametrize("F", [1, 3])
@pytest.mark.parametrize("fit_intercept", [False, True])
def test_pipelined_host_prepare_accumulates_gram_blocks(engine, F, fit_intercept):
    """A host-resident design travels in row blocks; every block enters its Gram as soon as it is
    packed (slm_gram_block_add): same Grams as numpy, exactly symmetric."""
    rng = _rng(17 + F)
    n, p = 5000, 70
    X = rng.standard_normal((n, p)) + 0.3
    y = rng.standard_normal(n)
    folds = None if F == 1 else [np.arange(f * n // F, (f + 1) * n // F) for f in range(F)]
    old = engine.PIPELINE_BLOCK_BYTES
    engine.PIPELINE_BLOCK_BYTES = 1 << 16  # floor of 1024 rows per block: five blocks
    try:
        fd = engine.prepare(X, y, folds, fit_intercept, None)
    finally:
        engine.PIPELINE_BLOCK_BYTES = old
    fd.check_finite()

    def gram(rows):
        Xr, yr = X[rows], y[rows]
        if fit_intercept:
            Xr, yr = Xr - Xr.mean(0), yr - yr.mean()
        return Xr.T @ Xr, Xr.T @ yr

    G, c = gram(np.arange(n))
    got = fd.G_full.cpu().numpy()
    np.testing.assert_allclose(got[:p, :p], G, rtol=1e-12, atol=1e-11 * np.abs(G).max())
    np.testing.assert_allclose(got[p, :p], c, rtol=1e-12, atol=1e-11 * np.abs(G).max())
    assert np.array_equal(got, got.T)
    for f in range(F if F > 1 else 0):
        Gf, cf = gram(np.setdiff1d(np.arange(n), folds[f]))
        gf = fd.G_train[f].cpu().numpy()
        np.testing.assert_allclose(gf[:p, :p], Gf, rtol=1e-12, atol=1e-11 * np.abs(Gf).max())
        np.testing.assert_allclose(gf[p, :p], cf, rtol=1e-12, atol=1e-11 * np.abs(Gf).max())
        assert np.array_equal(gf, gf.T)


def test_pack_with_permutations(engine):
    rng = _rng(5)
    n, p = 50, 13
    X = rng.standard_normal((n, p))
    y = rng.standard_normal(n)
    cp = rng.permutation(p)
    rp = rng.permutation(n)
    Xa = engine.pack(X, y, col_perm=cp, row_perm=rp).cpu().numpy()
    np.testing.assert_array_equal(Xa[:, :p], X[rp][:, cp])
    np.testing.assert_array_equal(Xa[:, p], y[rp])


def test_gram_center(engine):
    rng = _rng(7)
    n, p = 200, 33
    X = rng.random((n, p)) + 3.0
    y = rng.standard_normal(n) + 10.0
    Xa = engine.pack(X, y)
    G = engine.gram_blocks(Xa, np.array([0, n]))
    engine.gram_center(G, p)
    got = G[0].cpu().numpy()
    Xc, yc = X - X.mean(0), y - y.mean()
    np.testing.assert_allclose(got[:p, :p], Xc.T @ Xc, rtol=0, atol=1e-9)
    np.testing.assert_allclose(got[p, :p], Xc.T @ yc, rtol=0, atol=1e-9)
    np.testing.assert_allclose(got[p, p], yc @ yc, rtol=1e-10)
    np.testing.assert_allclose(got[p + 1, :p], X.sum(0), rtol=1e-13)  # ones row intact


# --------------------------------------------------------------------------- #
# K5: batched tensor-core apply, every tile shape
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("p,Ks", [(80, [1]), (80, [10, 7]), (515, [100, 100, 100]), (1030, [33]),
                                  (2048, [104] * 5), (1200, [128, 60, 17, 8]), (300, [250, 3])])
def test_gram_apply_matches_matmul(engine, p, Ks):
    torch = _torch()
    rng = _rng(p)
    F = len(Ks)
    pa = engine.padded_cols(p)
    A = rng.standard_normal((F, pa, pa))
    A = A + A.transpose(0, 2, 1)
    G = torch.from_numpy(A).cuda()
    ldz = max(8, (max(Ks) + 7) // 8 * 8)
    Zh = np.zeros((F, p, ldz))
    for f, k in enumerate(Ks):
        Zh[f, :, :k] = rng.standard_normal((p, k))
    Z = torch.from_numpy(Zh).cuda()
    GZ = engine.gram_apply(G, p, Ks, Z).cpu().numpy()
    for f, k in enumerate(Ks):
        ref = A[f, :p, :p] @ Zh[f]
        np.testing.assert_allclose(GZ[f][:, :k], ref[:, :k], rtol=0, atol=1e-11 * np.abs(ref).max())


@pytest.mark.parametrize("p,Ks,chunk_w,density", [
    (80, [10, 7], 32, 0.3), (515, [100, 100, 100], 32, 0.05), (1030, [33], 8, 0.5), (2048, [104] * 5, 32, 0.1),
    (1200, [128, 60, 17, 8], 24, 0.02), (300, [250, 3], 64, 1.0), (700, [40, 0, 40], 16, 0.0),
    (900, [100] * 9, 32, 0.2),  # 36 problems: more than one launch group
])
def test_rowsparse_apply_matches_dense(engine, p, Ks, chunk_w, density):
    """Row-sparse apply (device-built support lists per column chunk, device-side stream-K
    partition) == dense product; rows that are zero in one chunk only, empty supports and
    fully dense iterates included."""
    torch = _torch()
    rng = _rng(p + chunk_w)
    F = len(Ks)
    pa = engine.padded_cols(p)
    A = rng.standard_normal((F, pa, pa))
    A = A + A.transpose(0, 2, 1)
    G = torch.from_numpy(A).cuda()
    ldz = max(8, (max(Ks) + 7) // 8 * 8)
    Zh = np.zeros((F, p, ldz))
    for f, k in enumerate(Ks):
        vals = rng.standard_normal((p, k))
        # row sparsity that differs between column chunks, plus scattered single entries
        for c0 in range(0, k, chunk_w):
            keep = rng.random(p) < density
            vals[~keep, c0:c0 + chunk_w] = 0.0
        if k and density < 1.0:
            vals[rng.integers(p, size=3), rng.integers(k, size=3)] = 1.5
        Zh[f, :, :k] = vals
    # columns >= K[f] of Z hold garbage that must not enter the lists' products' results
    Zdev = Zh.copy()
    for f, k in enumerate(Ks):
        Zdev[f, :, k:] = 7.0
    Z = torch.from_numpy(Zdev).cuda()
    GZ = engine.gram_apply_rowsparse(G, p, Ks, Z, chunk_w=chunk_w).cpu().numpy()
    GZd = engine.gram_apply(G, p, Ks, Z).cpu().numpy()
    for f, k in enumerate(Ks):
        ref = A[f, :p, :p] @ Zh[f]
        scale = max(np.abs(ref).max(), 1.0)
        np.testing.assert_allclose(GZ[f][:, :k], ref[:, :k], rtol=0, atol=1e-11 * scale)
        np.testing.assert_allclose(GZ[f][:, :k], GZd[f][:, :k], rtol=0, atol=1e-11 * scale)
    # bit-reproducible run to run (fixed-order stream-K fix-up)
    GZ2 = engine.gram_apply_rowsparse(G, p, Ks, Z, chunk_w=chunk_w).cpu().numpy()
    for f, k in enumerate(Ks):
        assert np.array_equal(GZ[f][:, :k], GZ2[f][:, :k])


def test_lipschitz_is_tight_lower_bound(engine):
    torch = _torch()
    rng = _rng(3)
    for n, p in [(400, 100), (3000, 700), (60, 60), (6000, 1500), (500, 1000)]:
        X = rng.standard_normal((n, p))
        Xa = engine.pack(X, np.zeros(n))
        G = engine.gram_blocks(Xa, np.array([0, n]))
        lam = engine.lipschitz(G, p)[0]
        true = np.linalg.eigvalsh(X.T @ X)[-1]
        assert lam <= true * (1 + 1e-12)
        assert lam * engine.LIPSCHITZ_MARGIN >= true * 1.01, (lam, true)


# --------------------------------------------------------------------------- #
# K6/K7: one prox step and the gap against the numpy model of the kernels
# --------------------------------------------------------------------------- #
def _problem(n, p, Gn, seed, noise=1.0):
    rng = _rng(seed)
    X = rng.standard_normal((n, p))
    w = np.zeros(p)
    w[rng.choice(p, max(1, p // 10), replace=False)] = 10 * rng.random(max(1, p // 10))
    y = X @ w + noise * rng.standard_normal(n)
    sizes = rng.multinomial(p - Gn, np.ones(Gn) / Gn) + 1
    gptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    return X, y, gptr


def _grid(kind, p, gptr, alphas, rng):
    from sparselm_b200.engine import PenaltyGrid

    K = len(alphas)
    Gn = len(gptr) - 1
    gw = 0.5 + rng.random(Gn)
    if kind == "lasso":
        return PenaltyGrid(p=p, lam1=alphas)
    if kind == "group":
        return PenaltyGrid(p=p, lam1=np.zeros(K), gptr=gptr, W2=gw[:, None] * alphas[None, :])
    if kind == "sgl":
        return PenaltyGrid(p=p, lam1=0.5 * alphas, gptr=gptr, W2=gw[:, None] * (0.5 * alphas)[None, :])
    if kind == "ridged":
        return PenaltyGrid(p=p, lam1=np.zeros(K), gptr=gptr, W2=gw[:, None] * alphas[None, :],
                           D2=np.tile((0.1 + rng.random(Gn))[:, None], (1, K)))
    raise ValueError(kind)


def _oracle_pen(grid, k):
    p = grid.p
    gptr = np.arange(p + 1) if grid.gptr is None else grid.gptr
    Gn = len(gptr) - 1
    labels = np.repeat(np.arange(Gn), np.diff(gptr))
    w2 = np.zeros(Gn) if grid.W2 is None else grid.W2[:, k]
    d2 = np.zeros(Gn) if grid.D2 is None else grid.D2[:, k]
    return R.Penalty(labels, np.full(p, grid.lam1[k]), w2, d2)


@pytest.mark.parametrize("kind", ["lasso", "group", "sgl", "ridged"])
def test_solve_matches_oracle(engine, kind):
    n, p, Gn = 300, 90, 12
    X, y, gptr = _problem(n, p, Gn, seed=11)
    rng = _rng(1)
    amax = np.abs(X.T @ y).max() / n
    alphas = amax * np.logspace(0.1, -3, 13)
    grid = _grid(kind, p, gptr, alphas, rng)
    fd = engine.prepare(X, y)
    res = engine.solve(fd.G_full[None], p, [n], [fd.lipschitz(engine, "full")], [grid], tol=1e-12)
    B = res["B"][0].cpu().numpy()
    assert (res["status"][0, : grid.K] == 0).all(), res["status"]
    for k in range(grid.K):
        pen = _oracle_pen(grid, k)
        b_ref, info = R.solve(X, y, pen, tol=1e-14)
        scale = max(np.abs(b_ref).max(), 1e-12)
        assert np.abs(B[:, k] - b_ref).max() <= 1e-6 * scale, (kind, k, np.abs(B[:, k] - b_ref).max(), scale)
        assert np.array_equal(np.abs(B[:, k]) > 1e-6, np.abs(b_ref) > 1e-6)
        obj, obj_ref = R.objective(X, y, B[:, k], pen), R.objective(X, y, b_ref, pen)
        assert abs(obj - obj_ref) <= 1e-8 * max(abs(obj_ref), 1e-300)
        # reported primal/gap agree with an independent certificate from X, y
        cert = R.certificate(X, y, B[:, k], pen)
        assert abs(res["primal"][0, k] - cert["primal"]) <= 1e-9 * abs(cert["primal"])
        assert res["gap"][0, k] <= 1e-11 * abs(cert["primal"]) + 1e-12 * (y @ y) / (2 * n)


def test_solve_multi_fold_batch_and_model_iterations(engine):
    """5 folds x 20 columns in one batch; compare with the numpy model of the kernels."""
    n, p, Gn = 500, 120, 15
    X, y, gptr = _problem(n, p, Gn, seed=2)
    rng = _rng(2)
    folds = np.array_split(np.arange(n), 5)
    fd = engine.prepare(X, y, test_folds=folds)
    amax = np.abs(X.T @ y).max() / n
    alphas = amax * np.logspace(0, -2.5, 20)
    grid = _grid("sgl", p, gptr, alphas, rng)
    res = engine.solve(fd.G_train, p, fd.n_train, fd.lipschitz(engine, list(range(5))), [grid] * 5, tol=1e-11)
    B = res["B"].cpu().numpy()
    assert (res["status"][:, :20] == 0).all()
    for f in range(5):
        tr = np.setdiff1d(np.arange(n), folds[f])
        Xt, yt = X[tr], y[tr]
        G = Xt.T @ Xt
        # Gram of the fold agrees with a direct build
        got = fd.G_train[f].cpu().numpy()
        np.testing.assert_allclose(got[:p, :p], G, rtol=0, atol=1e-10 * np.abs(G).max())
        pb = M.BatchProblem(G, Xt.T @ yt, yt @ yt, len(tr), gptr, np.tile(grid.lam1, (p, 1)), grid.W2,
                            np.zeros_like(grid.W2), L=fd.lipschitz(engine, f))
        Bm, info = M.solve(pb, tol=1e-11)
        # same algorithm => same answer; p <= 160 runs the fused small-design kernel, whose
        # convergence checks sit at iterations 0, 10, 30, 70, 150, ...: a column is flagged at the
        # first check at or after the model's (which checks every 10 iterations)
        assert np.abs(B[f][:, :20] - Bm).max() <= 1e-7 * np.abs(Bm).max()
        checks = np.array([0, 10, 30, 70, 150, 310, 630])
        expect = checks[np.searchsorted(checks, info["iters"] - 10)]
        assert (res["n_iter"][f, :20] >= expect).all() and (res["n_iter"][f, :20] <= 2 * info["iters"] + 20).all()
        for k in (0, 10, 19):
            b_ref, _ = R.solve(Xt, yt, _oracle_pen(grid, k), tol=1e-14)
            assert np.abs(B[f][:, k] - b_ref).max() <= 1e-6 * max(np.abs(b_ref).max(), 1e-12)


def test_adaptive_passes_match_oracle(engine):
    from sparselm_b200.engine import PenaltyGrid

    n, p, Gn = 200, 40, 8
    X, y, gptr = _problem(n, p, Gn, seed=4)
    rng = _rng(4)
    alphas = np.array([0.05, 0.3, 1.0])
    K = len(alphas)
    gw = 0.5 + rng.random(Gn)
    labels = np.repeat(np.arange(Gn), np.diff(gptr))
    fd = engine.prepare(X, y)
    # AdaptiveLasso
    grid = PenaltyGrid(p=p, lam1=alphas, adaptive=dict(a1=alphas, a2=None, alpha=alphas, gw=None, eps=1e-6,
                                                       tol=1e-10, max_iter=3, update_function=None))
    res = engine.solve(fd.G_full[None], p, [n], [fd.lipschitz(engine, "full")], [grid], tol=1e-12)
    B = res["B"][0].cpu().numpy()
    for k, a in enumerate(alphas):
        b_ref, _ = R.fit("AdaptiveLasso", X, y, alpha=a)
        assert np.abs(B[:, k] - b_ref).max() <= 1e-6 * np.abs(b_ref).max(), (k, np.abs(B[:, k] - b_ref).max())
    # AdaptiveSparseGroupLasso
    l1r = 0.3
    grid = PenaltyGrid(p=p, lam1=l1r * alphas, gptr=gptr, W2=np.tile(((1 - l1r) * alphas)[None, :], (Gn, 1)),
                       adaptive=dict(a1=l1r * alphas, a2=(1 - l1r) * alphas, alpha=alphas, gw=gw, eps=1e-6,
                                     tol=1e-10, max_iter=3, update_function=None))
    res = engine.solve(fd.G_full[None], p, [n], [fd.lipschitz(engine, "full")], [grid], tol=1e-12)
    B = res["B"][0].cpu().numpy()
    for k, a in enumerate(alphas):
        b_ref, _ = R.fit("AdaptiveSparseGroupLasso", X, y, alpha=a, groups=labels, group_weights=gw, l1_ratio=l1r)
        assert np.abs(B[:, k] - b_ref).max() <= 1e-6 * np.abs(b_ref).max(), (k, np.abs(B[:, k] - b_ref).max())
    assert (res["n_pass"][0, :K] == 3).all()


def test_cv_score_and_intercepts(engine):
    torch = _torch()
    rng = _rng(9)
    n, p, K = 700, 45, 11
    X = rng.standard_normal((n, p)) + 1.0
    y = rng.standard_normal(n) + 2.0
    Xa = engine.pack(X, y)
    ldz = 16
    Bh = np.zeros((p, ldz))
    Bh[:, :K] = rng.standard_normal((p, K))
    B = torch.from_numpy(Bh).cuda()
    G = engine.gram_blocks(Xa, np.array([0, n]))
    icpt = engine.intercepts(G[0], p, B, K)
    ref_icpt = y.mean() - X.mean(0) @ Bh[:, :K]
    np.testing.assert_allclose(icpt.cpu().numpy()[:K], ref_icpt, rtol=1e-11, atol=1e-11)
    r0, r1 = 100, 433
    out = engine.cv_score(Xa, p, r0, r1, B, K, icpt).cpu().numpy()
    resid = y[r0:r1, None] - X[r0:r1] @ Bh[:, :K] - ref_icpt[None, :]
    np.testing.assert_allclose(out[0, :K], (resid ** 2).sum(0), rtol=1e-11)
    np.testing.assert_allclose(out[1, :K], np.abs(resid).sum(0), rtol=1e-11)
    # rows packed with sample weights: the scorer still sees unweighted residuals
    sw = 0.3 + rng.random(n)
    Xw = engine.pack(X, y, sample_weight=sw)
    outw = engine.cv_score(Xw, p, r0, r1, B, K, icpt, rows_scaled=True).cpu().numpy()
    np.testing.assert_allclose(outw[0, :K], (resid ** 2).sum(0), rtol=1e-11)
    np.testing.assert_allclose(outw[1, :K], np.abs(resid).sum(0), rtol=1e-11)


def test_overlap_gather_and_fold_back(engine):
    torch = _torch()
    rng = _rng(12)
    n, p = 120, 17
    X = rng.standard_normal((n, p))
    y = rng.standard_normal(n)
    group_list = [list(rng.choice(5, size=rng.integers(1, 3), replace=False)) for _ in range(p)]
    idx, ext_groups, Gn = R.expand_overlap(group_list, p)
    pe = len(idx)
    fd = engine.prepare(X, y)
    idx_dev = torch.from_numpy(idx.astype(np.int32)).cuda()
    Ge = engine.gram_gather(fd.G_full, p, idx_dev, pe)[0].cpu().numpy()
    Xe = X[:, idx]
    np.testing.assert_allclose(Ge[:pe, :pe], Xe.T @ Xe, rtol=0, atol=1e-11 * n)
    np.testing.assert_allclose(Ge[pe, :pe], Xe.T @ y, rtol=0, atol=1e-11 * n)
    np.testing.assert_allclose(Ge[pe, pe], y @ y, rtol=1e-13)
    # fold back
    order = np.argsort(idx, kind="stable")
    inv_ptr = np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=p))]).astype(np.int32)
    Be = np.zeros((pe, 8))
    Be[:, :3] = rng.standard_normal((pe, 3))
    coef = engine.fold_back(torch.from_numpy(Be).cuda(), torch.from_numpy(inv_ptr).cuda(),
                            torch.from_numpy(order.astype(np.int32)).cuda(), p, 3).cpu().numpy()
    for k in range(3):
        np.testing.assert_allclose(coef[:, k], R.fold_back(Be[:, k], idx, p), rtol=1e-14, atol=1e-15)


# --------------------------------------------------------------------------- #
# standardize=True: per-group whitening (slm_group_whiten_factors / slm_gram_whiten /
# slm_coef_unwhiten) against numpy Cholesky
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("sizes", [[1, 1, 1, 1, 1], [3, 7, 1, 20, 5], [64, 2, 130]])
@pytest.mark.parametrize("ridged", [False, True])
def test_group_whitening_matches_numpy(engine, sizes, ridged):
    torch = _torch()
    rng = _rng(sum(sizes) + ridged)
    p = int(sum(sizes))
    n = 3 * p + 10
    X = rng.standard_normal((n, p)) @ (np.eye(p) + 0.2 * rng.standard_normal((p, p)))
    y = rng.standard_normal(n)
    Xa = engine.pack(X, y)
    pa = engine.padded_cols(p)
    row_ptr = np.array([0, n // 2, n], dtype=np.int64)
    G = engine.gram_blocks(Xa, row_ptr)
    gptr = np.concatenate([[0], np.cumsum(sizes)])
    Gn = len(sizes)
    dl = 0.1 + rng.random(Gn) if ridged else None
    n_obs = np.array([n // 2, n - n // 2], dtype=float)
    Gw, wctx = engine.whiten(G, p, gptr, n_obs, shift=None if dl is None else np.sqrt(dl), ridge=dl)
    torch.cuda.synchronize()
    Gh, Gwh = G.cpu().numpy(), Gw.cpu().numpy()
    for f in range(2):
        W = np.zeros((pa, pa))
        W[p:, p:] = np.eye(pa - p)
        ridge = np.zeros((pa, pa))
        for g in range(Gn):
            a, b = gptr[g], gptr[g + 1]
            A = Gh[f, a:b, a:b] + (np.sqrt(dl[g]) * np.eye(b - a) if ridged else 0.0)
            Rg = np.linalg.cholesky(A).T  # A = R^T R, R upper
            Wg = np.linalg.inv(Rg)
            W[a:b, a:b] = Wg
            if ridged:
                ridge[a:b, a:b] = n_obs[f] * dl[g] * Wg.T @ Wg
        ref = W.T @ Gh[f] @ W + ridge
        np.testing.assert_allclose(Gwh[f], ref, rtol=1e-10, atol=1e-10 * np.abs(ref).max())
        if not ridged:  # whitened groups are orthonormal
            for g in range(Gn):
                a, b = gptr[g], gptr[g + 1]
                np.testing.assert_allclose(Gwh[f, a:b, a:b], np.eye(b - a), atol=1e-9)
        # coefficients back: b = W gamma
        ldz = 16
        gam = torch.from_numpy(rng.standard_normal((p, ldz))).to(engine.device)
        back = engine.unwhiten(gam, wctx, f, p, 11).cpu().numpy()
        ref_b = W[:p, :p] @ gam.cpu().numpy()
        np.testing.assert_allclose(back[:, :11], ref_b[:, :11], rtol=1e-11, atol=1e-12 * np.abs(ref_b).max())
        assert np.all(back[:, 11:] == 0.0)


def test_group_whitening_rejects_singular_block(engine):
    rng = _rng(77)
    n, p = 40, 6
    X = rng.standard_normal((n, p))
    X[:, 2] = X[:, 0] - X[:, 1]  # group 0 = columns 0..2 is rank deficient
    Xa = engine.pack(X, rng.standard_normal(n))
    G = engine.gram_blocks(Xa, np.array([0, n], dtype=np.int64))
    with pytest.raises(ValueError, match="positive definite"):
        engine.whiten(G, p, np.array([0, 3, 6]), np.array([float(n)]))


def test_tri_pack_roundtrip(engine):
    """Upper-triangle packing of the row-sharded Gram all-reduce: pack -> unpack restores a
    symmetric matrix exactly, and summing packed buffers equals summing the matrices."""
    import ctypes

    torch = _torch()
    rng = _rng(21)
    pa, F = 40, 3
    A = rng.standard_normal((2, F, pa, pa))
    A = A + A.transpose(0, 1, 3, 2)
    tri = int(engine.lib.slm_tri_size(pa))
    assert tri == pa * (pa + 1) // 2
    bufs = []
    for r in range(2):
        G = torch.from_numpy(A[r].copy()).to(engine.device)
        buf = torch.empty((F, tri), dtype=torch.float64, device=engine.device)
        engine._ck(engine.lib.slm_tri_pack(engine.h, engine._ptr(G), pa * pa, pa, F, engine._ptr(buf), engine.stream),
                   "slm_tri_pack")
        bufs.append(buf)
    iu = np.triu_indices(pa)
    np.testing.assert_array_equal(bufs[0].cpu().numpy()[1], A[0, 1][iu])
    total = bufs[0] + bufs[1]
    out = torch.full((F, pa, pa), np.nan, dtype=torch.float64, device=engine.device)
    engine._ck(engine.lib.slm_tri_unpack(engine.h, engine._ptr(total), pa, F, engine._ptr(out), pa * pa, engine.stream),
               "slm_tri_unpack")
    np.testing.assert_array_equal(out.cpu().numpy(), A[0] + A[1])


@pytest.mark.parametrize("kind", ["lasso", "group", "sgl", "ridged", "long_groups"])
def test_fused_iteration_matches_two_kernel_iteration(engine, kind):
    """prox_fused_kernel (the Gram applied to the iterate, momentum formed on the fly; csrc/solver_kernels.cuh)
    walks through the same iterates as prox_main + prox_momentum on the extrapolated point: same restart rule,
    same theta sequence.  Compared on a design large enough for the per-iteration kernels (no small-design /
    cooperative path), over several folds, with columns converging at different checks (compactions) and with
    groups longer than the kernel keeps in registers."""
    from sparselm_b200.engine import PenaltyGrid

    rng = _rng(5)
    n, p, F, K = 600, 360, 3, 40
    X = rng.standard_normal((n, p))
    w = np.zeros(p)
    w[rng.choice(p, 30, replace=False)] = 5 * rng.random(30)
    y = X @ w + rng.standard_normal(n)
    if kind == "long_groups":
        gptr = np.concatenate([[0], np.cumsum([90, 3, 150, 40, 77])]).astype(np.int32)  # > 32 rows: the two-pass path
    else:
        gptr = np.arange(0, p + 1, 9).astype(np.int32)
    Gn = len(gptr) - 1
    amax = np.abs(X.T @ y).max() / n
    alphas = amax * np.geomspace(1.0, 0.004, K)
    gw = np.sqrt(np.diff(gptr)).astype(float)
    if kind == "lasso":
        grid = PenaltyGrid(p=p, lam1=alphas)
    elif kind == "sgl":
        grid = PenaltyGrid(p=p, lam1=0.5 * alphas, gptr=gptr, W2=gw[:, None] * (0.5 * alphas)[None, :])
    elif kind == "ridged":
        grid = PenaltyGrid(p=p, lam1=np.zeros(K), gptr=gptr, W2=gw[:, None] * alphas[None, :], D2=np.full((Gn, K), 0.3))
    else:
        grid = PenaltyGrid(p=p, lam1=np.zeros(K), gptr=gptr, W2=gw[:, None] * alphas[None, :])
    folds = np.array_split(np.arange(n), F)
    fd = engine.prepare(X, y, test_folds=folds)
    L = fd.lipschitz(engine, list(range(F)))
    outs = []
    try:
        for fused in (1, 0):
            engine.set_option("fused_prox", fused)
            outs.append(engine.solve(fd.G_train, p, fd.n_train, L, [grid] * F, tol=1e-10, newton=False))
    finally:
        engine.set_option("fused_prox", 1)
    a, b = outs
    assert (a["status"][:, :K] == 0).all() and (b["status"][:, :K] == 0).all()
    Ba, Bb = a["B"].cpu().numpy()[:, :, :K], b["B"].cpu().numpy()[:, :, :K]
    # same iterates up to rounding (G*W exact vs. the momentum recurrence): identical iteration counts except
    # where a convergence check falls within rounding of the tolerance
    assert np.abs(Ba - Bb).max() <= 1e-8 * np.abs(Bb).max()
    assert np.mean(a["n_iter"][:, :K] == b["n_iter"][:, :K]) >= 0.9
    assert np.abs(a["n_iter"][:, :K] - b["n_iter"][:, :K]).max() <= 20

"""CPU check of the fused iteration's algebra (csrc/solver_kernels.cuh, prox_fused_kernel) on the numpy model of the
kernels: applying the Gram to the iterate and forming the extrapolated point on the fly, with the momentum of
iteration t decided at its start, walks through the same iterates as the two-kernel form that materialises Z and
decides the momentum at the end of iteration t-1.  (The GPU counterpart is
tests/test_gpu_engine.py::test_fused_iteration_matches_two_kernel_iteration.)"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import engine_model as M  # noqa: E402
import oracle.reference as R  # noqa: E402


def _problem(kind, seed):
    rng = np.random.default_rng(seed)
    n, p, Gn, K = 120, 48, 8, 9
    X = rng.standard_normal((n, p))
    w = np.zeros(p)
    w[rng.choice(p, 6, replace=False)] = 4 * rng.random(6)
    y = X @ w + 0.5 * rng.standard_normal(n)
    sizes = rng.multinomial(p - Gn, np.ones(Gn) / Gn) + 1
    gptr = np.concatenate([[0], np.cumsum(sizes)])
    amax = np.abs(X.T @ y).max() / n
    alphas = amax * np.geomspace(0.9, 0.01, K)
    gw = 0.5 + rng.random(Gn)
    W1 = np.zeros((p, K))
    W2 = np.zeros((Gn, K))
    D2 = np.zeros((Gn, K))
    if kind == "lasso":
        gptr = np.arange(p + 1)
        W1 = np.tile(alphas, (p, 1))
        W2 = np.zeros((p, K))
        D2 = np.zeros((p, K))
    elif kind == "group":
        W2 = gw[:, None] * alphas[None, :]
    elif kind == "sgl":
        W1 = np.tile(0.5 * alphas, (p, 1))
        W2 = gw[:, None] * (0.5 * alphas)[None, :]
    elif kind == "ridged":
        W2 = gw[:, None] * alphas[None, :]
        D2 = np.tile((0.1 + rng.random(Gn))[:, None], (1, K))
    pb = M.BatchProblem(X.T @ X, X.T @ y, y @ y, n, gptr, W1, W2, D2)
    return X, y, pb


@pytest.mark.parametrize("kind", ["lasso", "group", "sgl", "ridged"])
@pytest.mark.parametrize("seed", [0, 1])
def test_fused_model_walks_through_the_same_iterates(kind, seed):
    X, y, pb = _problem(kind, seed)
    for max_iter in (1, 2, 3, 7, 25, 60):  # iterate by iterate: momentum, restarts and freezes line up
        Ba, _ = M.solve(pb, tol=1e-11, max_iter=max_iter)
        Bb, _ = M.solve_fused(pb, tol=1e-11, max_iter=max_iter)
        assert np.abs(Ba - Bb).max() <= 1e-11 * max(np.abs(Ba).max(), 1.0), (kind, max_iter)
    Ba, ia = M.solve(pb, tol=1e-11)
    Bb, ib = M.solve_fused(pb, tol=1e-11)
    assert ia["done"].all() and ib["done"].all()
    assert np.abs(Ba - Bb).max() <= 1e-9 * np.abs(Ba).max()
    assert np.mean(ia["iters"] == ib["iters"]) >= 0.75 and np.abs(ia["iters"] - ib["iters"]).max() <= 20
    # and both are the oracle's solution
    gid = np.repeat(np.arange(len(pb.gptr) - 1), np.diff(pb.gptr))
    for k in (0, pb.K // 2, pb.K - 1):
        pen = R.Penalty(gid, pb.W1[:, k], pb.W2[:, k], pb.D2[:, k])
        b_ref, _ = R.solve(X, y, pen, tol=1e-14)
        # the relative duality-gap test bounds the objective, not the coefficients: a nearly-zero solution under
        # the heaviest penalty is only known to sqrt(gap * scale) -- compare on the scale of the grid's solutions
        assert np.abs(Bb[:, k] - b_ref).max() <= 1e-6 * np.abs(Bb).max()

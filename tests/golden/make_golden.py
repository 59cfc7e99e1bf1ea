"""Generates tests/golden/*.json -- the committed parity fixtures.

Sources of truth, in order of authority:
  1. the reference's own known-answer test (tests/test_lasso.py:29-61 of
     CederGroupHub/sparse-lm): copied numbers, not computed;
  2. scikit-learn's coordinate-descent Lasso (the implementation that KAT was
     "borrowed from"), which minimises exactly the reference's Lasso objective
     1/(2n)||y - Xw||^2 + alpha ||w||_1, run to tol=1e-15;
  3. closed forms on orthonormal designs (X^T X = n I) for every penalty.
The reference itself cannot be imported here (cvxpy is not installed), so no vector
below was produced by the reference's cvxpy path: group/sparse-group/ridged/adaptive
VALUES are pinned by closed forms and optimality certificates only.

Run:  python tests/golden/make_golden.py
"""

import json
import os

import numpy as np
from sklearn.datasets import make_regression
from sklearn.linear_model import Lasso

HERE = os.path.dirname(os.path.abspath(__file__))


def soft(v, t):
    return np.sign(v) * np.maximum(np.abs(v) - t, 0)


def main():
    out = {}
    # 1. reference KAT
    out["reference_kat"] = {
        "X": [[-1], [0], [1]], "y": [-1, 0, 1], "T": [[2], [3], [4]],
        "cases": [
            {"alpha": 1e-8, "coef": [1.0], "pred": [2, 3, 4]},
            {"alpha": 0.1, "coef": [0.85], "pred": [1.7, 2.55, 3.4]},
            {"alpha": 0.5, "coef": [0.25], "pred": [0.5, 0.75, 1.0]},
            {"alpha": 1.0, "coef": [0.0], "pred": [0, 0, 0]},
        ],
        "decimal": 6,
    }
    # 2. sklearn Lasso on seeded problems
    cases = []
    for seed, (n, p, ninf, noise) in enumerate([(40, 12, 4, 1.0), (30, 45, 6, 0.5), (80, 25, 10, 5.0)]):
        X, y = make_regression(n_samples=n, n_features=p, n_informative=ninf, noise=noise, random_state=seed,
                               bias=3.0 * seed)
        for fit_intercept in (False, True):
            amax = np.abs((X - X.mean(0) * fit_intercept).T @ (y - y.mean() * fit_intercept)).max() / n
            for frac in (0.5, 0.1, 0.01):
                a = float(frac * amax)
                m = Lasso(alpha=a, fit_intercept=fit_intercept, tol=1e-15, max_iter=2_000_000).fit(X, y)
                cases.append({"seed": seed, "n": n, "p": p, "n_informative": ninf, "noise": noise,
                              "bias": 3.0 * seed, "fit_intercept": fit_intercept, "alpha": a,
                              "coef": m.coef_.tolist(), "intercept": float(m.intercept_)})
    out["sklearn_lasso"] = cases
    # 3. orthonormal design closed forms
    rng = np.random.default_rng(7)
    n, p = 64, 12
    Q, _ = np.linalg.qr(rng.standard_normal((n, p)))
    X = np.sqrt(n) * Q  # X^T X = n I
    y = X @ np.array([3, -2, 0.5, 0, 0, 0, 1.5, 1.0, -0.2, 0, 0, 0.1]) + 0.3 * rng.standard_normal(n)
    z = X.T @ y / n
    groups = np.repeat(np.arange(4), 3)
    gw = np.array([1.0, 0.5, 2.0, 1.5])
    alpha, l1r, delta = 0.4, 0.3, np.array([0.5, 1.0, 0.0, 2.0])

    def gshrink(v, t):
        b = np.zeros_like(v)
        for g in range(4):
            m = groups == g
            nv = np.linalg.norm(v[m])
            b[m] = v[m] * max(0.0, 1 - t[g] / nv) if nv > 0 else 0
        return b

    out["orthonormal"] = {
        "X": X.tolist(), "y": y.tolist(), "groups": groups.tolist(), "group_weights": gw.tolist(),
        "alpha": alpha, "l1_ratio": l1r, "delta": delta.tolist(),
        "Lasso": soft(z, alpha).tolist(),
        "GroupLasso": gshrink(z, alpha * gw).tolist(),
        "SparseGroupLasso": gshrink(soft(z, l1r * alpha), (1 - l1r) * alpha * gw).tolist(),
        "RidgedGroupLasso": (gshrink(z, alpha * gw) / (1 + delta[groups])).tolist(),
    }
    # adaptive passes on the orthonormal design, reference update rules (alpha^2 quirk)
    eps, max_iter = 1e-6, 3
    w = alpha * np.ones(p)
    for _ in range(max_iter):
        b = soft(z, w)
        w = alpha * (alpha / (np.abs(b) + eps))
    out["orthonormal"]["AdaptiveLasso"] = b.tolist()
    v = alpha * np.ones(4)
    for _ in range(max_iter):
        b = gshrink(z, v)
        norms = np.array([np.linalg.norm(b[groups == g]) for g in range(4)])
        v = (alpha * gw) * (alpha / (norms + eps))
    out["orthonormal"]["AdaptiveGroupLasso"] = b.tolist()
    lam1, lam2 = l1r * alpha, (1 - l1r) * alpha
    w, v = lam1 * np.ones(p), lam2 * np.ones(4)
    for _ in range(max_iter):
        b = gshrink(soft(z, w), v)
        norms = np.array([np.linalg.norm(b[groups == g]) for g in range(4)])
        w = lam1 * (alpha / (np.abs(b) + eps))
        v = (lam2 * gw) * (alpha / (norms + eps))
    out["orthonormal"]["AdaptiveSparseGroupLasso"] = b.tolist()
    v = alpha * np.ones(4)
    for _ in range(max_iter):
        b = gshrink(z, v) / (1 + delta[groups])
        norms = np.array([np.linalg.norm(b[groups == g]) for g in range(4)])
        v = (alpha * gw) * (alpha / (norms + eps))
    out["orthonormal"]["AdaptiveRidgedGroupLasso"] = b.tolist()
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()

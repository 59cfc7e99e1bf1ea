"""Generates tests/golden/golden_nlp.json: group-penalty solutions from a THIRD-PARTY solver.

The reference hands every fit to cvxpy, which rewrites the group norms as second-order cone
constraints and calls a conic interior-point solver (src/sparselm/model/_base.py:512-519).
Neither cvxpy nor any conic solver is installed here, so this script writes the same epigraph
formulation by hand and gives it to a general constrained solver the image does have:
scipy.optimize.minimize(method="SLSQP") -- Kraft's sequential quadratic programming code, which
shares nothing with oracle/ (block coordinate descent) or the engine (proximal gradient):

    min_{b,t}  1/(2n)||y - X b||^2 + sum_j l_j |b_j| + sum_g w_g t_g + 1/2 sum_g d_g ||b_g||^2
    s.t.       t_g^2 >= ||b_g||^2,  t_g >= 0          (the cone, squared to be smooth)
    (the l1 term through the split b = b+ - b-, b+, b- >= 0: linear objective, bounds only)

Objective formulas: _lasso.py:109-121 (data term), :267-275 (group), :627-639 (sparse group),
:795-811 (ridged); group order np.unique (:248); default group weights 1 (:233-235).
Overlap: duplicated-column expansion, then the group problem (:440-461), folded back by
summation (:492-501).  Adaptive variants: the reweighting loops of _adaptive_lasso.py:206-232,
:364-374, :712-726 with every pass solved by SLSQP.

SLSQP is restarted from its own answer until the objective stops decreasing (it can stop early
at the apex of a cone).  Each record stores the objective reached; tests/test_oracle.py compares
coefficients at 1e-6 * ||b||_inf (north_star's coefficient tolerance) and the objective at 1e-9.

Run:  python tests/golden/make_golden_nlp.py        (seconds)
"""
import json
import os

import numpy as np
from scipy.optimize import minimize

HERE = os.path.dirname(os.path.abspath(__file__))


def _quad(b, m, Ms, g):
    """b_g' A_g b_g, A_g = I when no metric is given (standardize=False)."""
    return b[m] @ b[m] if Ms is None else b[m] @ Ms[g] @ b[m]


def _mvec(b, m, Ms, g):
    return b[m] if Ms is None else Ms[g] @ b[m]


def conic_solve(X, y, labels, lam1, w, d, w1=None, Ms=None):
    """labels: (p,) group index 0..G-1; w, d: (G,); w1: (p,) l1 weights or None (= lam1); Ms: per-group
    metric A_g of the norm ||b_g||_A = sqrt(b_g' A_g b_g) (standardize=True) or None.
    Returns (coef, OptimizeResult of the last SLSQP run)."""
    n, p = X.shape
    G = len(w)
    use_l1 = (w1 is not None and np.any(np.asarray(w1) > 0)) or lam1 > 0
    l1w = np.zeros(p) if not use_l1 else (np.full(p, lam1) if w1 is None else np.asarray(w1, float))
    nv = (2 * p if use_l1 else p) + G
    off = nv - G
    A = X.T @ X / n
    c = X.T @ y / n
    masks = [np.flatnonzero(labels == g) for g in range(G)]

    def unpack(z):
        return (z[:p] - z[p:2 * p], z[off:]) if use_l1 else (z[:p], z[off:])

    def f(z):
        b, t = unpack(z)
        v = 0.5 * b @ A @ b - c @ b + 0.5 * (y @ y) / n + w @ t + 0.5 * (d[labels] * b * b).sum()
        return v + (l1w @ (z[:p] + z[p:2 * p]) if use_l1 else 0.0)

    def grad(z):
        b, _ = unpack(z)
        gb = A @ b - c + d[labels] * b
        return np.concatenate([gb + l1w, -gb + l1w, w]) if use_l1 else np.concatenate([gb, w])

    def cone(z):
        b, t = unpack(z)
        return np.array([t[g] ** 2 - _quad(b, m, Ms, g) for g, m in enumerate(masks)])

    def cone_jac(z):
        b, t = unpack(z)
        J = np.zeros((G, nv))
        for g, m in enumerate(masks):
            J[g, m] = -2.0 * _mvec(b, m, Ms, g)
            if use_l1:
                J[g, p + m] = 2.0 * _mvec(b, m, Ms, g)
            J[g, off + g] = 2.0 * t[g]
        return J

    b0 = np.linalg.lstsq(X, y, rcond=None)[0] * 0.5
    t0 = [np.sqrt(_quad(b0, m, Ms, g)) + 0.1 for g, m in enumerate(masks)]
    if use_l1:
        z = np.concatenate([np.maximum(b0, 0), np.maximum(-b0, 0), t0])
        bounds = [(0, None)] * nv
    else:
        z = np.concatenate([b0, t0])
        bounds = [(None, None)] * p + [(0, None)] * G
    cons = [{"type": "ineq", "fun": cone, "jac": cone_jac}]
    best = np.inf
    for attempt in range(60):       # SLSQP sometimes stops early at a kink: restart it from its own answer
        res = minimize(f, z, jac=grad, bounds=bounds, constraints=cons, method="SLSQP",
                       options={"ftol": 1e-16, "maxiter": 3000})
        z = res.x.copy()
        b, t = unpack(z)
        z[off:] = [np.sqrt(_quad(b, m, Ms, g)) for g, m in enumerate(masks)]   # tight epigraph variables
        if use_l1:                                               # complementary split
            z[:p], z[p:2 * p] = np.maximum(b, 0), np.maximum(-b, 0)
        val = f(z)
        if best - val <= 1e-15 * max(1.0, abs(val)) and attempt >= 2:
            break
        best = min(best, val)
        z[off:] += 1e-3 * (1 + np.arange(G)) / G                 # step off the cone's apex before restarting
    return unpack(z)[0].copy(), res


def smooth_solve(X, y, labels, w, d, l1w=None, Ms=None):
    """Homotopy on the smoothed norms sqrt(||b_g||^2 + e^2) and sqrt(b_j^2 + e^2), e = 1e-1 ... 1e-9, every
    stage an exact-Hessian trust-region Newton solve (scipy 'trust-exact'), warm-started."""
    n, p = X.shape
    G = len(w)
    A = X.T @ X / n
    c = X.T @ y / n
    masks = [np.flatnonzero(labels == g) for g in range(G)]
    b = np.linalg.lstsq(X, y, rcond=None)[0] * 0.5
    l1w = np.zeros(p) if l1w is None else l1w
    for e in 10.0 ** -np.arange(1, 10):
        def f(b):
            return 0.5 * b @ A @ b - c @ b + 0.5 * (d[labels] * b * b).sum() + sum(
                w[g] * np.sqrt(_quad(b, m, Ms, g) + e * e) for g, m in enumerate(masks)) + l1w @ np.sqrt(b * b + e * e)

        def grad(b):
            g_ = A @ b - c + d[labels] * b + l1w * b / np.sqrt(b * b + e * e)
            for g, m in enumerate(masks):
                g_[m] += w[g] * _mvec(b, m, Ms, g) / np.sqrt(_quad(b, m, Ms, g) + e * e)
            return g_

        def hess(b):
            H = A + np.diag(d[labels] + l1w * e * e / (b * b + e * e) ** 1.5)
            for g, m in enumerate(masks):
                r = np.sqrt(_quad(b, m, Ms, g) + e * e)
                Ag, v = (np.eye(len(m)) if Ms is None else Ms[g]), _mvec(b, m, Ms, g)
                H[np.ix_(m, m)] += w[g] * (Ag / r - np.outer(v, v) / r**3)
            return H

        b = minimize(f, b, jac=grad, hess=hess, method="trust-exact", options={"gtol": 1e-13, "maxiter": 2000}).x
    # what the smoothing left at O(e) is exactly zero in the cone program
    b[(np.abs(b) < 1e-7) & (l1w > 0)] = 0.0
    for g, m in enumerate(masks):
        if np.sqrt(_quad(b, m, Ms, g)) < 1e-7 * (1.0 if Ms is None else np.sqrt(np.trace(Ms[g]) / len(m))):
            b[m] = 0.0
    return b


def kkt_violation(X, y, b, labels, l1w, w, d, Ms=None):
    """Largest violation of the optimality conditions of the penalised problem (plain numpy)."""
    n = len(y)
    g_ = X.T @ (X @ b - y) / n
    worst = 0.0
    for g in range(len(w)):
        m = np.flatnonzero(labels == g)
        nr = np.sqrt(_quad(b, m, Ms, g))
        if nr > 0:
            r = g_[m] + d[g] * b[m] + w[g] * _mvec(b, m, Ms, g) / nr
            nz = b[m] != 0
            worst = max(worst, np.abs(r[nz] + l1w[m][nz] * np.sign(b[m][nz])).max(initial=0.0),
                        np.maximum(np.abs(r[~nz]) - l1w[m][~nz], 0).max(initial=0.0))
        else:
            s = np.sign(g_[m]) * np.maximum(np.abs(g_[m]) - l1w[m], 0)
            dual = np.linalg.norm(s) if Ms is None else np.sqrt(s @ np.linalg.solve(Ms[g], s))   # dual norm of ||.||_A
            worst = max(worst, dual - w[g])
    return float(worst)


def objective(X, y, b, labels, lam1, w, d, w1=None, Ms=None):
    n = len(y)
    G = len(w)
    r = y - X @ b
    nr = np.array([np.sqrt(_quad(b, np.flatnonzero(labels == g), Ms, g)) for g in range(G)])
    l1 = lam1 * np.abs(b).sum() if w1 is None else (w1 * np.abs(b)).sum()
    n2 = np.array([b[labels == g] @ b[labels == g] for g in range(G)])       # the ridge is on the plain norm
    return float(r @ r / (2 * n) + l1 + w @ nr + 0.5 * (d * n2).sum())


def problem(seed, n, p, G, noise):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    labels = rng.permutation(np.arange(p) % G)
    beta = np.zeros(p)
    for g in rng.choice(G, max(1, G // 2), replace=False):
        beta[labels == g] = rng.standard_normal((labels == g).sum()) * 2.0
    y = X @ beta + noise * rng.standard_normal(n)
    return X, y, labels


def solve_checked(X, y, labels, l1w, w, d, Ms=None):
    """Best of the third-party solves + its KKT violation relative to ||X'y/n||_inf."""
    cands = [smooth_solve(X, y, labels, w, d, l1w, Ms=Ms)]
    cands.append(conic_solve(X, y, labels, 0.0, w, d, w1=l1w if np.any(l1w > 0) else None, Ms=Ms)[0])
    b = min(cands, key=lambda b: objective(X, y, b, labels, 0.0, w, d, w1=l1w, Ms=Ms))
    b = np.where(np.abs(b) < 1e-11 * np.abs(b).max(), 0.0, b)     # SLSQP leaves 1e-17 instead of 0
    return b, kkt_violation(X, y, b, labels, l1w, w, d, Ms=Ms) / (np.abs(X.T @ y).max() / len(y))


KKT_TOL = 1e-8


def main():
    cases, dropped, problems = [], [], {}
    for seed, (n, p, G, noise) in enumerate([(40, 12, 4, 0.5), (30, 15, 5, 1.0)]):
        X, y, labels = problem(100 + seed, n, p, G, noise)
        _, inv = np.unique(labels, return_inverse=True)
        amax = np.abs(X.T @ y).max() / n
        gw = 0.5 + np.arange(G) / G
        for frac in (0.3, 0.05):
            alpha = float(frac * amax)
            problems[str(100 + seed)] = dict(n=n, p=p, G=G, noise=noise, X=X.tolist(), y=y.tolist(), groups=labels.tolist())
            base = dict(seed=100 + seed, alpha=alpha)

            def rec(name, b, kkt, extra):
                if kkt <= KKT_TOL:
                    cases.append(dict(base, name=name, coef=b.tolist(), kkt=kkt, **extra))
                else:
                    dropped.append((name, 100 + seed, frac, kkt))

            zero, nol1 = np.zeros(G), np.zeros(p)
            ones = np.ones(G)
            # GroupLasso, default and explicit group weights
            b, k = solve_checked(X, y, inv, nol1, alpha * ones, zero)
            rec("GroupLasso", b, k, dict(objective=objective(X, y, b, inv, 0.0, alpha * ones, zero)))
            b, k = solve_checked(X, y, inv, nol1, alpha * gw, zero)
            rec("GroupLasso", b, k, dict(group_weights=gw.tolist(), objective=objective(X, y, b, inv, 0.0, alpha * gw, zero)))
            # SparseGroupLasso
            l1r = 0.4
            lam1, lam2 = l1r * alpha, (1 - l1r) * alpha
            b, k = solve_checked(X, y, inv, lam1 * np.ones(p), lam2 * ones, zero)
            rec("SparseGroupLasso", b, k, dict(l1_ratio=l1r, objective=objective(X, y, b, inv, lam1, lam2 * ones, zero)))
            # RidgedGroupLasso
            delta = 0.3 + 0.2 * np.arange(G)
            b, k = solve_checked(X, y, inv, nol1, alpha * ones, delta)
            rec("RidgedGroupLasso", b, k, dict(delta=delta.tolist(), objective=objective(X, y, b, inv, 0.0, alpha * ones, delta)))
            # OverlapGroupLasso: every third feature also belongs to the next group
            group_list = [[int(inv[j])] + ([int((inv[j] + 1) % G)] if j % 3 == 0 else []) for j in range(p)]
            ext_cols, ext_lab = [], []
            for g in range(G):                      # _lasso.py:446-449
                for j in range(p):
                    if g in group_list[j]:
                        ext_cols.append(j)
                        ext_lab.append(g)
            ext_cols, ext_lab = np.array(ext_cols), np.array(ext_lab)
            Xe = X[:, ext_cols]
            be, k = solve_checked(Xe, y, ext_lab, np.zeros(len(ext_cols)), alpha * ones, zero)
            b = np.zeros(p)
            np.add.at(b, ext_cols, be)               # _lasso.py:492-501
            # the expanded problem has flat directions (duplicated columns): the folded coefficients and
            # the objective are unique, the split between the copies need not be
            rec("OverlapGroupLasso", b, k, dict(group_list=group_list,
                                                objective=objective(Xe, y, be, ext_lab, 0.0, alpha * ones, zero)))
            # AdaptiveGroupLasso: _adaptive_lasso.py:343-374 (v0 = alpha, v <- alpha*gw * alpha/(||b_g||+eps))
            eps = 1e-6
            v, worst = alpha * ones, 0.0
            for _ in range(3):
                b, k = solve_checked(X, y, inv, nol1, v, zero)
                worst = max(worst, k)
                nr = np.array([np.linalg.norm(b[inv == g]) for g in range(G)])
                v = alpha * gw * (alpha / (nr + eps))
            rec("AdaptiveGroupLasso", b, worst, dict(group_weights=gw.tolist(), max_iter=3, eps=eps))
            # AdaptiveSparseGroupLasso: _adaptive_lasso.py:654-726
            w1, v, worst = lam1 * np.ones(p), lam2 * ones, 0.0
            for _ in range(3):
                b, k = solve_checked(X, y, inv, w1, v, zero)
                worst = max(worst, k)
                nr = np.array([np.linalg.norm(b[inv == g]) for g in range(G)])
                w1 = lam1 * (alpha / (np.abs(b) + eps))
                v = lam2 * gw * (alpha / (nr + eps))
            rec("AdaptiveSparseGroupLasso", b, worst, dict(l1_ratio=l1r, group_weights=gw.tolist(), max_iter=3, eps=eps))
    # ---- preprocessing and standardize=True variants (kept apart: "cases_preprocess") -------------------
    def preprocess(X, y, sw, fit_intercept):
        """_base.py:207-227 restated: weights rescaled to sum n, (weighted) centring, rows scaled by sqrt(sw)."""
        n = len(y)
        if sw is not None:
            sw = sw * (n / sw.sum())
        mu, yb = np.zeros(X.shape[1]), 0.0
        if fit_intercept:
            mu, yb = np.average(X, axis=0, weights=sw), np.average(y, weights=sw)
        Xc, yc = X - mu, y - yb
        if sw is not None:
            Xc, yc = Xc * np.sqrt(sw)[:, None], yc * np.sqrt(sw)
        return Xc, yc, mu, yb

    pre = []
    for seed, (n, p, G, noise) in enumerate([(40, 12, 4, 0.5), (30, 15, 5, 1.0)]):
        X, y, labels = problem(100 + seed, n, p, G, noise)
        X = X + 0.5                     # non-zero column means: the intercept matters
        y = y + 3.0
        rng = np.random.default_rng(200 + seed)
        sw = 0.5 + 2.0 * rng.random(n)
        _, inv = np.unique(labels, return_inverse=True)
        problems[f"pre{100 + seed}"] = dict(n=n, p=p, G=G, noise=noise, X=X.tolist(), y=y.tolist(),
                                            groups=labels.tolist(), sample_weight=sw.tolist())
        ones, zero, nol1 = np.ones(G), np.zeros(G), np.zeros(p)
        delta = 0.3 + 0.2 * np.arange(G)
        for frac in (0.3, 0.05):
            for name, fi, use_sw, std in (("GroupLasso", True, False, False), ("GroupLasso", True, True, False),
                                          ("SparseGroupLasso", False, True, False), ("GroupLasso", True, False, True),
                                          ("GroupLasso", False, True, True), ("RidgedGroupLasso", True, False, True)):
                Xc, yc, mu, yb = preprocess(X, y, sw if use_sw else None, fi)
                alpha = float(frac * np.abs(Xc.T @ yc).max() / n)
                masks = [np.flatnonzero(inv == g) for g in range(G)]
                Ms, d, extra = None, zero, {}
                if std and name == "GroupLasso":            # ||X_g b_g||_2 (_lasso.py:249-252)
                    Ms = [Xc[:, m].T @ Xc[:, m] for m in masks]
                if name == "RidgedGroupLasso":              # ||sqrtm(X_g'X_g + sqrt(delta_g) I) b_g||_2 (:776-789) + ridge
                    Ms = [Xc[:, m].T @ Xc[:, m] + np.sqrt(delta[g]) * np.eye(len(m)) for g, m in enumerate(masks)]
                    d, extra = delta, {"delta": delta.tolist()}
                if name == "SparseGroupLasso":
                    l1w, w = 0.4 * alpha * np.ones(p), 0.6 * alpha * ones
                    extra = {"l1_ratio": 0.4}
                else:
                    l1w, w = nol1, alpha * ones
                if Ms is not None:                          # the standardized norms have the scale of ||X_g||: alpha from them
                    alpha = float(frac * max(np.sqrt(c_ @ np.linalg.solve(M_, c_)) for M_, c_ in
                                             ((Ms[g], (Xc.T @ yc)[m] / n) for g, m in enumerate(masks))))
                    w = alpha * ones
                b, k = solve_checked(Xc, yc, inv, l1w, w, d, Ms=Ms)
                rec_ = dict(problem=f"pre{100 + seed}", name=name, alpha=alpha, fit_intercept=fi, use_sample_weight=use_sw,
                            standardize=std, coef=b.tolist(), intercept=float(yb - mu @ b) if fi else 0.0, kkt=k, **extra)
                (pre if k <= KKT_TOL else dropped).append(rec_ if k <= KKT_TOL else (name, f"pre{100 + seed}", frac, k))
    with open(os.path.join(HERE, "golden_nlp.json"), "w") as fh:
        json.dump({"solver": "scipy.optimize.minimize: SLSQP on the epigraph (SOCP) formulation, restarted until the "
                             "objective stalls, and a smoothed-norm homotopy with 'trust-exact'; the lower objective is kept",
                   "kkt_tol": KKT_TOL, "problems": problems, "cases": cases, "cases_preprocess": pre}, fh)
    print(len(cases), "+", len(pre), "cases kept;", len(dropped), "dropped (third-party solve not converged):", dropped)


if __name__ == "__main__":
    main()

"""Fixtures from the REFERENCE'S OWN estimator code, run in the build container.

    python tests/golden/make_golden_reference.py      ->  tests/golden/golden_reference.json

cvxpy is not installable here, so `problem.solve()` cannot run -- but everything around it can.
This script installs tests/golden/cvxpy_shim.py as `cvxpy`, loads the reference's
`model/_base.py`, `_lasso.py`, `_adaptive_lasso.py` BY PATH from /root/reference (unmodified) and
calls the reference's own `Estimator.fit(X, y, sample_weight)`.  That executes, in the reference's
code: `_preprocess_data` (_base.py:207-227), `_validate_params`, `generate_problem` with
`_generate_params` / `_generate_auxiliaries` / `_generate_objective` (_base.py:414-467; group
norms _lasso.py:239-255, standardized norms :249-252 and :776-789, lambda1/lambda2 :616-639, delta
:755-765, overlap expansion :440-484), the adaptive loop `AdaptiveLasso._solve`
(_adaptive_lasso.py:206-232) with `_iterative_update` (:196-204, :364-374, :712-726) and
`_check_convergence` (:189-194, :698-710), the overlap fold-back (_lasso.py:486-502) and
`_set_intercept`.

The only substituted step is the conic solve.  For every `problem.solve()` the hook
  1. reads the problem the reference built -- `cvxpy_shim.analyze` probes the reference's
     objective expression into  sum c||Ab-r||^2 + sum ||w o b||_1 + sum v||M b_S||_2  with the
     CURRENT parameter values (so the adaptive weights are the reference's, not ours);
  2. minimises that generic form with the oracle's block-coordinate-descent core (a solver, told
     nothing about which estimator this is; non-identity M are whitened here);
  3. certifies the minimiser with a KKT residual computed in this file from the same generic
     form (numpy / scipy.optimize.lsq_linear only -- independent of oracle/), relative to the
     gradient scale, and refuses to write a fixture above 1e-9.
The reference's own expression is also evaluated at random points (`objective_probes`) so that a
test can check "oracle objective == reference objective" as FUNCTIONS.

Environment adaptation (not reference code): scikit-learn 1.9 removed `BaseEstimator._validate_data`
and made `_preprocess_data` keyword-only with a sixth return value and its own sqrt(sw) rescaling
(SURVEY F3).  Two adapters restore the scikit-learn >= 1.2 behaviour the reference was written for:
`_validate_data -> sklearn.utils.validation.validate_data`, `_preprocess_data(...,
rescale_with_sw=False)[:5]` (the reference rescales itself, _base.py:224-225).
"""

from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/sparselm"
for pth in (ROOT, HERE):
    if pth not in sys.path:
        sys.path.insert(0, pth)

import cvxpy_shim as shim  # noqa: E402


# --------------------------------------------------------------------------- #
# load the reference modules with the shim as cvxpy
# --------------------------------------------------------------------------- #
def load_reference():
    sys.modules["cvxpy"] = shim
    for name, sub in (("sparselm", ""), ("sparselm.model", "model"), ("sparselm._utils", "_utils")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REF, sub)]
            sys.modules[name] = pkg

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    load("sparselm._utils.validation", "_utils/validation.py")
    base = load("sparselm.model._base", "model/_base.py")
    lasso = load("sparselm.model._lasso", "model/_lasso.py")
    adaptive = load("sparselm.model._adaptive_lasso", "model/_adaptive_lasso.py")

    # --- scikit-learn 1.9 adapters (see module docstring) ---
    from sklearn.linear_model._base import _preprocess_data as _sk_preprocess
    from sklearn.utils.validation import validate_data

    def _preprocess_data(X, y, copy=True, fit_intercept=True, sample_weight=None):
        return _sk_preprocess(X, y, fit_intercept=fit_intercept, copy=copy, sample_weight=sample_weight,
                              rescale_with_sw=False)[:5]

    base._preprocess_data = _preprocess_data
    base.CVXRegressor._validate_data = lambda self, X, y, **kw: validate_data(self, X, y, **kw)
    return base, lasso, adaptive


# --------------------------------------------------------------------------- #
# generic form -> minimiser (oracle BCD core) and an independent KKT certificate
# --------------------------------------------------------------------------- #
def _sym_sqrt_and_inv(A):
    w, V = np.linalg.eigh((A + A.T) / 2)
    if w.min() <= 1e-12 * w.max():
        raise ValueError("group norm matrix is singular")
    return (V * np.sqrt(w)) @ V.T, (V / np.sqrt(w)) @ V.T


def _generic_structure(sf, p):
    """Parse the probed standard form: stacked least-squares rows, l1 weights per coordinate,
    disjoint l2 groups (support S, matrix M on the support, weight v)."""
    rows_A, rows_r = [], []
    for c, A, r in sf["quad"]:
        if c < 0:
            raise ValueError("concave quadratic")
        rows_A.append(np.sqrt(c) * A)
        rows_r.append(np.sqrt(c) * r)
    A_all, r_all = np.vstack(rows_A), np.concatenate(rows_r)
    w1 = np.zeros(p)
    for w, A, r in sf["l1"]:
        if np.any(r != 0):
            raise ValueError("shifted l1 term")
        for i in range(A.shape[0]):
            nz = np.flatnonzero(A[i])
            if len(nz) != 1 or A[i, nz[0]] != 1.0:
                raise ValueError("l1 term is not on plain coordinates")
            w1[nz[0]] += w[i]
    groups = []
    used = np.zeros(p, dtype=bool)
    for v, A, r in sf["l2"]:
        if np.any(r != 0):
            raise ValueError("shifted l2 term")
        S = np.flatnonzero(np.any(A != 0, axis=0))
        if used[S].any():
            raise ValueError("l2 groups overlap in the solver variables")
        used[S] = True
        groups.append((S, A[:, S], float(v)))
    return A_all, r_all, w1, groups, used


def solve_generic(sf, p):
    import oracle.reference as R

    A_all, r_all, w1, groups, used = _generic_structure(sf, p)
    n_rows = A_all.shape[0]
    labels = np.full(p, -1, dtype=np.int64)
    w2 = []
    Xs = A_all.copy()
    backs = []
    for gi, (S, M, v) in enumerate(groups):
        labels[S] = gi
        w2.append(v)
        ident = M.shape[0] == len(S) and np.array_equal(M, np.eye(len(S)))
        if not ident:
            if np.any(w1[S] != 0):
                return solve_generic_split(A_all, r_all, w1, groups, used, p)
            Rg, Rinv = _sym_sqrt_and_inv(M.T @ M)  # ||M b|| = ||R b||, gamma = R b
            Xs[:, S] = A_all[:, S] @ Rinv
            backs.append((S, Rinv))
    nxt = len(groups)
    for j in np.flatnonzero(~used):
        labels[j] = nxt
        w2.append(0.0)
        nxt += 1
    sc = 1.0 / (2.0 * n_rows)  # oracle data term is 1/(2 n_rows) ||.||^2
    pen = R.Penalty(labels, w1 * sc, np.asarray(w2) * sc, np.zeros(nxt))
    # tol < 0: never stop on the duality gap (a gap of eps still leaves sqrt(eps) in beta); iterate
    # until a further sweep cannot move the iterate (status 2)
    gamma, info = R.solve(Xs, r_all, pen, tol=-1.0, max_sweeps=5000000, check_every=1000)
    beta = gamma.copy()
    for S, Rinv in backs:
        beta[S] = Rinv @ gamma[S]
    return beta, info


def solve_generic_split(A_all, r_all, w1, groups, used, p, rho=10.0, tol=1e-13, max_outer=5000):
    """Generic form with an l1 term on coordinates whose group norm has a non-identity matrix
    (||M b_S||, M'M = R^2): not separable in any single set of variables.  Method of multipliers on
    s_S = R b_S / sqrt(m) (m = rows of the least-squares part): the inner problem in (b, s) has
    separable penalties and goes to the same BCD core; the multiplier moves by the constraint
    residual until it vanishes.  The result is certified by kkt_residual_generic like every other."""
    import oracle.reference as R

    m = A_all.shape[0]
    Rbd = np.zeros((p, p))
    for S, M, v in groups:
        Rg, _ = _sym_sqrt_and_inv(M.T @ M)
        Rbd[np.ix_(S, S)] = Rg
    in_group = used.copy()
    Rh = Rbd / np.sqrt(m)
    sr = np.sqrt(m * rho)
    # columns of s only for grouped coordinates (ungrouped coordinates carry no l2 term)
    gs = np.flatnonzero(in_group)
    q = len(gs)
    Xaug = np.block([[A_all, np.zeros((m, q))], [sr * Rh[np.ix_(gs, np.arange(p))], -sr * np.eye(q)]])
    nG = len(groups)
    lab_b = nG + np.arange(p)
    lab_s = np.empty(q, dtype=np.int64)
    pos = {int(j): i for i, j in enumerate(gs)}
    w2 = np.zeros(nG + p)
    for gi, (S, M, v) in enumerate(groups):
        lab_s[[pos[int(j)] for j in S]] = gi
        w2[gi] = np.sqrt(m) * v
    rows = m + q
    sc = 1.0 / (2.0 * rows)  # the form is ||A b - r||^2 + pen: oracle data term 1/(2 rows)||.||^2
    pen = R.Penalty(np.concatenate([lab_b, lab_s]).astype(np.int64), np.concatenate([w1, np.zeros(q)]) * sc,
                    w2 * sc * 1.0, np.zeros(nG + p))
    # note: the split adds (rho'/2)-type terms in the same ||.||^2 scaling, so the penalties keep the form's scale
    u = np.zeros(q)
    z = None
    info = {"sweeps": 0, "status": 1}
    for it in range(max_outer):
        yaug = np.concatenate([r_all, -sr * u])
        z, inf = R.solve(Xaug, yaug, pen, tol=-1.0, max_sweeps=5000000, beta0=z, check_every=1000)
        info["sweeps"] += inf["sweeps"]
        res = Rh[np.ix_(gs, np.arange(p))] @ z[:p] - z[p:]
        u = u + res
        if np.abs(res).max() <= tol * max(1.0, np.abs(z[p:]).max()):
            info["status"] = 0
            break
    beta = z[:p].copy()
    for S, M, v in groups:  # a group whose split variable is exactly zero is a zero group (R b = s in the limit)
        if not np.any(z[p + np.array([pos[int(j)] for j in S])]):
            beta[S] = 0.0
    return beta, info


def kkt_residual_generic(sf, p, beta):
    """max violation of 0 in d f(beta) for the probed generic form, relative to the gradient
    scale.  Independent of oracle/: numpy + scipy.optimize.lsq_linear only."""
    from scipy.optimize import lsq_linear

    A_all, r_all, w1, groups, used = _generic_structure(sf, p)
    g = -2.0 * A_all.T @ (A_all @ beta - r_all)  # minus the gradient of the smooth part
    scale = max(float(np.abs(g).max()), float(np.abs(2.0 * A_all.T @ r_all).max()), 1e-300)
    worst = 0.0

    def l1_part(t, b, w):  # distance of t from w o d|b|
        return np.where(b != 0, t - w * np.sign(b), np.sign(t) * np.maximum(np.abs(t) - w, 0.0))

    for S, M, v in groups:
        b, gs, w = beta[S], g[S], w1[S]
        Mb = M @ b
        nrm = float(np.linalg.norm(Mb))
        if nrm > 0:
            t = gs - v * (M.T @ Mb) / nrm
            worst = max(worst, float(np.abs(l1_part(t, b, w)).max()))
        else:
            # need s in [-w, w], ||u|| <= v with gs = s + M^T u.  With M^T M = R^2 (R symmetric,
            # invertible): min_s ||R^{-1}(gs - s)|| <= v
            Rg, Rinv = _sym_sqrt_and_inv(M.T @ M)
            if np.all(w == 0):
                dist = float(np.linalg.norm(Rinv @ gs))
            else:
                res = lsq_linear(Rinv, Rinv @ gs, bounds=(-w - 1e-300, w + 1e-300), method="bvls", tol=1e-14)
                dist = float(np.linalg.norm(Rinv @ (gs - res.x)))
            worst = max(worst, max(0.0, dist - v) * float(np.linalg.norm(Rg, 2)))
    rest = np.flatnonzero(~used)
    if len(rest):
        worst = max(worst, float(np.abs(l1_part(g[rest], beta[rest], w1[rest])).max()))
    return worst / scale


PASS_LOG = []


def solve_hook(problem, **kwargs):
    var = problem.variables()[0]
    p = var.shape[0]
    sf = shim.analyze(problem.objective.expr, var)
    beta, info = solve_generic(sf, p)
    kkt = kkt_residual_generic(sf, p, beta)
    var.value = beta
    obj = float(problem.objective.value)
    PASS_LOG.append({"objective": obj, "kkt_rel": kkt, "sweeps": info["sweeps"], "oracle_status": info["status"]})
    return obj


# --------------------------------------------------------------------------- #
# cases
# --------------------------------------------------------------------------- #
def make_data(seed, n, p, n_inf=5, noise=0.3):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    w = np.zeros(p)
    w[rng.choice(p, n_inf, replace=False)] = rng.standard_normal(n_inf) * 2.0
    y = X @ w + noise * rng.standard_normal(n) + 0.7
    sw = rng.random(n) + 0.25
    return X, y, sw


GROUPS16 = [7, 7, 3, 3, 3, 0, 0, 11, 11, 11, 11, 5, 2, 2, 7, 3]          # unsorted, non-contiguous labels
GW16 = [1.0, 0.5, 2.0, 1.5, 0.8, 1.2]                                     # one per sorted unique label
GROUP_LIST16 = [[0], [0, 1], [1], [1, 2], [2], [2], [3], [3, 0], [3], [4], [4, 5], [5], [5], [], [4], [2, 5]]
GW_OVERLAP = [1.0, 0.7, 1.3, 0.9, 1.1, 1.6]
DELTA6 = [0.5, 0.0, 1.0, 2.0, 0.3, 0.8]


def case_list():
    cases = []

    def add(name, kwargs, seed=0, n=40, p=16, fit_intercept=False, weighted=False):
        cases.append(dict(estimator=name, kwargs=kwargs, seed=seed, n=n, p=p, fit_intercept=fit_intercept,
                          weighted=weighted))

    for a in (0.05, 0.4):
        add("Lasso", dict(alpha=a))
        add("Lasso", dict(alpha=a), seed=1, fit_intercept=True, weighted=True)
        add("GroupLasso", dict(groups=GROUPS16, alpha=a))
        add("GroupLasso", dict(groups=GROUPS16, alpha=a, group_weights=GW16), seed=2, fit_intercept=True)
        add("GroupLasso", dict(groups=GROUPS16, alpha=a, standardize=True), seed=3, weighted=True)
        add("GroupLasso", dict(groups=GROUPS16, alpha=a, standardize=True, group_weights=GW16), seed=4,
            fit_intercept=True, weighted=True)
        add("OverlapGroupLasso", dict(group_list=GROUP_LIST16, alpha=a))
        add("OverlapGroupLasso", dict(group_list=GROUP_LIST16, alpha=a, group_weights=GW_OVERLAP, standardize=True),
            seed=5, fit_intercept=True)
        add("SparseGroupLasso", dict(groups=GROUPS16, alpha=a, l1_ratio=0.5))
        add("SparseGroupLasso", dict(groups=GROUPS16, alpha=a, l1_ratio=0.2, group_weights=GW16), seed=6,
            fit_intercept=True, weighted=True)
        add("RidgedGroupLasso", dict(groups=GROUPS16, alpha=a, delta=(1.0,)))
        add("RidgedGroupLasso", dict(groups=GROUPS16, alpha=a, delta=DELTA6, group_weights=GW16), seed=7,
            fit_intercept=True)
        add("RidgedGroupLasso", dict(groups=GROUPS16, alpha=a, delta=DELTA6, standardize=True), seed=8, weighted=True)
        add("AdaptiveLasso", dict(alpha=a))
        add("AdaptiveLasso", dict(alpha=a, max_iter=5, eps=1e-4), seed=9, fit_intercept=True, weighted=True)
        add("AdaptiveGroupLasso", dict(groups=GROUPS16, alpha=a, group_weights=GW16))
        add("AdaptiveGroupLasso", dict(groups=GROUPS16, alpha=a, standardize=True), seed=10, fit_intercept=True)
        add("AdaptiveOverlapGroupLasso", dict(group_list=GROUP_LIST16, alpha=a, group_weights=GW_OVERLAP))
        add("AdaptiveOverlapGroupLasso", dict(group_list=GROUP_LIST16, alpha=a, standardize=True), seed=11,
            weighted=True)
        add("AdaptiveSparseGroupLasso", dict(groups=GROUPS16, alpha=a, l1_ratio=0.5, group_weights=GW16))
        add("AdaptiveSparseGroupLasso", dict(groups=GROUPS16, alpha=a, l1_ratio=0.7, max_iter=4), seed=12,
            fit_intercept=True, weighted=True)
        add("AdaptiveRidgedGroupLasso", dict(groups=GROUPS16, alpha=a, delta=(0.5,), group_weights=GW16))
        add("AdaptiveRidgedGroupLasso", dict(groups=GROUPS16, alpha=a, delta=DELTA6, standardize=True), seed=13,
            fit_intercept=True, weighted=True)
    # SparseGroupLasso with standardize=True: l1 on b, group norms ||X_g b_g|| (_lasso.py:249-252, :627-639)
    for a in (0.05, 0.4):
        add("SparseGroupLasso", dict(groups=GROUPS16, alpha=a, l1_ratio=0.5, standardize=True), seed=19)
        add("SparseGroupLasso", dict(groups=GROUPS16, alpha=a, l1_ratio=0.3, standardize=True, group_weights=GW16),
            seed=20, fit_intercept=True, weighted=True)
        add("AdaptiveSparseGroupLasso", dict(groups=GROUPS16, alpha=min(a, 0.12), l1_ratio=0.5 if a < 0.1 else 0.3,
                                             standardize=True), seed=21, fit_intercept=True)
    # groups=None / group_list=None degenerate to singleton groups (with the reference's warning)
    add("GroupLasso", dict(groups=None, alpha=0.1), seed=14)
    add("OverlapGroupLasso", dict(group_list=None, alpha=0.1), seed=15)
    # p > n
    add("SparseGroupLasso", dict(groups=GROUPS16, alpha=0.2, l1_ratio=0.5), seed=16, n=12)
    # adaptive chain that stops early (tol large) and a single pass
    add("AdaptiveLasso", dict(alpha=0.3, tol=1.0, max_iter=6), seed=17)
    add("AdaptiveGroupLasso", dict(groups=GROUPS16, alpha=0.3, max_iter=2), seed=18)
    return cases


def weights_after(est):
    prm = est.canonicals_.parameters
    out = {}
    for nm in ("adaptive_weights", "adaptive_coef_weights", "adaptive_group_weights"):
        if hasattr(prm, nm):
            out[nm] = np.asarray(getattr(prm, nm).value, dtype=float).tolist()
    return out


def main():
    base, lasso, adaptive = load_reference()
    shim.SOLVE_HOOK = solve_hook
    classes = {**{k: getattr(lasso, k) for k in ("Lasso", "GroupLasso", "OverlapGroupLasso", "SparseGroupLasso",
                                                  "RidgedGroupLasso")},
               **{k: getattr(adaptive, k) for k in ("AdaptiveLasso", "AdaptiveGroupLasso", "AdaptiveOverlapGroupLasso",
                                                     "AdaptiveSparseGroupLasso", "AdaptiveRidgedGroupLasso")}}
    out = []
    worst_kkt = 0.0
    for ci, c in enumerate(case_list()):
        X, y, sw = make_data(c["seed"], c["n"], c["p"])
        kw = dict(c["kwargs"])
        for k in ("groups", "group_weights", "delta"):
            if kw.get(k) is not None:
                kw[k] = np.asarray(kw[k], dtype=float if k != "groups" else int)
        est = classes[c["estimator"]](fit_intercept=c["fit_intercept"], **kw)
        PASS_LOG.clear()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)
            est.fit(X, y, sample_weight=sw if c["weighted"] else None)
        passes = [dict(v) for v in PASS_LOG]
        kmax = max(v["kkt_rel"] for v in passes)
        worst_kkt = max(worst_kkt, kmax)
        if kmax > 1e-9 or any(v["oracle_status"] not in (0, 2) for v in passes):  # 2 = stationary to rounding
            raise SystemExit(f"case {ci} {c['estimator']} {c['kwargs']}: KKT {kmax:.2e} -- no fixture written")
        rec = dict(c)
        rec["x_sha"] = hashlib.sha256(np.ascontiguousarray(X).tobytes() + np.ascontiguousarray(y).tobytes()
                                      + np.ascontiguousarray(sw).tobytes()).hexdigest()[:16]
        rec["coef"] = np.asarray(est.coef_, dtype=float).tolist()
        rec["intercept"] = float(est.intercept_)
        rec["n_iter"] = int(getattr(est, "n_iter_", 0)) or None
        rec["passes"] = [{"objective": v["objective"], "kkt_rel": v["kkt_rel"]} for v in passes]
        rec["weights_after"] = weights_after(est)
        # the reference's own objective expression at random points (state: parameters as they are
        # AFTER fit; for the non-adaptive estimators that is the problem that was solved)
        var = est.canonicals_.beta
        keep = var.value
        probes = []
        rng = np.random.default_rng(1000 + ci)
        for t in range(3):
            b = rng.standard_normal(var.shape[0]) * (rng.random(var.shape[0]) < (0.5 if t else 1.0))
            var.value = b
            probes.append({"beta": b.tolist(), "objective": float(est.canonicals_.objective.value)})
        var._value = keep
        rec["objective_probes"] = probes
        aux = est.canonicals_.auxiliaries
        if aux is not None and hasattr(aux, "extended_coef_indices"):
            rec["extended_coef_indices"] = np.asarray(aux.extended_coef_indices).astype(int).tolist()
        out.append(rec)
        print(f"{ci:3d} {c['estimator']:28s} passes {len(passes)} kkt {kmax:.1e} nnz {int(np.sum(np.abs(est.coef_) > 1e-9))}")
    with open(os.path.join(HERE, "golden_reference.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden_reference.py", "worst_kkt_rel": worst_kkt, "cases": out}, f)
    print(f"wrote golden_reference.json: {len(out)} cases, worst KKT residual {worst_kkt:.2e}")


if __name__ == "__main__":
    main()

"""oracle/reference.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the hot path of CederGroupHub/sparse-lm (the convex
estimators' ``fit``) used as the parity checker for the CUDA engine.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package
(``sparselm_b200``) never does.

What is restated, with the reference lines it follows (paths relative to
/root/reference/src/sparselm):

* preprocessing                         model/_base.py:207-227
* Lasso objective                       model/_lasso.py:99-121
* GroupLasso penalty / default weights  model/_lasso.py:230-275
* OverlapGroupLasso expansion/fold-back model/_lasso.py:440-461, 486-502
* SparseGroupLasso lambda1/lambda2      model/_lasso.py:616-639
* RidgedGroupLasso delta handling       model/_lasso.py:755-765, 795-811
* adaptive reweighting loop             model/_adaptive_lasso.py:206-232
* adaptive updates (alpha^2 quirk)      model/_adaptive_lasso.py:177-204, 343-374, 654-726
* standardize=True group norms          model/_lasso.py:249-252, 776-789
  (SparseGroupLasso + standardize: l1 on b with ||X_g b_g|| group norms, method of multipliers)

The arithmetic of the reference lives in cvxpy (>=1.2, unpinned,
pyproject.toml:15) and whichever conic solver it selects; neither is under
/root/reference nor installable in this image.  The solve itself is therefore a
restatement of the *problem* (convex, so any exact solver agrees) by a cyclic
block-coordinate descent in C (slm_oracle.c) that is certified by a duality gap
computed from X and y.

PARITY PINNING: pinned by the reference's OWN problem-building code.  tests/golden/make_golden_reference.py
installs a numeric stand-in for the cvxpy names the reference uses (tests/golden/cvxpy_shim.py), loads
/root/reference/src/sparselm/model/_lasso.py, _adaptive_lasso.py and _base.py by path and calls the reference's
_generate_params / _generate_auxiliaries / _generate_objective (_base.py:453-458) and _iterative_update
(_adaptive_lasso.py:196-204, 364-374, 712-726).  The committed fixtures (tests/golden/golden_reference.json, 57
cases over all ten estimators) hold the reference's objective values, its adaptive weight updates and a minimiser
of the reference's objective with an independent KKT certificate; tests/test_reference_fixtures.py checks this
module against them (objective to 1e-13, minimiser, weights).  In addition: the reference's only numeric
known-answer test (tests/test_lasso.py:29-61), sklearn's coordinate-descent Lasso on random problems, third-party
conic solutions (tests/golden/make_golden_nlp.py), closed forms on orthonormal designs and the reference's
structural tests.  What cannot run here is the reference's solver (cvxpy + a conic backend): the solve is a
restatement of the convex problem, certified by a duality gap computed from X and y.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libslm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile slm_oracle.c with gcc (OpenMP). Returns the .so path."""
    src = os.path.join(_HERE, "slm_oracle.c")
    if (
        not force
        and os.path.exists(_LIB_PATH)
        and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)
    ):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    cmd = ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", _LIB_PATH, src, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        lib.slmo_bcd.restype = ctypes.c_int
        lib.slmo_bcd.argtypes = [
            ctypes.c_int64, ctypes.c_int64, dp, dp, ctypes.c_int64, ip, dp, dp, dp,
            ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_int64, dp, dp,
        ]
        lib.slmo_certificate.restype = None
        lib.slmo_certificate.argtypes = [
            ctypes.c_int64, ctypes.c_int64, dp, dp, ctypes.c_int64, ip, dp, dp, dp, dp, dp,
        ]
        lib.slmo_bcd_many.restype = ctypes.c_int
        lib.slmo_bcd_many.argtypes = [
            ctypes.c_int64, ip, ctypes.c_int64, ctypes.POINTER(dp), ctypes.POINTER(dp),
            ctypes.c_int64, ip, dp, dp, dp, ctypes.c_double, ctypes.c_double,
            ctypes.c_int64, ctypes.c_int64, dp, dp,
        ]
        lib.slmo_gram_path.restype = ctypes.c_int
        lib.slmo_gram_path.argtypes = [
            ctypes.c_int64, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_int64, ip, ctypes.c_int64,
            dp, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_int64, dp, dp, dp,
        ]
        lib.slmo_gram_path_many.restype = ctypes.c_int
        lib.slmo_gram_path_many.argtypes = [
            ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(dp), ctypes.POINTER(dp), dp, dp, ctypes.c_int64, ip, ip,
            dp, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_int64, dp, dp,
        ]
        lib.slmo_num_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


# --------------------------------------------------------------------------- #
# preprocessing   (model/_base.py:207-227 + sklearn _preprocess_data/_rescale_data)
# --------------------------------------------------------------------------- #
def preprocess(X, y, sample_weight=None, fit_intercept=False):
    X = np.array(X, dtype=np.float64)
    y = np.array(y, dtype=np.float64)
    n = X.shape[0]
    sw = None
    if sample_weight is not None:
        sw = np.asarray(sample_weight, dtype=np.float64)
        if sw.ndim == 0:
            sw = np.full(n, float(sw))
        sw = sw * (n / np.sum(sw))  # _base.py:214
    if fit_intercept:
        X_offset = np.average(X, axis=0, weights=sw)
        y_offset = np.average(y, axis=0, weights=sw)
        X = X - X_offset
        y = y - y_offset
    else:
        X_offset = np.zeros(X.shape[1])
        y_offset = 0.0
    if sw is not None:  # _base.py:224-225 (rows scaled by sqrt(sw))
        s = np.sqrt(sw)
        X = X * s[:, None]
        y = y * s
    return X, y, X_offset, y_offset


# --------------------------------------------------------------------------- #
# penalty description
# --------------------------------------------------------------------------- #
@dataclass
class Penalty:
    """sum_j w1_j|b_j| + sum_g w2_g||b_g|| + 1/2 sum_g delta_g||b_g||^2 over
    groups given as integer labels per feature (sorted-unique order,
    _lasso.py:248)."""

    labels: np.ndarray  # (p,) group index 0..G-1 per feature
    w1: np.ndarray  # (p,)
    w2: np.ndarray  # (G,)
    delta: np.ndarray  # (G,)

    @property
    def n_groups(self):
        return len(self.w2)


def group_labels(groups, p):
    """Map arbitrary labels to 0..G-1 in np.sort(np.unique()) order (_lasso.py:248);
    groups=None -> every feature its own group (_lasso.py:261)."""
    if groups is None:
        return np.arange(p, dtype=np.int64), p
    g = np.asarray(groups)
    uniq, inv = np.unique(g, return_inverse=True)
    return inv.astype(np.int64), len(uniq)


def expand_overlap(group_list, p):
    """Column-duplication expansion of OverlapGroupLasso (_lasso.py:440-461).
    Returns (beta_indices (p_ext,), extended_groups (p_ext,), n_groups)."""
    if group_list is None:
        group_list = [[i] for i in range(p)]
    group_ids = np.sort(np.unique([gid for grp in group_list for gid in grp]))
    beta_inds_list = [
        [i for i, grp in enumerate(group_list) if gid in grp] for gid in group_ids
    ]
    extended = np.concatenate([len(g) * [i] for i, g in enumerate(beta_inds_list)])
    beta_indices = np.concatenate(beta_inds_list)
    return beta_indices.astype(np.int64), extended.astype(np.int64), len(group_ids)


def fold_back(beta_ext, beta_indices, p):
    """_lasso.py:492-501: coef[j] = sum of duplicated coefficients."""
    out = np.zeros(p)
    np.add.at(out, beta_indices, beta_ext)
    return out


# --------------------------------------------------------------------------- #
# solve + certificates
# --------------------------------------------------------------------------- #
def _contiguous(pen: Penalty):
    order = np.argsort(pen.labels, kind="stable")
    counts = np.bincount(pen.labels, minlength=pen.n_groups)
    gptr = np.zeros(pen.n_groups + 1, dtype=np.int64)
    np.cumsum(counts, out=gptr[1:])
    return order, gptr


def solve(X, y, pen: Penalty, tol=1e-13, max_sweeps=200000, beta0=None, check_every=5,
          floor_rel_y=1e-16):
    """argmin of the penalised LS problem by BCD (C). Returns (beta, info dict)."""
    lib = _load()
    n, p = X.shape
    order, gptr = _contiguous(pen)
    Xc = np.asfortranarray(X[:, order], dtype=np.float64)
    yc = np.ascontiguousarray(y, dtype=np.float64)
    w1 = np.ascontiguousarray(pen.w1[order], dtype=np.float64)
    w2 = np.ascontiguousarray(pen.w2, dtype=np.float64)
    dl = np.ascontiguousarray(pen.delta, dtype=np.float64)
    beta = np.zeros(p) if beta0 is None else np.ascontiguousarray(beta0[order], dtype=np.float64)
    info = np.zeros(4)
    floor_abs = floor_rel_y * float(yc @ yc) / (2.0 * n)
    lib.slmo_bcd(n, p, _dp(Xc), _dp(yc), pen.n_groups, _ip(gptr), _dp(w1), _dp(w2), _dp(dl),
                 tol, floor_abs, max_sweeps, check_every, _dp(beta), _dp(info))
    out = np.empty(p)
    out[order] = beta
    return out, {"sweeps": int(info[0]), "primal": info[1], "gap": info[2], "status": int(info[3])}


def penalty_value(beta, pen: Penalty):
    nrm2 = np.bincount(pen.labels, weights=beta * beta, minlength=pen.n_groups)
    return float(pen.w1 @ np.abs(beta) + pen.w2 @ np.sqrt(nrm2) + 0.5 * pen.delta @ nrm2)


def objective(X, y, beta, pen: Penalty):
    """Reference objective value: 1/(2n)||Xb-y||^2 + reg (_lasso.py:120)."""
    r = X @ beta - y
    return float(r @ r) / (2.0 * X.shape[0]) + penalty_value(beta, pen)


def certificate(X, y, beta, pen: Penalty):
    """(primal, dual, gap, omega*) computed in C from X and y."""
    lib = _load()
    n, p = X.shape
    order, gptr = _contiguous(pen)
    Xc = np.asfortranarray(X[:, order], dtype=np.float64)
    yc = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(4)
    lib.slmo_certificate(n, p, _dp(Xc), _dp(yc), pen.n_groups, _ip(gptr),
                         _dp(np.ascontiguousarray(pen.w1[order])), _dp(np.ascontiguousarray(pen.w2)),
                         _dp(np.ascontiguousarray(pen.delta)), _dp(np.ascontiguousarray(beta[order])), _dp(out))
    return {"primal": out[0], "dual": out[1], "gap": out[2], "omega": out[3]}


def kkt_residual(X, y, beta, pen: Penalty):
    """Solver-independent optimality residual in pure numpy (no C): the distance of
    -grad f(beta) from the subdifferential of the penalty, max over groups, in
    gradient units.  Zero iff beta is optimal."""
    n = X.shape[0]
    g = X.T @ (y - X @ beta) / n - pen.delta[pen.labels] * beta  # -grad of smooth part
    worst = 0.0
    for gi in range(pen.n_groups):
        idx = np.flatnonzero(pen.labels == gi)
        bg, gg, w1g, w2g = beta[idx], g[idx], pen.w1[idx], pen.w2[gi]
        nb = np.linalg.norm(bg)
        if nb > 0:
            # need g = w1*s + w2*b/||b||, s_j = sign(b_j) where b_j != 0 else |s_j|<=1
            t = gg - w2g * bg / nb
            res = np.where(bg != 0, t - w1g * np.sign(bg), np.sign(t) * np.maximum(np.abs(t) - w1g, 0))
            worst = max(worst, float(np.max(np.abs(res))))
        else:
            u = np.sign(gg) * np.maximum(np.abs(gg) - w1g, 0)
            worst = max(worst, max(0.0, float(np.linalg.norm(u)) - w2g))
    return worst


# --------------------------------------------------------------------------- #
# standardize=True transforms (_lasso.py:249-252, 776-789)
# --------------------------------------------------------------------------- #
def _sym_sqrt(A):
    w, V = np.linalg.eigh((A + A.T) / 2)
    w = np.maximum(w, 0)
    return (V * np.sqrt(w)) @ V.T, w, V


def standardize_transform(X, labels, n_groups, ridge_sqrt_delta=None):
    """Per-group change of variables gamma_g = R_g beta_g with R_g symmetric PSD,
    R_g^2 = X_g^T X_g (+ sqrt(delta_g) I for the ridged variant, _lasso.py:779-786),
    so that the penalty ||R_g beta_g|| becomes ||gamma_g||.  Returns (X_tilde, Rinv
    list) with X_tilde_g = X_g R_g^+ ; beta_g = R_g^+ gamma_g."""
    Xt = np.zeros_like(X)
    rinv = []
    for gi in range(n_groups):
        idx = np.flatnonzero(labels == gi)
        A = X[:, idx].T @ X[:, idx]
        if ridge_sqrt_delta is not None:
            A = A + ridge_sqrt_delta[gi] * np.eye(len(idx))
        _, w, V = _sym_sqrt(A)
        tol = max(w.max(), 0) * len(idx) * np.finfo(float).eps * 16
        inv_sqrt = np.where(w > tol, 1.0 / np.sqrt(np.where(w > tol, w, 1.0)), 0.0)
        Ri = (V * inv_sqrt) @ V.T
        rinv.append((idx, Ri))
        Xt[:, idx] = X[:, idx] @ Ri
    return Xt, rinv



# --------------------------------------------------------------------------- #
# SparseGroupLasso with standardize=True (_lasso.py:249-252 + :627-639): the group norm is
# ||X_g b_g|| while the l1 term stays on b -- not separable in the whitened variables.  Solved by
# the method of multipliers on the split  s_g = R_g b_g / sqrt(n)  (R_g^2 = X_g^T X_g):
#   min  1/(2n)||y - Xb||^2 + sum w1|b| + sum sqrt(n) w2_g ||s_g||   s.t.  R b / sqrt(n) = s.
# The inner problem in (b, s) is a penalised least-squares problem with SEPARABLE penalties on the
# augmented design [[X, 0], [sqrt(rho) R, -sqrt(n rho) I]] (targets [y; -sqrt(n rho) u]) -- exactly what
# solve() handles; the scaled multiplier moves by the constraint residual until it vanishes.
# --------------------------------------------------------------------------- #
def group_sqrt_blocks(X, labels, n_groups):
    """blockdiag of the symmetric square roots R_g of X_g^T X_g as one [p, p] matrix."""
    p = X.shape[1]
    Rbd = np.zeros((p, p))
    for gi in range(n_groups):
        idx = np.flatnonzero(labels == gi)
        Rg, _, _ = _sym_sqrt(X[:, idx].T @ X[:, idx])
        Rbd[np.ix_(idx, idx)] = Rg
    return Rbd


def objective_sgl_standardized(X, y, beta, labels, w1, w2):
    """1/(2n)||Xb - y||^2 + sum w1|b| + sum_g w2_g ||X_g b_g||."""
    r = X @ beta - y
    val = float(r @ r) / (2.0 * X.shape[0]) + float(np.asarray(w1) @ np.abs(beta))
    for gi in range(len(w2)):
        idx = np.flatnonzero(labels == gi)
        val += float(w2[gi]) * float(np.linalg.norm(X[:, idx] @ beta[idx]))
    return val


def solve_sgl_standardized(X, y, labels, w1, w2, rho=10.0, tol=1e-12, max_outer=2000, solver_tol=-1.0,
                           max_sweeps=200000, start=None):
    """argmin of the standardized sparse-group problem by the method of multipliers.
    Returns (beta, info) with info = dict(outer, residual, u, z): the state to warm-start from."""
    n, p = X.shape
    G = len(w2)
    Rh = group_sqrt_blocks(X, labels, G) / np.sqrt(n)
    sr = np.sqrt(n * rho)
    Xaug = np.block([[X, np.zeros((n, p))], [sr * Rh, -sr * np.eye(p)]])
    lab2 = np.concatenate([G + np.arange(p), labels]).astype(np.int64)  # b: singleton groups G..G+p-1; s: groups 0..G-1
    ps = n / (n + p)  # solve() divides the data term by its own row count
    pen = Penalty(lab2, np.concatenate([w1, np.zeros(p)]) * ps,
                  np.concatenate([np.sqrt(n) * np.asarray(w2, dtype=float), np.zeros(p)]) * ps, np.zeros(G + p))
    u = np.zeros(p) if start is None else start["u"].copy()
    z = None if start is None else start["z"].copy()
    res, it = np.inf, 0
    for it in range(1, max_outer + 1):
        yaug = np.concatenate([y, -sr * u])
        z, _ = solve(Xaug, yaug, pen, tol=solver_tol, max_sweeps=max_sweeps, beta0=z, check_every=50)
        r = Rh @ z[:p] - z[p:]
        u = u + r
        res = float(np.abs(r).max())
        if res <= tol * max(1.0, float(np.abs(z[p:]).max())):
            break
    beta = z[:p].copy()
    for gi in range(G):  # a group whose split variable is exactly zero is a zero group (R b = s in the limit)
        idx = np.flatnonzero(labels == gi)
        if not np.any(z[p + idx]):
            beta[idx] = 0.0
    return beta, {"outer": it, "residual": res, "u": u, "z": z}

# --------------------------------------------------------------------------- #
# estimators (fit on already validated inputs)
# --------------------------------------------------------------------------- #
ADAPTIVE = {
    "AdaptiveLasso", "AdaptiveGroupLasso", "AdaptiveOverlapGroupLasso",
    "AdaptiveSparseGroupLasso", "AdaptiveRidgedGroupLasso",
}
ESTIMATORS = {
    "Lasso", "GroupLasso", "OverlapGroupLasso", "SparseGroupLasso", "RidgedGroupLasso",
} | ADAPTIVE


def _default_update(alpha):
    # _adaptive_lasso.py:181 -- NOTE alpha is already inside the update
    return lambda beta, eps: alpha / (np.abs(beta) + eps)


def fit(name, X, y, *, alpha=1.0, groups=None, group_list=None, group_weights=None,
        l1_ratio=0.5, delta=(1.0,), standardize=False, fit_intercept=False,
        sample_weight=None, max_iter=3, eps=1e-6, tol=1e-10, update_function=None,
        solver_tol=1e-13, max_sweeps=200000, return_details=False):
    """Fit one reference estimator. Returns (coef, intercept[, details])."""
    assert name in ESTIMATORS, name
    Xp, yp, X_off, y_off = preprocess(X, y, sample_weight, fit_intercept)
    n, p = Xp.shape
    adaptive = name in ADAPTIVE
    base = name.replace("Adaptive", "")

    beta_indices = None
    if base == "OverlapGroupLasso":
        beta_indices, labels, G = expand_overlap(group_list, p)
        Xs = Xp[:, beta_indices]
    else:
        if base == "Lasso":
            labels, G = np.arange(p, dtype=np.int64), p
        else:
            labels, G = group_labels(groups, p)
        Xs = Xp
    pe = Xs.shape[1]
    gw = np.ones(G) if group_weights is None else np.asarray(group_weights, dtype=float)

    dl = np.zeros(G)
    if base == "RidgedGroupLasso":
        d = np.asarray(delta, dtype=float)
        dl = d * np.ones(G) if len(d) != G else d  # _lasso.py:762-764

    # standardize (group estimators only)
    rinv = None
    Xsolve = Xs
    if standardize and base != "Lasso":
        rs = np.sqrt(dl) if base == "RidgedGroupLasso" else None  # delta**0.5, _lasso.py:783
        Xsolve, rinv = standardize_transform(Xs, labels, G, rs)

    def to_beta(gamma):
        if rinv is None:
            return gamma
        b = np.zeros_like(gamma)
        for idx, Ri in rinv:
            b[idx] = Ri @ gamma[idx]
        return b

    def ridge_in_gamma():
        return rinv is not None and base == "RidgedGroupLasso" and np.any(dl > 0)

    # initial weights
    w1 = np.zeros(pe)
    w2 = np.zeros(G)
    lam1 = lam2 = None
    if base == "Lasso":
        w1[:] = alpha  # _lasso.py:107 / _adaptive_lasso.py:162-164
    elif base in ("GroupLasso", "OverlapGroupLasso", "RidgedGroupLasso"):
        # adaptive variants start from alpha*ones -- no group_weights in pass 1
        # (_adaptive_lasso.py:347-351,362)
        w2[:] = alpha if adaptive else alpha * gw
    elif base == "SparseGroupLasso":
        lam1, lam2 = l1_ratio * alpha, (1 - l1_ratio) * alpha  # _lasso.py:621-624
        w1[:] = lam1
        w2[:] = lam2 if adaptive else lam2 * gw  # _adaptive_lasso.py:658-667

    # SparseGroupLasso + standardize: l1 on b, group norms ||X_g b_g|| -- solved in b by the method of
    # multipliers (solve_sgl_standardized), not in whitened variables
    sgl_std = rinv is not None and base == "SparseGroupLasso"
    if sgl_std:
        rinv, Xsolve = None, Xs
        Rbd_std = group_sqrt_blocks(Xs, labels, G)

    # ridged + standardize: the ridge 1/2 delta_g ||R_g^+ gamma_g||^2 is smooth; it is the data
    # term of the rows sqrt(n delta_g) R_g^+ appended to the whitened design (targets 0).  The
    # solver's data term is 1/(2 n_rows), so the whole objective is rescaled by n / n_rows.
    ysolve, dl_solve, pen_scale = yp, dl, 1.0
    if ridge_in_gamma():
        blocks = []
        for gi, (idx, Ri) in enumerate(rinv):
            if dl[gi] > 0:
                blk = np.zeros((len(idx), pe))
                blk[:, idx] = np.sqrt(n * dl[gi]) * Ri
                blocks.append(blk)
        Xsolve = np.vstack([Xsolve] + blocks)
        ysolve = np.concatenate([yp, np.zeros(Xsolve.shape[0] - n)])
        dl_solve = np.zeros(G)
        pen_scale = n / Xsolve.shape[0]
    elif rinv is not None:
        dl_solve = np.zeros(G)

    update = update_function if update_function is not None else _default_update(alpha)

    alm_state = {}

    def solve_once(beta0):
        if sgl_std:
            # inner solves run to stationarity (solver_tol < 0): a multiplier iteration cannot push the
            # constraint residual below the accuracy of its inner solutions
            b, st = solve_sgl_standardized(Xs, yp, labels, w1, w2, solver_tol=-1.0, max_sweeps=max_sweeps,
                                           start=alm_state.get("st"))
            alm_state["st"] = st
            return b, {"sweeps": st["outer"], "primal": objective_sgl_standardized(Xs, yp, b, labels, w1, w2),
                       "gap": st["residual"], "status": 0 if st["residual"] <= 1e-9 else 1}
        pen = Penalty(labels, w1 * pen_scale, w2 * pen_scale, dl_solve.copy())
        return solve(Xsolve, ysolve, pen, tol=solver_tol, max_sweeps=max_sweeps, beta0=beta0)

    details = {"passes": []}
    if not adaptive:
        gamma, info = solve_once(None)
        details["passes"].append(info)
        n_iter = None
    else:
        gamma = None
        n_iter = 0
        prev = np.concatenate([w2.copy(), w1.copy()])
        for i in range(max_iter):  # _adaptive_lasso.py:212
            gamma, info = solve_once(gamma)
            details["passes"].append(info)
            n_iter = i + 1
            norms = np.sqrt(np.bincount(labels, weights=gamma * gamma, minlength=G))
            if sgl_std:  # the problem's norm expression is ||X_g b_g|| = ||R_g b_g||
                rb = Rbd_std @ gamma
                norms = np.sqrt(np.bincount(labels, weights=rb * rb, minlength=G))
            # group_norms.value is the *problem's* norm expression: in whitened
            # variables it equals ||X_g b_g|| when standardize (_adaptive_lasso.py:374)
            if base == "Lasso":
                w1 = alpha * np.asarray(update(to_beta(gamma), eps), dtype=float)  # :204
            elif base in ("GroupLasso", "OverlapGroupLasso", "RidgedGroupLasso"):
                w2 = (alpha * gw) * np.asarray(update(norms, eps), dtype=float)  # :372-374
            else:
                w1 = lam1 * np.asarray(update(to_beta(gamma), eps), dtype=float)  # :721-723
                w2 = (lam2 * gw) * np.asarray(update(norms, eps), dtype=float)  # :724-726
            cur = np.concatenate([w2, w1])
            if np.linalg.norm(cur - prev) <= tol:  # :189-194, :698-710
                break
            prev = cur
    beta = to_beta(gamma)
    if beta_indices is not None:
        beta = fold_back(beta, beta_indices, p)
    intercept = float(y_off - X_off @ beta) if fit_intercept else 0.0
    details.update(n_iter=n_iter, w1=w1, w2=w2, labels=labels, delta=dl, sgl_standardized=bool(sgl_std),
                   beta_solve=gamma, X_solve=Xsolve, y_solve=ysolve, pen_scale=pen_scale)
    if return_details:
        return beta, intercept, details
    return beta, intercept


def gram_path(G, c, yty, n, pen_list, labels, tol=1e-13, max_sweeps=200000, check_every=5, floor_rel_y=1e-16):
    """Gram-form BCD along a path (slmo_gram_path): G = X^T X [p, p], c = X^T y, yty, n of one
    training set; pen_list = Penalty objects sharing `labels` (solved in order, each warm-started
    from the previous solution).  Returns (betas [K, p], infos [K, 4])."""
    lib = _load()
    p = len(c)
    n_groups = pen_list[0].n_groups
    order = np.argsort(labels, kind="stable")
    counts = np.bincount(labels, minlength=n_groups)
    gptr = np.zeros(n_groups + 1, dtype=np.int64)
    np.cumsum(counts, out=gptr[1:])
    Gp = np.ascontiguousarray(G[np.ix_(order, order)], dtype=np.float64)
    cp = np.ascontiguousarray(c[order], dtype=np.float64)
    K = len(pen_list)
    w1 = np.ascontiguousarray(np.stack([pn.w1[order] for pn in pen_list]))
    w2 = np.ascontiguousarray(np.stack([pn.w2 for pn in pen_list]))
    dl = np.ascontiguousarray(np.stack([pn.delta for pn in pen_list]))
    betas = np.zeros((K, p))
    infos = np.zeros((K, 4))
    lib.slmo_gram_path(p, _dp(Gp), _dp(cp), float(yty), float(n), n_groups, _ip(gptr), K, _dp(w1), _dp(w2), _dp(dl),
                       tol, floor_rel_y, max_sweeps, check_every, None, _dp(betas), _dp(infos))
    out = np.empty_like(betas)
    out[:, order] = betas
    return out, infos


def num_threads():
    return _load().slmo_num_threads()

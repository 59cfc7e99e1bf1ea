"""Pins the CPU oracle (oracle/) before it is trusted as the parity checker:
reference KAT, sklearn's Lasso, orthonormal closed forms, KKT/gap certificates."""

import json
import os

import numpy as np
import pytest
from sklearn.datasets import make_regression

import oracle.reference as R

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def test_reference_known_answer_lasso_toy():
    kat = GOLD["reference_kat"]
    X, y, T = np.array(kat["X"], float), np.array(kat["y"], float), np.array(kat["T"], float)
    for c in kat["cases"]:
        b, icpt = R.fit("Lasso", X, y, alpha=c["alpha"])
        np.testing.assert_array_almost_equal(b, c["coef"], decimal=kat["decimal"])
        np.testing.assert_array_almost_equal(T @ b + icpt, c["pred"], decimal=kat["decimal"])


@pytest.mark.parametrize("case", GOLD["sklearn_lasso"], ids=lambda c: f"s{c['seed']}-{c['fit_intercept']}-{c['alpha']:.3g}")
def test_oracle_lasso_matches_sklearn_golden(case):
    X, y = make_regression(n_samples=case["n"], n_features=case["p"], n_informative=case["n_informative"],
                           noise=case["noise"], random_state=case["seed"], bias=case["bias"])
    b, icpt = R.fit("Lasso", X, y, alpha=case["alpha"], fit_intercept=case["fit_intercept"])
    ref = np.array(case["coef"])
    assert np.abs(b - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1.0)
    assert abs(icpt - case["intercept"]) <= 1e-9 * max(abs(case["intercept"]), 1.0)


@pytest.mark.parametrize("name", ["Lasso", "GroupLasso", "SparseGroupLasso", "RidgedGroupLasso", "AdaptiveLasso",
                                  "AdaptiveGroupLasso", "AdaptiveSparseGroupLasso", "AdaptiveRidgedGroupLasso"])
def test_oracle_closed_forms_on_orthonormal_design(name):
    o = GOLD["orthonormal"]
    X, y = np.array(o["X"]), np.array(o["y"])
    kw = {}
    if name not in ("Lasso", "AdaptiveLasso"):
        kw.update(groups=np.array(o["groups"]), group_weights=np.array(o["group_weights"]))
    if "Sparse" in name:
        kw["l1_ratio"] = o["l1_ratio"]
    if "Ridged" in name:
        kw["delta"] = np.array(o["delta"])
    b, _ = R.fit(name, X, y, alpha=o["alpha"], **kw)
    np.testing.assert_allclose(b, np.array(o[name]), rtol=0, atol=1e-10)


@pytest.mark.parametrize("name", ["GroupLasso", "SparseGroupLasso", "RidgedGroupLasso", "OverlapGroupLasso"])
def test_oracle_solution_satisfies_kkt_and_gap(name):
    rng = np.random.default_rng(3)
    n, p = 60, 24
    X = rng.standard_normal((n, p))
    y = X[:, :4] @ [2.0, -1.0, 0.5, 1.0] + 0.2 * rng.standard_normal(n)
    groups = rng.integers(0, 6, size=p)
    kw = dict(groups=groups, group_weights=0.5 + rng.random(len(np.unique(groups))))
    if name == "OverlapGroupLasso":
        gl = [list(rng.choice(5, size=rng.integers(1, 3), replace=False)) for _ in range(p)]
        kw = dict(group_list=gl)
    if name == "RidgedGroupLasso":
        kw["delta"] = (0.3,)
    b, _, det = R.fit(name, X, y, alpha=0.1, return_details=True, **kw)
    pen = R.Penalty(det["labels"], det["w1"], det["w2"], det["delta"])
    Xs, ys, bs = det["X_solve"], det["y_solve"], det["beta_solve"]
    assert R.kkt_residual(Xs, ys, bs, pen) <= 1e-9
    cert = R.certificate(Xs, ys, bs, pen)
    assert 0 <= cert["gap"] + 1e-15 and cert["gap"] <= 1e-11 * max(cert["primal"], 1e-300)
    assert cert["primal"] == pytest.approx(R.objective(Xs, ys, bs, pen), rel=1e-12)
    # any perturbation increases the objective (convexity + optimality)
    for _ in range(5):
        d = 1e-4 * rng.standard_normal(len(bs))
        assert R.objective(Xs, ys, bs + d, pen) >= cert["primal"] - 1e-14


def test_oracle_overlap_expansion_and_fold_back():
    group_list = [[0], [0, 1], [1], [2, 0], []]
    idx, ext, ng = R.expand_overlap(group_list, 5)
    np.testing.assert_array_equal(idx, [0, 1, 3, 1, 2, 3])
    np.testing.assert_array_equal(ext, [0, 0, 0, 1, 1, 2])
    assert ng == 3
    np.testing.assert_allclose(R.fold_back(np.arange(1.0, 7.0), idx, 5), [1, 2 + 4, 5, 3 + 6, 0])


def test_oracle_preprocess_matches_weighted_least_squares():
    # reference tests/test_ols.py:36-66 pins sample_weight + intercept handling: a vanishing
    # penalty must reproduce closed-form weighted least squares
    rng = np.random.default_rng(4)
    n, p = 50, 5
    X = rng.standard_normal((n, p))
    y = X @ rng.standard_normal(p) + 0.1 * rng.standard_normal(n) + 2.0
    sw = rng.random(n) + 0.2
    b, icpt = R.fit("Lasso", X, y, alpha=1e-12, fit_intercept=True, sample_weight=sw)
    Xa = np.hstack([X, np.ones((n, 1))])
    W = np.diag(sw)
    sol = np.linalg.solve(Xa.T @ W @ Xa, Xa.T @ W @ y)
    np.testing.assert_allclose(b, sol[:p], atol=1e-8)
    assert icpt == pytest.approx(sol[p], abs=1e-8)


@pytest.mark.parametrize("name", ["GroupLasso", "RidgedGroupLasso"])
@pytest.mark.parametrize("fit_intercept", [False, True])
def test_oracle_standardize_satisfies_kkt_in_original_variables(name, fit_intercept):
    """standardize=True (_lasso.py:249-252, 776-789): the oracle solves in whitened variables;
    its answer is checked here against the optimality conditions of the ORIGINAL problem
    min 1/(2n)||y - X b||^2 + sum_g v_g ||M_g b_g|| + 1/2 sum_g delta_g ||b_g||^2,
    M_g^2 = X_g^T X_g (+ sqrt(delta_g) I), computed in plain numpy."""
    rng = np.random.default_rng(11)
    n, p = 80, 18
    X = rng.standard_normal((n, p)) @ (np.eye(p) + 0.3 * rng.standard_normal((p, p)))
    y = X[:, :4] @ [2.0, -1.0, 0.5, 1.0] + 0.2 * rng.standard_normal(n) + (1.0 if fit_intercept else 0.0)
    groups = rng.integers(0, 5, size=p)
    G = len(np.unique(groups))
    gw = 0.5 + rng.random(G)
    kw = dict(groups=groups, group_weights=gw)
    delta = np.zeros(G)
    if name == "RidgedGroupLasso":
        delta = 0.1 + rng.random(G)
        kw["delta"] = delta
    alpha = 0.01
    b, icpt, det = R.fit(name, X, y, alpha=alpha, standardize=True, fit_intercept=fit_intercept,
                         return_details=True, **kw)
    Xp, yp, _, _ = R.preprocess(X, y, None, fit_intercept)
    labels, _ = R.group_labels(groups, p)
    v = det["w2"]
    r = yp - Xp @ b
    grad = -Xp.T @ r / n
    n_active = 0
    for g in range(G):
        idx = np.flatnonzero(labels == g)
        M2 = Xp[:, idx].T @ Xp[:, idx] + np.sqrt(delta[g]) * np.eye(len(idx))
        bg = b[idx]
        nrm = np.sqrt(bg @ M2 @ bg)
        if nrm > 1e-9:
            n_active += 1
            res = grad[idx] + delta[g] * bg + v[g] * (M2 @ bg) / nrm
            assert np.abs(res).max() <= 1e-8 * max(1.0, np.abs(grad).max())
        else:
            # 0 in grad_g + v_g M_g B(0,1)  <=>  ||M_g^{-1} grad_g|| <= v_g
            L = np.linalg.cholesky(M2)
            assert np.linalg.norm(np.linalg.solve(L, grad[idx])) <= v[g] * (1 + 1e-8)
    assert 0 < n_active < G


# ---- group-type estimators against a third-party solver ---------------------------------------
NLP = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_nlp.json")))


def nlp_case_kwargs(case):
    """(X, y, fit kwargs) of a golden_nlp.json record (shared with the GPU parity test)."""
    pb = NLP["problems"][str(case["seed"])]
    X, y = np.array(pb["X"]), np.array(pb["y"])
    kw = {k: case[k] for k in ("group_weights", "l1_ratio", "group_list", "max_iter", "eps") if k in case}
    if "delta" in case:
        kw["delta"] = tuple(case["delta"])
    if "group_list" not in kw:
        kw["groups"] = np.array(pb["groups"])
    return X, y, kw


@pytest.mark.parametrize("case", NLP["cases"], ids=lambda c: f"{c['name']}-s{c['seed']}-{c['alpha']:.3g}")
def test_oracle_matches_third_party_conic_solution(case):
    """The values the reference's tests leave unpinned (group / sparse-group / ridged / overlap /
    adaptive coefficients): scipy's SLSQP and trust-region Newton codes on the cone program cvxpy
    would build (tests/golden/make_golden_nlp.py), kept only where they reach a KKT violation <= 1e-8.
    Tolerances are north_star's: coefficients 1e-6 * ||b||_inf, support above 1e-6, objective 1e-8."""
    X, y, kw = nlp_case_kwargs(case)
    b, icpt = R.fit(case["name"], X, y, alpha=case["alpha"], fit_intercept=False, **kw)
    ref = np.array(case["coef"])
    assert icpt == 0.0
    assert np.abs(b - ref).max() <= 1e-6 * np.abs(ref).max()
    assert np.array_equal(np.abs(b) > 1e-6, np.abs(ref) > 1e-6)
    if "objective" in case and case["name"] != "OverlapGroupLasso":
        n = len(y)
        G = len(np.unique(kw["groups"]))
        labels = np.unique(kw["groups"], return_inverse=True)[1]
        gw = np.array(kw.get("group_weights", np.ones(G)))
        l1r = kw.get("l1_ratio", 0.0) if case["name"] == "SparseGroupLasso" else 0.0
        nr = np.array([np.linalg.norm(b[labels == g]) for g in range(G)])
        dl = np.array(kw.get("delta", np.zeros(G))) if case["name"] == "RidgedGroupLasso" else np.zeros(G)
        obj = (np.sum((y - X @ b) ** 2) / (2 * n) + l1r * case["alpha"] * np.abs(b).sum()
               + (1 - l1r) * case["alpha"] * (gw * nr).sum() + 0.5 * (dl * nr * nr).sum())
        assert abs(obj - case["objective"]) <= 1e-8 * abs(case["objective"])


def nlp_pre_case(case):
    """(X, y, sample_weight or None, constructor kwargs) of a "cases_preprocess" record."""
    pb = NLP["problems"][case["problem"]]
    X, y = np.array(pb["X"]), np.array(pb["y"])
    sw = np.array(pb["sample_weight"]) if case["use_sample_weight"] else None
    kw = dict(groups=np.array(pb["groups"]), alpha=case["alpha"], fit_intercept=case["fit_intercept"],
              standardize=case["standardize"])
    if "l1_ratio" in case:
        kw["l1_ratio"] = case["l1_ratio"]
    if "delta" in case:
        kw["delta"] = tuple(case["delta"])
    return X, y, sw, kw


def _pre_id(c):
    return (f"{c['name']}-{c['problem']}-{c['alpha']:.3g}-fi{int(c['fit_intercept'])}"
            f"-sw{int(c['use_sample_weight'])}-std{int(c['standardize'])}")


@pytest.mark.parametrize("case", NLP["cases_preprocess"], ids=_pre_id)
def test_oracle_preprocessing_and_standardize_match_third_party(case):
    """Intercepts, sample weights (rescaled to sum n, rows scaled by sqrt(sw), _base.py:207-227) and the
    standardized group norms ||X_g b_g|| / ||sqrtm(X_g'X_g + sqrt(delta_g) I) b_g|| (_lasso.py:249-252, :776-789):
    the generator restates them independently and solves with scipy's SLSQP / trust-exact."""
    X, y, sw, kw = nlp_pre_case(case)
    b, icpt = R.fit(case["name"], X, y, sample_weight=sw, **kw)
    ref = np.array(case["coef"])
    assert np.abs(b - ref).max() <= 1e-6 * np.abs(ref).max()
    assert np.array_equal(np.abs(b) > 1e-6, np.abs(ref) > 1e-6)
    assert abs(icpt - case["intercept"]) <= 1e-6 * max(1.0, abs(case["intercept"]))

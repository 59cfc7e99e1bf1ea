"""The CPU oracle against fixtures produced by the REFERENCE'S OWN code.

tests/golden/golden_reference.json was written by tests/golden/make_golden_reference.py, which runs
the reference's unmodified `Estimator.fit` (problem building, overlap expansion, adaptive loop,
fold-back, intercept) with a numeric cvxpy stand-in; only the conic solve is substituted, and its
result is KKT-certified on the problem the reference built (residual <= 8.5e-15 of the gradient
scale for every kept fixture).  Here the oracle's own restatement of the whole chain has to land
on the same numbers, and its objective has to BE the reference's objective as a function."""

import hashlib
import json
import os

import numpy as np
import pytest

import oracle.reference as R

REF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_reference.json")))


def ref_data(case):
    """Same stream as make_golden_reference.make_data (numpy Generator streams are stable)."""
    rng = np.random.default_rng(case["seed"])
    n, p, n_inf, noise = case["n"], case["p"], 5, 0.3
    X = rng.standard_normal((n, p))
    w = np.zeros(p)
    w[rng.choice(p, n_inf, replace=False)] = rng.standard_normal(n_inf) * 2.0
    y = X @ w + noise * rng.standard_normal(n) + 0.7
    sw = rng.random(n) + 0.25
    sha = hashlib.sha256(np.ascontiguousarray(X).tobytes() + np.ascontiguousarray(y).tobytes()
                         + np.ascontiguousarray(sw).tobytes()).hexdigest()[:16]
    assert sha == case["x_sha"], "fixture inputs no longer reproduce: regenerate golden_reference.json"
    return X, y, (sw if case["weighted"] else None)


def ref_kwargs(case):
    kw = dict(case["kwargs"])
    for k in ("groups", "group_weights", "delta"):
        if kw.get(k) is not None:
            kw[k] = np.asarray(kw[k], dtype=int if k == "groups" else float)
    return kw


def ref_id(case):
    kw = case["kwargs"]
    tags = [case["estimator"], f"a{kw.get('alpha')}"]
    tags += [t for t, on in (("std", kw.get("standardize")), ("icpt", case["fit_intercept"]), ("sw", case["weighted"]),
                             ("gw", kw.get("group_weights") is not None)) if on]
    return "-".join(tags) + f"-s{case['seed']}"


_FITS = {}


def _oracle_fit(case):
    """The oracle's chain for one fixture case, inner solves run to stationarity (solver_tol < 0: a
    duality gap of eps still leaves sqrt(eps) in the coefficients; the fixtures are stationary to
    rounding).  Shared by the two tests below."""
    key = ref_id(case)
    if key not in _FITS:
        X, y, sw = ref_data(case)
        _FITS[key] = R.fit(case["estimator"], X, y, fit_intercept=case["fit_intercept"], sample_weight=sw,
                           solver_tol=-1.0, max_sweeps=5000000, return_details=True, **ref_kwargs(case))
    return _FITS[key]


def test_fixture_file_is_certified():
    assert len(REF["cases"]) >= 56
    assert REF["worst_kkt_rel"] <= 1e-12
    names = {c["estimator"] for c in REF["cases"]}
    assert names == R.ESTIMATORS


@pytest.mark.parametrize("case", REF["cases"], ids=ref_id)
def test_oracle_chain_matches_reference_code(case):
    X, y, sw = ref_data(case)
    kw = ref_kwargs(case)
    b, icpt, det = _oracle_fit(case)
    ref = np.array(case["coef"])
    scale = max(np.abs(ref).max(), 1e-12)
    assert np.abs(b - ref).max() <= 1e-9 * scale + 1e-13
    assert np.array_equal(np.abs(b) > 1e-9 * scale, np.abs(ref) > 1e-9 * scale)
    assert abs(icpt - case["intercept"]) <= 1e-9 * max(1.0, abs(case["intercept"]))
    assert det["n_iter"] == case["n_iter"]
    # the reference's adaptive weights after its last update (_adaptive_lasso.py:196-204, 364-374, 712-726)
    wa = case["weights_after"]
    if "adaptive_coef_weights" in wa:
        np.testing.assert_allclose(det["w1"], wa["adaptive_coef_weights"], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(det["w2"], wa["adaptive_group_weights"], rtol=1e-7, atol=1e-12)
    elif "adaptive_weights" in wa:
        mine = det["w1"] if case["estimator"] == "AdaptiveLasso" else det["w2"]
        np.testing.assert_allclose(mine, wa["adaptive_weights"], rtol=1e-7, atol=1e-12)
    if "extended_coef_indices" in case:  # overlap expansion, _lasso.py:440-461
        idx, _, _ = R.expand_overlap(kw.get("group_list"), case["p"])
        assert idx.tolist() == case["extended_coef_indices"]


@pytest.mark.parametrize("case", REF["cases"], ids=ref_id)
def test_oracle_objective_is_the_reference_objective(case):
    """The reference's own objective expression, evaluated at random (dense and sparse) points,
    against the oracle's objective for the problem it thinks this is.  Equal as functions +
    the oracle's certified minimiser  =>  a minimiser of the reference's problem."""
    X, y, sw = ref_data(case)
    kw = ref_kwargs(case)
    _, _, det = _oracle_fit(case)
    if det.get("sgl_standardized"):
        # l1 on b with ||X_g b_g|| group norms: evaluated directly on the preprocessed design
        Xp, yp, _, _ = R.preprocess(X, y, sw, case["fit_intercept"])
        for probe in case["objective_probes"]:
            val = R.objective_sgl_standardized(Xp, yp, np.array(probe["beta"]), det["labels"], det["w1"], det["w2"])
            assert abs(val - probe["objective"]) <= 1e-11 * max(1.0, abs(probe["objective"]))
        return
    ps = det["pen_scale"]
    labels, G = det["labels"], len(det["w2"])
    standardized = bool(kw.get("standardize")) and case["estimator"].replace("Adaptive", "") != "Lasso"
    dl = np.zeros(G) if standardized else det["delta"]
    pen = R.Penalty(labels, det["w1"] * ps, det["w2"] * ps, dl)
    Xs, ys = det["X_solve"], det["y_solve"]
    if standardized:
        # the oracle solves in gamma_g = R_g b_g; recover R_g from X_solve_g = X_g R_g^{-1}
        Xp, _, _, _ = R.preprocess(X, y, sw, case["fit_intercept"])
        if "extended_coef_indices" in case:
            Xp = Xp[:, np.array(case["extended_coef_indices"])]
        n = Xp.shape[0]
    for probe in case["objective_probes"]:
        beta = np.array(probe["beta"])
        if standardized:
            gamma = np.zeros_like(beta)
            for g in range(G):
                idx = np.flatnonzero(labels == g)
                Rinv = np.linalg.lstsq(Xp[:, idx], Xs[:n, idx], rcond=None)[0]  # X_g Rinv = Xs_g
                gamma[idx] = np.linalg.solve(Rinv, beta[idx])
            val = R.objective(Xs, ys, gamma, pen) / ps
        else:
            val = R.objective(Xs, ys, beta, pen) / ps
        assert abs(val - probe["objective"]) <= 1e-11 * max(1.0, abs(probe["objective"]))

"""CPU tests: host-side logic of the estimator API, the C-ABI surface and the oracle."""

import ctypes
import os
import re

import numpy as np
import pytest

import oracle.reference as R
from sparselm_b200 import _build, _lib
from sparselm_b200.model import (
    AdaptiveGroupLasso,
    AdaptiveLasso,
    AdaptiveOverlapGroupLasso,
    AdaptiveRidgedGroupLasso,
    AdaptiveSparseGroupLasso,
    GroupLasso,
    Lasso,
    OverlapGroupLasso,
    RidgedGroupLasso,
    SparseGroupLasso,
)
from sparselm_b200.model._base import stack_specs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI surface ----------------------------------------------------------------------
def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "sparselm_b200.h")).read()
    declared = set(re.findall(r"\b(slm_[a-z_0-9]+)\s*\(", header))
    declared -= {"slm_ctx", "slm_batch"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    assert lib.slm_version() >= 1
    assert lib.slm_padded_cols(4096) == 4104 and lib.slm_padded_cols(6) == 8


def test_batch_struct_layout_matches_header():
    # field order of the ctypes mirror == declaration order in the header
    header = open(os.path.join(ROOT, "include", "sparselm_b200.h")).read()
    body = header[header.index("typedef struct slm_batch {"):header.index("} slm_batch;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b([A-Za-z_0-9]+)(?:\[SLM_MAX_FOLDS\])?\s*;", body)
    assert names == [f[0] for f in _lib.SlmBatch._fields_]


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sparselm_b200.engine import EngineError

    with pytest.raises(EngineError):
        Lasso(alpha=0.1).fit(np.eye(3), np.ones(3))
    h = ctypes.c_void_p()
    assert _lib.load().slm_create(0, ctypes.byref(h)) != 0


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sparselm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


# ---- error / warning contract (reference tests/test_lasso.py:203-260) -------------------
@pytest.fixture
def data():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((25, 20))
    y = rng.standard_normal(25)
    groups = rng.integers(0, 4, size=20)
    groups[:4] = np.arange(4)
    return X, y, groups


def test_bad_inputs(data):
    X, y, groups = data
    rng = np.random.default_rng(1)
    bad_groups = rng.integers(0, 6, size=X.shape[1] - 1)
    gw = np.ones(len(np.unique(bad_groups)))
    with pytest.raises(ValueError):
        GroupLasso(bad_groups, group_weights=gw).fit(X, y)
    with pytest.raises(TypeError):
        GroupLasso("groups", group_weights=gw).fit(X, y)
    with pytest.raises(ValueError):
        GroupLasso(bad_groups, group_weights=np.ones(len(np.unique(bad_groups)) - 1)).fit(X, y)
    with pytest.raises(TypeError):
        GroupLasso(groups, group_weights="weights").fit(X, y)
    lasso = SparseGroupLasso(groups)
    for bad in (-1.0, 2.0):
        lasso.l1_ratio = bad
        with pytest.raises(ValueError):
            lasso.fit(X, y)
        with pytest.raises(ValueError):
            SparseGroupLasso(groups, l1_ratio=bad).fit(X, y)
    with pytest.raises(ValueError):
        OverlapGroupLasso(group_list=[[0]] * (X.shape[1] - 1)).fit(X, y)
    with pytest.raises(ValueError):
        RidgedGroupLasso(groups, delta=(1.0, 2.0)).fit(X, y)
    with pytest.raises(ValueError):
        Lasso(alpha=-1.0).fit(X, y)
    with pytest.raises(TypeError):
        Lasso(solver_options=[1]).fit(X, y)


def test_constructor_signatures_match_reference():
    import inspect

    sig = lambda c: list(inspect.signature(c).parameters)  # noqa: E731
    assert sig(Lasso) == ["alpha", "fit_intercept", "copy_X", "warm_start", "solver", "solver_options"]
    assert sig(GroupLasso)[:4] == ["groups", "alpha", "group_weights", "standardize"]
    assert sig(OverlapGroupLasso)[:4] == ["group_list", "alpha", "group_weights", "standardize"]
    assert sig(SparseGroupLasso)[:4] == ["groups", "l1_ratio", "alpha", "group_weights"]
    assert sig(RidgedGroupLasso)[:4] == ["groups", "alpha", "delta", "group_weights"]
    assert sig(AdaptiveLasso)[:5] == ["alpha", "max_iter", "eps", "tol", "update_function"]
    assert sig(AdaptiveGroupLasso)[:7] == ["groups", "alpha", "group_weights", "max_iter", "eps", "tol",
                                           "update_function"]
    for cls in (AdaptiveLasso, AdaptiveGroupLasso, AdaptiveOverlapGroupLasso, AdaptiveSparseGroupLasso,
                AdaptiveRidgedGroupLasso):
        e = cls()
        assert (e.max_iter, e.eps, e.tol, e.warm_start) == (3, 1e-6, 1e-10, True)
    assert RidgedGroupLasso().delta == (1.0,) and SparseGroupLasso().l1_ratio == 0.5


def test_get_set_params_clone_roundtrip(data):
    from sklearn.base import clone

    _, _, groups = data
    for est in (Lasso(alpha=0.3), SparseGroupLasso(groups, l1_ratio=0.2, alpha=2.0),
                AdaptiveOverlapGroupLasso(group_list=[[0, 1]] * 20, max_iter=5),
                AdaptiveRidgedGroupLasso(groups, delta=(3.0,))):
        c = clone(est)
        assert type(c) is type(est)
        for k, v in est.get_params().items():
            assert c.get_params()[k] is v or np.array_equal(c.get_params()[k], v)
        c.set_params(alpha=7.0)
        assert c.alpha == 7.0


# ---- penalty plumbing (reference tests/test_lasso.py:263-309) -----------------------------
def test_problem_specs_follow_reference_rules(data):
    _, _, groups = data
    p = 20
    G = len(np.unique(groups))
    gw = np.arange(1, G + 1, dtype=float)
    s = SparseGroupLasso(groups, l1_ratio=0.25, alpha=0.5, group_weights=gw)._problem_spec(p)
    assert s.lam1 == 0.25 * 0.5
    np.testing.assert_allclose(s.w2, 0.75 * 0.5 * gw)
    # groups made contiguous: permuted labels are sorted
    assert np.all(np.diff(groups[s.col_perm]) >= 0)
    assert s.gptr[-1] == p and len(s.gptr) == G + 1
    r = RidgedGroupLasso(groups, alpha=2.0, delta=(4.0,))._problem_spec(p)
    np.testing.assert_array_equal(r.d2, 4.0 * np.ones(G))
    r = RidgedGroupLasso(groups, alpha=2.0, delta=3.0 * np.ones(G))._problem_spec(p)
    np.testing.assert_array_equal(r.d2, 3.0 * np.ones(G))
    a = AdaptiveGroupLasso(groups, alpha=0.5, group_weights=gw)._problem_spec(p)
    np.testing.assert_array_equal(a.w2, 0.5 * np.ones(G))  # no group_weights in pass 1
    assert a.adaptive["a2"] == 0.5 and a.adaptive["a1"] is None
    np.testing.assert_array_equal(a.gw, gw)
    asg = AdaptiveSparseGroupLasso(groups, l1_ratio=0.5, alpha=0.5)._problem_spec(p)
    assert asg.lam1 == 0.25 and asg.adaptive["a1"] == 0.25 and asg.adaptive["a2"] == 0.25
    al = AdaptiveLasso(alpha=0.3)._problem_spec(p)
    assert al.lam1 == 0.3 and al.adaptive["a1"] == 0.3 and al.adaptive["alpha"] == 0.3


def test_overlap_expansion_matches_reference_scan():
    rng = np.random.default_rng(2)
    p = 17
    group_list = [list(rng.choice(6, size=rng.integers(1, 4), replace=False)) for _ in range(p)]
    est = OverlapGroupLasso(group_list=group_list)
    ext_idx, gptr, ng = est._expansion(p)
    ref_idx, ref_groups, ref_ng = R.expand_overlap(group_list, p)
    np.testing.assert_array_equal(ext_idx, ref_idx)
    assert ng == ref_ng
    np.testing.assert_array_equal(np.repeat(np.arange(ng), np.diff(gptr)), ref_groups)


def test_stack_specs_shapes(data):
    _, _, groups = data
    specs = [SparseGroupLasso(groups, alpha=a)._problem_spec(20) for a in (0.1, 0.2, 0.3)]
    assert len({s.key for s in specs}) == 1
    g = stack_specs(specs)
    assert g.K == 3 and g.W2.shape == (len(np.unique(groups)), 3)
    np.testing.assert_allclose(g.lam1, [0.05, 0.1, 0.15])
    assert SparseGroupLasso(groups, alpha=0.1, fit_intercept=True)._problem_spec(20).key != specs[0].key


# ---- model selection host logic --------------------------------------------------------------
def test_one_std_rule_and_partition_helpers():
    from sklearn.model_selection import KFold, ShuffleSplit

    from sparselm_b200.model_selection import _is_partition, _metric, _select_best_index_onestd

    res = {
        "rank_test_score": np.array([3, 1, 2, 4]),
        "mean_test_score": np.array([-1.3, -1.0, -1.05, -2.0]),
        "std_test_score": np.array([0.1, 0.1, 0.1, 0.1]),
        "param_alpha": np.ma.MaskedArray([0.01, 0.1, 1.0, 10.0]),
        "params": [{}] * 4,
    }
    # candidates with alpha >= 0.1: target mean = -1.0 - 0.1 = -1.1 -> alpha=1.0 (-1.05) is closest
    assert _select_best_index_onestd(True, "score", res) == 2
    n = 23
    assert _is_partition(list(KFold(5).split(np.zeros(n))), n)
    assert _is_partition(list(KFold(5, shuffle=True, random_state=0).split(np.zeros(n))), n)
    assert not _is_partition(list(ShuffleSplit(3, random_state=0).split(np.zeros(n))), n)
    sse, sae = np.array([4.0, 0.0]), np.array([2.0, 0.0])
    np.testing.assert_allclose(_metric("neg_root_mean_squared_error", sse, sae, 4, 8.0), [-1.0, 0.0])
    np.testing.assert_allclose(_metric("neg_mean_squared_error", sse, sae, 4, 8.0), [-1.0, 0.0])
    np.testing.assert_allclose(_metric("neg_mean_absolute_error", sse, sae, 4, 8.0), [-0.5, 0.0])
    np.testing.assert_allclose(_metric("r2", sse, sae, 4, 8.0), [0.5, 1.0])


def test_miqp_names_exist_and_fail_loudly():
    from sparselm_b200.model import L1L0, L2L0, BestSubsetSelection, RegularizedL0, RidgedBestSubsetSelection

    for cls in (L1L0, L2L0, BestSubsetSelection, RegularizedL0, RidgedBestSubsetSelection):
        with pytest.raises(NotImplementedError):
            cls().fit(np.eye(3), np.ones(3))


# ---- host logic added with the widening rows (no GPU needed) ------------------------------
def test_device_metrics_forms():
    from sparselm_b200.model_selection import _device_metrics

    assert _device_metrics(None) == "r2"
    assert _device_metrics("neg_mean_absolute_error") == "neg_mean_absolute_error"
    assert _device_metrics("explained_variance") is None
    assert _device_metrics(["r2", "neg_mean_squared_error"]) == {"r2": "r2",
                                                                 "neg_mean_squared_error": "neg_mean_squared_error"}
    assert _device_metrics(("r2", "r2")) is None  # duplicate names: sklearn's own error path
    assert _device_metrics({"a": "r2", "b": "neg_root_mean_squared_error"}) == {
        "a": "r2", "b": "neg_root_mean_squared_error"}
    assert _device_metrics({"a": "r2", "b": "max_error"}) is None
    assert _device_metrics(lambda est, X, y: 0.0) is None
    assert _device_metrics([]) is None


def test_metric_formulas_match_sklearn():
    from sklearn.metrics import mean_absolute_error, mean_squared_error, r2_score

    from sparselm_b200.model_selection import _metric

    rng = np.random.default_rng(0)
    y, yp = rng.normal(size=40), rng.normal(size=40)
    sse, sae, sst = ((y - yp) ** 2).sum(), np.abs(y - yp).sum(), ((y - y.mean()) ** 2).sum()
    assert _metric("r2", sse, sae, 40, sst) == pytest.approx(r2_score(y, yp))
    assert _metric("neg_mean_squared_error", sse, sae, 40, sst) == pytest.approx(-mean_squared_error(y, yp))
    assert _metric("neg_root_mean_squared_error", sse, sae, 40, sst) == pytest.approx(-np.sqrt(mean_squared_error(y, yp)))
    assert _metric("neg_mean_absolute_error", sse, sae, 40, sst) == pytest.approx(-mean_absolute_error(y, yp))
    # constant target: sklearn's force_finite convention
    assert _metric("r2", np.array([0.0, 1.0]), None, 5, 0.0).tolist() == [1.0, 0.0]


def test_warm_start_views_are_lazy_and_first_fold_wins():
    import torch

    from sparselm_b200.model_selection import _WarmStarts

    B = torch.arange(2 * 3 * 8, dtype=torch.float64).reshape(2, 3, 8)  # [F, p, ldz]
    w = _WarmStarts()
    idxs = np.array([5, 2, 9])  # batch columns -> candidate indices
    w.add(B, idxs, [np.array([0, 2]), np.array([1, 2])])  # fold 0 solved 5 and 9, fold 1 solved 2 and 9
    assert torch.equal(w.get(5), B[0, :, 0])
    assert torch.equal(w.get(9), B[0, :, 1])  # the first fold that solved it
    assert torch.equal(w[2], B[1, :, 0])
    assert w.get(7) is None
    with pytest.raises(KeyError):
        w[7]

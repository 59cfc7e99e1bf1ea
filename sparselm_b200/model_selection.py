"""Hyper-parameter search: GridSearchCV with the one-standard-error rule and
LineSearchCV (reference: src/sparselm/model_selection.py).

The reference runs ``n_candidates x n_splits`` independent Python fits through
joblib (model_selection.py:304-323), each rebuilding its cvxpy problem.  Here the
whole (candidate, fold) grid of an engine-backed estimator is ONE batch on the
GPU: the per-fold Gram matrices are built once, every candidate is a column of
the batched accelerated proximal-gradient solve, and the CV scores are computed
on the device.  Anything the batched seam does not cover (foreign estimators,
fit params, exotic scorers or splitters) takes sklearn's generic per-fit path,
which still fits on the engine through ``estimator.fit``.
"""

from __future__ import annotations

import numbers
import re
import time
from copy import deepcopy

import numpy as np
from numpy.ma import MaskedArray
from scipy.stats import rankdata
from sklearn.base import clone
from sklearn.model_selection import GridSearchCV as _SkGridSearchCV
from sklearn.model_selection import ParameterGrid, check_cv
from sklearn.model_selection._search import BaseSearchCV
from sklearn.utils.validation import check_X_y

from .engine import EngineError, get_engine
from .model._base import EngineRegressor, _to_original_order, solve_specs

__all__ = ["GridSearchCV", "LineSearchCV"]

_DEVICE_SCORERS = {None, "r2", "neg_root_mean_squared_error", "neg_mean_squared_error",
                   "neg_mean_absolute_error"}


def _device_metrics(scoring):
    """`scoring` as the batched seam understands it: one device scorer name, or a dict
    {metric name: device scorer name} for sklearn's list / tuple / dict multi-metric forms;
    None when some scorer has to run on the host (callables, other metrics)."""
    if scoring is None:
        return "r2"
    if isinstance(scoring, str):
        return scoring if scoring in _DEVICE_SCORERS else None
    if isinstance(scoring, (list, tuple, set)):
        names = list(scoring)
        if not names or len(set(names)) != len(names):
            return None
        scoring = {nm: nm for nm in names}
    if isinstance(scoring, dict) and scoring and all(
            isinstance(k, str) and isinstance(v, str) and v in _DEVICE_SCORERS for k, v in scoring.items()):
        return dict(scoring)
    return None


def _select_best_index_onestd(refit, refit_metric, results):
    """One-standard-error rule (reference model_selection.py:190-223): among the
    candidates whose summed non-negative numeric hyper-parameters are at least those
    of the best-scoring candidate, take the one whose mean score is closest to
    (best mean - its std)."""
    if callable(refit):
        best_index = refit(results)
        if not isinstance(best_index, numbers.Integral):
            raise TypeError("best_index_ returned is not an integer")
        if best_index < 0 or best_index >= len(results["params"]):
            raise IndexError("best_index_ index out of range")
        return best_index
    opt_index = results[f"rank_test_{refit_metric}"].argmin()
    m = results[f"mean_test_{refit_metric}"][opt_index]
    sig = results[f"std_test_{refit_metric}"][opt_index]
    metrics = results[f"mean_test_{refit_metric}"]
    params = []
    for name in [key for key in results if re.match(r"^param_(\w+)", key)]:
        if all(isinstance(val, numbers.Number) for val in results[name]):
            pv = np.array(results[name], dtype=float)
            if np.all(pv > -1e-9):
                params.append(pv)
    params_sum = np.sum(params, axis=0)
    one_std_dists = np.abs(metrics - m + sig)
    candidates = np.arange(len(metrics))[params_sum >= params_sum[opt_index]]
    return candidates[np.argmin(one_std_dists[candidates])]


def _is_partition(splits, n):
    seen = np.zeros(n, dtype=np.int64)
    for train, test in splits:
        seen[test] += 1
        if len(train) + len(test) != n:
            return False
        mask = np.ones(n, dtype=bool)
        mask[test] = False
        if not np.array_equal(np.flatnonzero(mask), np.sort(train)):
            return False
    return bool(np.all(seen == 1))


def _metric(scoring, sse, sae, n_rows, sst):
    """Score per column from residual sums (greater is better, sklearn convention)."""
    if scoring == "neg_root_mean_squared_error":
        return -np.sqrt(sse / n_rows)
    if scoring == "neg_mean_squared_error":
        return -sse / n_rows
    if scoring == "neg_mean_absolute_error":
        return -sae / n_rows
    # r2 (estimator.score / "r2"), with sklearn's force_finite convention
    if sst == 0.0:
        return np.where(sse == 0.0, 1.0, 0.0)
    return 1.0 - sse / sst


class _WarmStarts:
    """candidate index -> device view [pe] of its solution on the first training fold that
    solved it (start point of the refit).  The views are made on demand: building one per
    (fold, candidate) eagerly costs more host time than the scoring kernels it delays."""

    def __init__(self):
        self._batches = []

    def add(self, B, idxs, mine):
        self._batches.append((B, np.asarray(idxs), [np.asarray(m) for m in mine]))

    def get(self, ci, default=None):
        for B, idxs, mine in self._batches:
            for f, m in enumerate(mine):
                pos = np.nonzero(idxs[m] == ci)[0] if len(m) else ()
                if len(pos):
                    return B[f, :, int(pos[0])]
        return default

    def __getitem__(self, ci):
        v = self.get(ci)
        if v is None:
            raise KeyError(ci)
        return v


def _fold_key(est, spec):
    return (bool(est.fit_intercept), None if spec.col_perm is None else spec.col_perm.tobytes())


def prepare_folds(engine, X, yv, test_folds, est, spec, cache=None, cache_key=None, shard=None, sample_weight=None,
                  score_folds=None):
    """Device-resident design + Grams for (fit_intercept, column order) of `est`/`spec`,
    through the FoldData cache of a LineSearchCV when there is one.  Only enqueues GPU
    work (H2D copies, packing, Gram build): the caller can keep working on the host."""
    ck = None
    if cache is not None and cache_key is not None:
        # keyed on the full content of the split: a shuffling splitter draws a new partition per
        # line, and two partitions can agree in fold sizes and first indices
        import hashlib

        fold_hash = hashlib.sha1(np.concatenate([np.asarray(t, dtype=np.int64) for t in test_folds]).tobytes()).hexdigest()
        ck = (cache_key, _fold_key(est, spec), tuple(len(t) for t in test_folds), fold_hash)
        if ck in cache:
            return cache[ck]
    fd = engine.prepare(X, yv, test_folds, est.fit_intercept, sample_weight, col_perm=spec.col_perm, shard=shard,
                        score_folds=score_folds)
    if ck is not None:
        cache[ck] = fd
    return fd


def batched_cv(engine, X, y, test_folds, ests, specs, opts, scoring, return_train_score=False, cache=None,
               cache_key=None, shard=None, fds=None, sample_weight=None):
    """Solve every (candidate, fold) problem as device batches and score them.

    X: (n, p) numpy array or torch tensor (host, pinned or already on the device);
    y: (n,) numpy array; test_folds: list of index arrays partitioning the rows;
    ests/specs: one configured estimator and its ProblemSpec per candidate.
    shard: optional parallel.GridShard -- this rank builds the Gram of its rows, solves
    its share of the (fold, candidate) grid, and the score tables are summed over ranks.
    sample_weight: optional (n,) strictly positive fit weights, split per fold like the
    reference's fit_params (model_selection.py:266); the scores stay unweighted.
    Returns test_scores [n_cand, n_splits] (+ train scores, timings, solver info and
    the device-resident FoldData objects keyed by (fit_intercept, column order)).
    """
    n, p = X.shape
    n_splits, n_cand = len(test_folds), len(specs)
    yv = np.asarray(y, dtype=np.float64)
    torch = engine.torch
    # one scorer name, or {metric name: scorer name}: every device scorer derives from the same
    # two residual sums, so a multi-metric search costs nothing extra on the GPU
    multi = not isinstance(scoring, str)
    metrics = dict(scoring) if multi else {"score": scoring}
    sharded = shard is not None and shard.world > 1
    # residual sums per (candidate, fold): sum of squares and of absolute values over the test rows
    # (and over the training rows when train scores are asked for).  A sharded search fills in the
    # sums of ITS rows of every fold (scoring is row-sharded like the Gram build) and the tables are
    # summed over the ranks; the metrics are formed from the complete sums afterwards.
    want_train = bool(return_train_score)
    SSE, SAE = np.zeros((n_cand, n_splits)), np.zeros((n_cand, n_splits))
    TSSE, TSAE = (np.zeros((n_cand, n_splits)), np.zeros((n_cand, n_splits))) if want_train else (None, None)
    fit_time = np.zeros(n_cand)
    score_time = np.zeros(n_cand)
    info = dict(n_iter=np.zeros((n_cand, n_splits), dtype=int), status=np.zeros((n_cand, n_splits), dtype=int),
                gap=np.zeros((n_cand, n_splits)), n_pass=np.zeros((n_cand, n_splits), dtype=int))
    need_r2 = "r2" in metrics.values()
    sst_test = np.array([float(((yv[t] - yv[t].mean()) ** 2).sum()) if need_r2 else 0.0 for t in test_folds])
    tot_sum, tot_sq = float(yv.sum()), float((yv * yv).sum())
    n_unconverged = 0
    newton_stats = {}
    iters_run = 0

    fds = {} if fds is None else dict(fds)  # may arrive pre-populated (prepare started early)
    if n_cand and _fold_key(ests[0], specs[0]) not in fds:
        # enqueue the GPU side of the preparation (H2D, packing, Gram build) before the host-side grouping and
        # ordering of the candidates below: that Python work then runs while the GPU is busy
        x_dev = isinstance(X, torch.Tensor) and X.is_cuda
        score_folds = set() if (sharded and cache is None and not x_dev) else None
        fds[_fold_key(ests[0], specs[0])] = prepare_folds(engine, X, yv, test_folds, ests[0], specs[0], cache, cache_key,
                                                         shard, sample_weight, score_folds)
    warm = _WarmStarts()  # candidate -> device view [pe] of its solution on some training fold (refit start)
    batches = {}
    for ci, s in enumerate(specs):
        batches.setdefault(s.key, []).append(ci)
    for key, idxs in batches.items():
        s0, e0 = specs[idxs[0]], ests[idxs[0]]
        fkey = _fold_key(e0, s0)
        # batch columns in order of increasing penalty strength (dense iterates first): the
        # row-sparse apply shares one support list per chunk of adjacent columns
        idxs = np.asarray(sorted(idxs, key=lambda ci: (specs[ci].strength, ci)))
        K = len(idxs)
        if sharded:
            from .parallel import assign_columns

            owner = assign_columns(n_splits, K, shard.world)  # owner[f][k]: rank that solves column k of fold f
            mine = [np.flatnonzero(owner[f] == shard.rank) for f in range(n_splits)]
        else:
            owner = None
            mine = [np.arange(K)] * n_splits
        if fkey not in fds:
            # a sharded rank only needs its own slice of every fold's rows on the device (Gram build
            # and scoring are both row-sharded) when the design comes from the host; a design that is already
            # device-resident, or the cached one of a LineSearchCV, is kept whole
            x_dev = isinstance(X, torch.Tensor) and X.is_cuda
            score_folds = set() if (sharded and cache is None and not x_dev) else None
            fds[fkey] = prepare_folds(engine, X, yv, test_folds, e0, s0, cache, cache_key, shard, sample_weight,
                                      score_folds)
        fd = fds[fkey]
        t0 = time.perf_counter()
        out = solve_specs(engine, fd, [[specs[idxs[k]] for k in mine[f]] for f in range(n_splits)], **opts)
        t1 = time.perf_counter()
        warm.add(out["B"], idxs, mine)
        wtd = bool(fd.extra.get("weighted"))
        if not sharded or fd.extra.get("partial_folds") is None:
            # every row is resident: each rank scores the columns it solved on whole folds (a sharded grid then
            # needs no exchange of coefficients; the partial tables are summed with the info tables below).
            # All (fold, row range) problems go through one batched scoring launch.
            icpt = (lambda f: out["intercept"][f]) if fd.fit_intercept else (lambda f: None)
            items, where = [], []
            for f in range(n_splits):
                kf = len(mine[f])
                if kf == 0:
                    continue
                items.append((fd.row_ptr[f], fd.row_ptr[f + 1], out["coef"][f], kf, icpt(f)))
                where.append((0, f))
                if want_train:
                    for r0, r1 in ((0, fd.row_ptr[f]), (fd.row_ptr[f + 1], n)):
                        if r1 > r0:
                            items.append((r0, r1, out["coef"][f], kf, icpt(f)))
                            where.append((1, f))
            sc = np.zeros((n_splits, 2, K))
            tsc = np.zeros((n_splits, 2, K)) if want_train else None
            if items:
                res = engine.cv_score_many(fd.Xa, p, items, rows_scaled=wtd).cpu().numpy()  # one D2H
                for i, (kind, f) in enumerate(where):
                    tgt = sc if kind == 0 else tsc
                    tgt[f][:, mine[f]] += res[i][:, :len(mine[f])]
        else:
            sc, tsc = _sharded_residual_sums(engine, shard, fd, out, owner, mine, p, K, n_splits, wtd, want_train)
        t2 = time.perf_counter()
        for f in range(n_splits):
            SSE[idxs, f], SAE[idxs, f] = sc[f, 0], sc[f, 1]
            if want_train:
                TSSE[idxs, f], TSAE[idxs, f] = tsc[f, 0], tsc[f, 1]
            kf = len(mine[f])
            ci = idxs[mine[f]]
            info["n_iter"][ci, f] = out["n_iter"][f, :kf]
            info["status"][ci, f] = out["status"][f, :kf]
            info["gap"][ci, f] = out["gap"][f, :kf]
            info["n_pass"][ci, f] = out["n_pass"][f, :kf]
        fit_time[idxs] = (t1 - t0) / (K * n_splits)
        score_time[idxs] = (t2 - t1) / (K * n_splits)
        n_unconverged += int(out["n_unconverged"])
        iters_run += int(out["iters_run"])
        for k, v in (out.get("newton") or {}).items():  # second-order phase of this rank's batches (engine._run_batch)
            newton_stats[k] = newton_stats.get(k, 0) + v
    if sharded:
        # the last exchange of the sharded grid: one sum of the residual-sum and info tables
        tabs = [SSE, SAE] + ([TSSE, TSAE] if want_train else [])
        keys = list(info)
        tabs += [info[k].astype(np.float64) for k in keys] + [np.array([[float(n_unconverged)]])]
        flat = shard.allreduce_sum_numpy(np.concatenate([t.ravel() for t in tabs]), engine.device)
        parts, o = [], 0
        for t in tabs:
            parts.append(flat[o:o + t.size].reshape(t.shape))
            o += t.size
        SSE, SAE = parts.pop(0), parts.pop(0)
        if want_train:
            TSSE, TSAE = parts.pop(0), parts.pop(0)
        for k in keys:
            info[k] = parts.pop(0).astype(info[k].dtype)
        n_unconverged = int(round(float(parts.pop(0)[0, 0])))
    test_tabs = {m: np.empty((n_cand, n_splits)) for m in metrics}
    train_tabs = {m: np.empty((n_cand, n_splits)) for m in metrics} if want_train else None
    for f in range(n_splits):
        nt = len(test_folds[f])
        for m, scorer in metrics.items():
            test_tabs[m][:, f] = _metric(scorer, SSE[:, f], SAE[:, f], nt, sst_test[f])
        if want_train:
            ntr = n - nt
            s_tr = tot_sum - float(yv[test_folds[f]].sum())
            q_tr = tot_sq - float((yv[test_folds[f]] ** 2).sum())
            for m, scorer in metrics.items():
                train_tabs[m][:, f] = _metric(scorer, TSSE[:, f], TSAE[:, f], ntr, q_tr - s_tr * s_tr / ntr)
    test_scores = test_tabs if multi else test_tabs["score"]
    train_scores = None
    if want_train:
        train_scores = train_tabs if multi else train_tabs["score"]
    return dict(test_scores=test_scores, train_scores=train_scores, fit_time=fit_time, score_time=score_time,
                info=info, fds=fds, n_unconverged=n_unconverged, iters_run=iters_run, warm=warm, newton=newton_stats)


def _sharded_residual_sums(engine, shard, fd, out, owner, mine, p, K, n_splits, wtd, want_train):
    """Row-sharded scoring of a sharded grid (north_star: "NCCL all-gather of CV scores and
    coefficients").  Every rank publishes the coefficients (+ intercepts) of the columns it solved,
    ONE all-gather makes all K x n_splits solutions resident everywhere (C3: 10 MB per rank), and each
    rank forms the residual sums of ITS slice of every fold's rows -- the slice it uploaded for the
    Gram build, so a host-resident design crosses the bus once, 1/world per rank.  Returns this
    rank's partial sums [n_splits, 2, K] (test rows) and the same for the training rows (or None)."""
    import ctypes

    from .parallel import gather_layout, scoring_plan

    torch = engine.torch
    W = shard.world
    # send layout: [p + 1][ldg], the columns this rank solved, fold after fold (only what was solved travels: C3 on
    # 8 ranks, 2.6 MB per rank)
    kfr, pad, off, ldg = gather_layout(owner, W)
    send = torch.zeros((p + 1, ldg), dtype=torch.float64, device=engine.device)
    for f in range(n_splits):
        kf = len(mine[f])
        if kf:
            o = int(off[shard.rank, f])
            send[:p, o:o + kf] = out["coef"][f][:, :kf]
            if fd.fit_intercept:
                send[p, o:o + kf] = out["intercept"][f][:kf]
    recv = torch.empty((W,) + tuple(send.shape), dtype=torch.float64, device=engine.device)
    comm = shard.comm_ptr(engine.device)
    if comm is not None:
        engine._ck(engine.lib.slm_gather_results(engine.h, ctypes.c_void_p(comm), engine._ptr(send), engine._ptr(recv),
                                                 send.numel(), engine.stream), "slm_gather_results")
    else:
        shard.all_gather_(send, recv)
    # every (fold, solving rank) group against this rank's slice of the fold's rows -- and, for train scores,
    # against its slices of the other folds -- as ONE batched scoring launch
    plan = scoring_plan(owner, shard.fold_row_ranges(fd.row_ptr), W, want_train)
    items = []
    for lo, hi, r, f, kind, cols in plan:
        o, kf = int(off[r, f]), len(cols)
        items.append((lo, hi, recv[r, :p, o:o + int(pad[r, f])], kf, recv[r, p, o:o + kf] if fd.fit_intercept else None))
    part = np.zeros((n_splits, 2, K))
    tpart = np.zeros((n_splits, 2, K)) if want_train else None
    if items:
        res = engine.cv_score_many(fd.Xa, p, items, rows_scaled=wtd).cpu().numpy()  # one D2H
        for i, (lo, hi, r, f, kind, cols) in enumerate(plan):
            tgt = part if kind == 0 else tpart
            tgt[f][:, cols] += res[i][:, :len(cols)]
    return part, tpart


class GridSearchCV(_SkGridSearchCV):
    """Exhaustive search over a parameter grid, batched on the GPU for engine-backed
    estimators.  Same constructor as the reference (model_selection.py:160-187):
    sklearn's GridSearchCV plus ``opt_selection_method`` in {"max_score",
    "one_std_score"} and default scoring "neg_root_mean_squared_error"; fitted
    attributes ``best_params_``, ``best_score_``, ``best_score_std_``,
    ``best_estimator_``, ``cv_results_`` (:376-422).
    """

    def __init__(self, estimator, param_grid, *, opt_selection_method="max_score",
                 scoring="neg_root_mean_squared_error", n_jobs=None, refit=True, cv=None, verbose=0,
                 pre_dispatch="2*n_jobs", error_score=np.nan, return_train_score=False):
        super().__init__(estimator=estimator, param_grid=param_grid, scoring=scoring, n_jobs=n_jobs,
                         refit=refit, cv=cv, verbose=verbose, pre_dispatch=pre_dispatch,
                         error_score=error_score, return_train_score=return_train_score)
        self.opt_selection_method = opt_selection_method

    # sklearn calls self._select_best_index(self.refit, refit_metric, results)
    def _select_best_index(self, refit, refit_metric, results):
        if self.opt_selection_method == "max_score":
            return _SkGridSearchCV._select_best_index(refit, refit_metric, results)
        if self.opt_selection_method == "one_std_score":
            return _select_best_index_onestd(refit, refit_metric, results)
        raise NotImplementedError(f"Method {self.opt_selection_method} not implemented!")

    _select_best_index_onestd = staticmethod(_select_best_index_onestd)

    # ------------------------------------------------------------------ #
    def fit(self, X, y=None, **params):
        """Run the search.  Engine-backed estimators with batchable grids are solved as
        one device batch; everything else goes through sklearn's per-fit loop."""
        plan = self._batch_plan(X, y, params)
        if plan is None:
            super().fit(X, y, **params)
            if hasattr(self, "best_index_") and not callable(self.refit):
                self.best_score_std_ = self.cv_results_["std_test_score"][self.best_index_] \
                    if "std_test_score" in self.cv_results_ else None
            self.batched_ = False
            return self
        return self._fit_batched(plan)

    # ------------------------------------------------------------------ #
    def _batch_plan(self, X, y, params):
        est = self.estimator
        if not isinstance(est, EngineRegressor) or y is None or not getattr(est, "_batchable", True):
            return None
        # fit params: only sample_weight is understood by the batched seam (strictly positive:
        # the unweighted CV scores are recovered from the sqrt(sw)-scaled rows)
        sw = None
        for k, v in params.items():
            if v is None:
                continue
            if k != "sample_weight":
                return None
            sw = v
        if _device_metrics(self.scoring) is None:
            return None
        if callable(self.refit) or hasattr(X, "columns"):
            return None
        if not isinstance(_device_metrics(self.scoring), str) and self.refit is not False and (
                not isinstance(self.refit, str) or self.refit not in _device_metrics(self.scoring)):
            return None  # sklearn's own loop raises its multi-metric refit error
        if self.opt_selection_method not in ("max_score", "one_std_score"):
            raise NotImplementedError(f"Method {self.opt_selection_method} not implemented!")
        try:
            # finiteness is checked on the device (engine.prepare) instead of a host pass over X
            Xv, yv = check_X_y(X, y, dtype=np.float64, y_numeric=True, ensure_min_samples=2,
                               ensure_all_finite=False)
        except Exception:
            return None
        n, p = Xv.shape
        if sw is not None:
            try:
                sw = np.asarray(sw, dtype=np.float64).reshape(-1)
            except Exception:
                return None
            if sw.shape[0] != n or not np.all(np.isfinite(sw)) or not np.all(sw > 0):
                return None
        cv = check_cv(self.cv, yv, classifier=False)
        splits = list(cv.split(Xv, yv, None))
        if len(splits) < 2 or len(splits) > 15 or not _is_partition(splits, n):
            return None
        candidates = list(ParameterGrid(self.param_grid))
        valid = set(est.get_params(deep=False))
        if any(not set(c) <= valid for c in candidates):
            return None
        ests, specs = [], []
        pre_fds = {}
        try:
            # one working estimator re-parametrised per candidate (cloning 100 estimators and
            # re-deriving their group structure costs more host time than the GPU solve)
            work = clone(est)
            import warnings as _w
            from types import SimpleNamespace

            for ci, c in enumerate(candidates):
                work.set_params(**c)
                if ci == 1:
                    # from the second candidate on only the grid's own parameters are re-validated, and the
                    # group structure memoised for the first candidate is reused without re-hashing `groups`
                    # (the working estimator is private to this loop)
                    work.__dict__["_validate_only"] = set().union(*[set(cc) for cc in candidates])
                    work.__dict__["_groups_frozen"] = True
                with _w.catch_warnings():
                    if ci > 0:  # structural warnings (e.g. groups=None) are emitted once
                        _w.simplefilter("ignore", UserWarning)
                    work._validate_hyperparams(Xv, yv)
                ests.append(SimpleNamespace(fit_intercept=bool(work.fit_intercept)))
                specs.append(work._problem_spec(p))
                if ci == 0:
                    # the design of the first candidate starts its way to the device (H2D, packing,
                    # Gram build are only enqueued) while the host describes the other candidates
                    opts0 = est._engine_options()
                    engine = get_engine(opts0.pop("device", None))
                    cache = getattr(self, "_fd_cache", None)
                    shard = getattr(self, "_shard", None)
                    score_folds = None
                    if shard is not None and shard.world > 1 and cache is None:
                        # host rows this rank needs: its slice of every fold (Gram build and scoring
                        # are both row-sharded)
                        score_folds = set()
                    pre_fds[_fold_key(ests[0], specs[0])] = prepare_folds(
                        engine, Xv, yv, [np.asarray(test) for _, test in splits], ests[0], specs[0], cache,
                        None if cache is None else (id(Xv), id(yv), None if sw is None else sw.tobytes()),
                        shard, sw, score_folds)
        except (NotImplementedError, EngineError):
            raise
        except Exception:
            return None  # sklearn's loop applies error_score semantics per candidate
        return dict(X=Xv, y=yv, n=n, p=p, splits=splits, candidates=candidates, ests=ests, specs=specs,
                    fds=pre_fds, sample_weight=sw)

    def _fit_batched(self, plan):
        Xv, yv, n, p = plan["X"], plan["y"], plan["n"], plan["p"]
        splits, candidates, specs, ests = plan["splits"], plan["candidates"], plan["specs"], plan["ests"]
        n_splits, n_cand = len(splits), len(candidates)
        base = self.estimator
        opts = base._engine_options()
        engine = get_engine(opts.pop("device", None))
        scoring = _device_metrics(self.scoring)
        multi = not isinstance(scoring, str)
        test_folds = [np.asarray(test) for _, test in splits]
        cache = getattr(self, "_fd_cache", None)
        sw = plan.get("sample_weight")
        cache_key = None if cache is None else (id(plan["X"]), id(plan["y"]), None if sw is None else sw.tobytes())
        res = batched_cv(engine, Xv, yv, test_folds, ests, specs, opts, scoring,
                         return_train_score=self.return_train_score, cache=cache, cache_key=cache_key,
                         shard=getattr(self, "_shard", None), fds=plan.get("fds"), sample_weight=sw)
        test_scores, train_scores = res["test_scores"], res["train_scores"]
        fit_time, score_time, info, fds = res["fit_time"], res["score_time"], res["info"], res["fds"]
        if res["n_unconverged"]:
            import warnings

            from sklearn.exceptions import ConvergenceWarning

            warnings.warn(f"{res['n_unconverged']} (candidate, fold) problems did not reach the gap tolerance",
                          ConvergenceWarning)

        # ---- cv_results_ in sklearn's layout ------------------------------------------
        results = {}

        def _store(name, arr, splits_=False, rank=False):
            if splits_:
                for i in range(n_splits):
                    results[f"split{i}_{name}"] = arr[:, i]
            mean = arr.mean(axis=1)
            results[f"mean_{name}"] = mean
            results[f"std_{name}"] = np.sqrt(((arr - mean[:, None]) ** 2).mean(axis=1))
            if rank:
                if np.isnan(mean).all():
                    ranks = np.ones_like(mean, dtype=np.int32)
                else:
                    m = np.nan_to_num(mean, nan=np.nanmin(mean) - 1)
                    ranks = rankdata(-m, method="min").astype(np.int32, copy=False)
                results[f"rank_{name}"] = ranks

        _store("fit_time", np.tile(fit_time[:, None], (1, n_splits)))
        _store("score_time", np.tile(score_time[:, None], (1, n_splits)))
        names = sorted({k for c in candidates for k in c})
        for name in names:
            vals = [c.get(name, None) for c in candidates]
            try:
                arr = np.array(vals)
                if arr.ndim != 1 or arr.dtype.kind not in "fiub":
                    raise ValueError
                ma = MaskedArray(arr, mask=[name not in c for c in candidates])
            except Exception:
                ma = MaskedArray(np.empty(n_cand, dtype=object), mask=[name not in c for c in candidates])
                for i, v in enumerate(vals):
                    ma[i] = v
            results[f"param_{name}"] = ma
        results["params"] = candidates
        for m in (scoring if multi else ["score"]):
            _store(f"test_{m}", test_scores[m] if multi else test_scores, splits_=True, rank=True)
            if train_scores is not None:
                _store(f"train_{m}", train_scores[m] if multi else train_scores, splits_=True)

        self.multimetric_ = multi
        refit_metric = self.refit if multi else "score"
        if self.refit or not self.multimetric_:
            self.best_index_ = int(self._select_best_index(self.refit, refit_metric, results))
            self.best_score_ = results[f"mean_test_{refit_metric}"][self.best_index_]
            self.best_score_std_ = results[f"std_test_{refit_metric}"][self.best_index_]
            self.best_params_ = results["params"][self.best_index_]
        if self.refit:
            bi = self.best_index_
            self.best_estimator_ = clone(base).set_params(**clone(self.best_params_, safe=False))
            spec = specs[bi]
            fkey = _fold_key(ests[bi], spec)
            t0 = time.time()
            self.best_estimator_.n_features_in_ = p
            self.best_estimator_._fit_prepared(engine, fds[fkey], spec, dict(opts), B0=res["warm"].get(bi))
            self.refit_time_ = time.time() - t0
        from sklearn.metrics import check_scoring, get_scorer

        self.scorer_ = {m: get_scorer(v) for m, v in scoring.items()} if multi else \
            check_scoring(base, scoring=self.scoring)
        self.cv_results_ = results
        self.n_splits_ = n_splits
        self.solver_info_ = info
        self.batched_ = True
        return self


class LineSearchCV(BaseSearchCV):
    """Coordinate-wise line search over several hyper-parameters
    (reference model_selection.py:427-703): ``n_iter`` successive 1-D grid searches,
    parameter ``i % n_params`` swept while the others stay at their last best value
    (initially the first value of their grid).

    Args:
        estimator: estimator object.
        param_grid (list[tuple[str, list]]): ordered (name, values) pairs.
        opt_selection_method (str | list[str] | None): selection rule per parameter.
        n_iter (int | None): number of line searches; default ``2 * n_params``.
        (remaining arguments as GridSearchCV)
    """

    def __init__(self, estimator, param_grid, *, opt_selection_method=None, n_iter=None,
                 scoring="neg_root_mean_squared_error", n_jobs=None, refit=True, cv=None, verbose=0,
                 pre_dispatch="2*n_jobs", error_score=np.nan, return_train_score=False):
        super().__init__(estimator=estimator, scoring=scoring, n_jobs=n_jobs, refit=refit, cv=cv, verbose=verbose,
                         pre_dispatch=pre_dispatch, error_score=error_score, return_train_score=return_train_score)
        self.param_grid = param_grid
        self.opt_selection_method = opt_selection_method
        self.n_iter = n_iter

    def fit(self, X, y=None, *, groups=None, **fit_params):
        grid = self.param_grid
        if not (isinstance(grid, (list, tuple)) and len(grid) > 0 and isinstance(grid[0], (tuple, list))
                and isinstance(grid[0][0], str)):
            raise ValueError("Parameter grid is not given in the correct format!")
        n_params = len(grid)
        if self.opt_selection_method is None:
            methods = ["max_score"] * n_params
        elif isinstance(self.opt_selection_method, str):
            methods = [self.opt_selection_method] * n_params
        elif (isinstance(self.opt_selection_method, (list, tuple))
              and all(isinstance(m, str) for m in self.opt_selection_method)
              and len(self.opt_selection_method) == n_params):
            methods = list(self.opt_selection_method)
        else:
            raise ValueError(
                "Optimal hyperparams selection methods should be given as a single string, or as a list of "
                "strings with the same amount of parameters!")
        n_iter = self.n_iter if (self.n_iter is not None and self.n_iter > 0) else 2 * n_params

        history = []
        fd_cache = {}  # device-resident design shared by every line (same X, y, folds)
        best = None
        if isinstance(self.estimator, EngineRegressor) and y is not None and not hasattr(X, "columns"):
            # one float64 conversion for all lines: the FoldData cache recognises the design by identity
            try:
                X, y = check_X_y(X, y, dtype=np.float64, y_numeric=True, ensure_min_samples=2, ensure_all_finite=False)
            except Exception:
                pass  # the first GridSearchCV raises the proper error
        for i in range(n_iter):
            pid = i % n_params
            last = [values[0] for _, values in grid] if best is None else [best[name] for name, _ in grid]
            line = {name: (values if j == pid else [last[j]]) for j, (name, values) in enumerate(grid)}
            gs = GridSearchCV(estimator=self.estimator, param_grid=line, opt_selection_method=methods[pid],
                              scoring=self.scoring, n_jobs=self.n_jobs, refit=self.refit, cv=self.cv,
                              verbose=self.verbose, pre_dispatch=self.pre_dispatch, error_score=self.error_score,
                              return_train_score=self.return_train_score)
            gs._fd_cache = fd_cache
            if getattr(self, "_shard", None) is not None:
                gs._shard = self._shard
            if groups is not None:
                fit_params = dict(fit_params, groups=groups)
            gs.fit(X, y, **fit_params)
            best = deepcopy(gs.best_params_)
            history.append(gs)
        for attr in [v for v in vars(history[-1]) if v.endswith("_") and not v.startswith("__")]:
            setattr(self, attr, getattr(history[-1], attr))
        self.history_ = history
        return self

    def _run_search(self, evaluate_candidates):
        """Unused: the search is driven by ``fit``."""
        return

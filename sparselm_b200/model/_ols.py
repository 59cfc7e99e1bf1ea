"""Ordinary least squares on the engine (reference: src/sparselm/model/_ols.py).

The reference minimises ``1/(2n) ||X b - y||^2`` through cvxpy (_ols.py:57-65).  Here the
design goes through the same pack + FP64 tensor-core Gram build as every other estimator
and the normal equations ``G b = c`` are solved by conjugate gradients whose products run on
the tensor-core Gram apply (``slm_gram_cg``).  For an under-determined design the
minimum-norm least-squares solution is returned (cvxpy returns *a* minimiser).
"""

from __future__ import annotations

import warnings

import numpy as np
from sklearn.utils.validation import validate_data

from ..engine import get_engine
from ._base import EngineRegressor, ProblemSpec


class OrdinaryLeastSquares(EngineRegressor):
    r"""Ordinary least squares: ``min_b ||X b - y||_2^2`` (reference _ols.py:16-65).

    Args:
        fit_intercept, copy_X, warm_start, solver, solver_options: see EngineRegressor.
            ``solver_options``: ``tol`` (relative residual ``||c - G b|| / ||c||`` of the normal
            equations, default 1e-13), ``max_iter`` (Gram products), ``device``.
    """

    _batchable = False  # no penalty grid: a CV search over it goes through sklearn's per-fit loop

    def _problem_spec(self, n_features):
        return ProblemSpec(p=n_features, pe=n_features, lam1=0.0,
                           key=("OrdinaryLeastSquares", n_features, bool(self.fit_intercept)))

    def fit(self, X, y, sample_weight=None):
        # finiteness is checked on the device (FoldData.check_finite, raised at the first host sync
        # of the solve) instead of a host pass over X: 50 ms for the 640 MB of a 20000 x 4000 design
        X, y = validate_data(self, X, y, accept_sparse=False, y_numeric=True, multi_output=False,
                             dtype=np.float64, ensure_all_finite=False)
        self._validate_hyperparams(X, y)
        opts = self._engine_options()
        engine = get_engine(opts.pop("device", None))
        sw = None
        if sample_weight is not None:
            from sklearn.utils.validation import _check_sample_weight

            sw = _check_sample_weight(sample_weight, X, dtype=np.float64)
        p = X.shape[1]
        fd = engine.prepare(X, y, None, self.fit_intercept, sw)
        tol = float(opts.get("tol", 1e-13))
        fd.check_finite()  # conjugate gradients synchronise at once anyway: validate before them
        X8, iters, rel = engine.gram_cg(fd.G_full, p, tol=tol, max_iter=opts.get("max_iter"))
        self.coef_ = X8[:, 0].cpu().numpy()
        if self.fit_intercept:
            self.intercept_ = float(engine.intercepts(fd.G_full, p, X8, 1)[0].item())
        else:
            self.intercept_ = 0.0
        self.solver_info_ = {"iterations": iters, "relative_residual": rel, "status": int(rel > 4.0 * tol)}
        if self.solver_info_["status"] != 0:
            from sklearn.exceptions import ConvergenceWarning

            warnings.warn(f"OrdinaryLeastSquares: conjugate gradients stopped at relative residual {rel:.3e} "
                          f"after {iters} Gram products; raise solver_options['max_iter']", ConvergenceWarning)
        return self

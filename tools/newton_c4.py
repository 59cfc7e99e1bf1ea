"""C4 (AdaptiveOverlapGroupLasso CV search) with and without the Newton phase: wall time per
GridSearchCV.fit, unconverged columns, agreement of the score tables and of the refit coefficients.
Usage: python tools/newton_c4.py [on|off|both] [n_alphas]"""
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sklearn.base import clone  # noqa: E402

from sparselm_b200.model_selection import GridSearchCV  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "both"
wl = bench.workload("c4")
alphas = wl["alphas"]
if len(sys.argv) > 2:
    alphas = alphas[:: max(1, len(alphas) // int(sys.argv[2]))]
X, y = wl["X"], wl["y"]
out = {}
for tag in (["off", "on"] if mode == "both" else [mode]):
    est = clone(wl["est"])
    est.set_params(solver_options={"tol": 1e-9, "max_iter": 50000, "newton": tag == "on"})
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        t0 = time.perf_counter()
        gs = GridSearchCV(est, {"alpha": list(alphas)}, cv=5).fit(X, y)
        dt = time.perf_counter() - t0
    out[tag] = gs
    print(f"newton {tag}: {dt:.2f} s, best alpha {gs.best_params_['alpha']:.5g}, warnings {len(wlist)}"
          f" {[str(w.message)[:90] for w in wlist[:2]]}", flush=True)
if len(out) == 2:
    a, b = out["off"], out["on"]
    print("max |mean_test_score diff|", np.abs(a.cv_results_["mean_test_score"] - b.cv_results_["mean_test_score"]).max(),
          "scale", np.abs(a.cv_results_["mean_test_score"]).max())
    print("refit coef diff", np.abs(a.best_estimator_.coef_ - b.best_estimator_.coef_).max(), "scale",
          np.abs(a.best_estimator_.coef_).max())

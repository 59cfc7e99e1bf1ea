"""sklearn's estimator conformance checks over every engine-backed estimator, all failures listed
(reference: tests/test_common.py:95-108 runs check_estimator on each estimator).

    python tools/check_estimators.py > gpurun_out/check_estimators.txt
"""
import os
import sys
import traceback
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.utils.estimator_checks import check_estimator  # noqa: E402

import sparselm_b200.model as M  # noqa: E402

NAMES = ["OrdinaryLeastSquares", "Lasso", "GroupLasso", "OverlapGroupLasso", "SparseGroupLasso", "RidgedGroupLasso",
         "AdaptiveLasso", "AdaptiveGroupLasso", "AdaptiveOverlapGroupLasso", "AdaptiveSparseGroupLasso",
         "AdaptiveRidgedGroupLasso"]
for nm in sys.argv[1:] or NAMES:
    est = getattr(M, nm)(fit_intercept=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = check_estimator(est, on_fail=None, on_skip=None)
    bad = [r for r in res if r["status"] not in ("passed", "skipped", "xfail")]
    print(f"== {nm}: {len(res)} checks, {sum(r['status'] == 'passed' for r in res)} passed, "
          f"{sum(r['status'] == 'skipped' for r in res)} skipped, {len(bad)} failed", flush=True)
    for r in bad:
        exc = r.get("exception")
        msg = "".join(traceback.format_exception_only(type(exc), exc)).strip().replace("\n", " | ")[:400] if exc else ""
        print(f"   {r['check_name']}: {r['status']}: {msg}", flush=True)

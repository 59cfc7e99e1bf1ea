"""Generates tests/golden/golden_host.json from the REFERENCE'S OWN python, imported in the
build container from /root/reference (it does not exist on the GPU box, hence the fixture):

  * sparselm.dataset.make_group_regression (dataset.py:15-139) -- X, y, groups, coefs for seeds;
  * sparselm.tools.constrain_coefficients (tools.py:14-103) around a plain least-squares fit,
    and r2_score_to_cv_error (tools.py:106-131).

Both modules are pure numpy / scikit-learn (no cvxpy), so unlike the solve they run here.
The modules are loaded by file path: importing the `sparselm` package would pull in cvxpy.

Run:  python tests/golden/make_golden_host.py
"""

import importlib.util
import json
import os
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/sparselm"


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def lstsq_fit(X, y):
    return np.linalg.lstsq(X, y, rcond=None)[0]


def main():
    ds, tools = _load("dataset"), _load("tools")
    out = {"dataset": [], "constrain": [], "r2_to_cv": []}
    for seed, kw in enumerate([
        dict(n_samples=12, n_groups=4, n_features_per_group=3, n_informative_groups=2),
        dict(n_samples=10, n_groups=3, n_features_per_group=[2, 4, 3], n_informative_groups=2,
             frac_informative_in_group=0.5, shuffle=False),
        dict(n_samples=15, n_groups=5, n_features_per_group=2, n_informative_groups=5, noise=1.5, bias=2.0),
    ]):
        X, y, groups, coefs = ds.make_group_regression(coef=True, random_state=seed, **kw)
        out["dataset"].append({"kwargs": kw, "seed": seed, "X": X.tolist(), "y": y.tolist(),
                               "groups": groups.tolist(), "coefs": coefs.tolist()})
    rng = np.random.default_rng(11)
    for t in range(6):
        X, y = rng.normal(size=(10, 8)), rng.normal(size=10)
        inds = rng.choice(8, size=3, replace=False)
        low = rng.random(3) - 0.5
        high = rng.random(3) + low
        for kw in (dict(high=2, low=0), dict(high=high.tolist(), low=low.tolist()), dict(high=high.tolist()),
                   dict(low=low.tolist())):
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                coefs = tools.constrain_coefficients(inds, **kw)(lstsq_fit)(X, y)
            out["constrain"].append({"X": X.tolist(), "y": y.tolist(), "indices": inds.tolist(), "kwargs": kw,
                                     "coefs": coefs.tolist(), "warned": len(w) > 0})
    for t in range(3):
        y = rng.normal(size=9)
        yp = y + 0.2 * rng.normal(size=9)
        w = rng.random(9) + 0.1
        out["r2_to_cv"].append({"score": 0.8 - 0.1 * t, "y": y.tolist(), "y_pred": yp.tolist(), "weights": w.tolist(),
                                "weighted": float(tools.r2_score_to_cv_error(0.8 - 0.1 * t, y, yp, w)),
                                "unweighted": float(tools.r2_score_to_cv_error(0.8 - 0.1 * t, y, yp))})
    with open(os.path.join(HERE, "golden_host.json"), "w") as f:
        json.dump(out, f)
    print("wrote golden_host.json:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()

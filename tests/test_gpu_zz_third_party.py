"""Engine estimators against the third-party fixtures of the preprocessing / standardize variants
(tests/golden/make_golden_nlp.py, "cases_preprocess"): intercepts, sample weights, standardized group
norms.  north_star tolerances: coefficients 1e-6 * ||b||_inf, support above 1e-6."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_oracle import NLP, _pre_id, nlp_pre_case  # noqa: E402


@pytest.mark.parametrize("case", NLP["cases_preprocess"], ids=_pre_id)
def test_estimators_match_third_party_preprocessing_and_standardize(case):
    import sparselm_b200.model as M

    X, y, sw, kw = nlp_pre_case(case)
    est = getattr(M, case["name"])(solver_options={"tol": 1e-12}, **kw).fit(X, y, sample_weight=sw)
    assert est.solver_info_["status"] == 0
    ref = np.array(case["coef"])
    assert np.abs(est.coef_ - ref).max() <= 1e-6 * np.abs(ref).max()
    assert np.array_equal(np.abs(est.coef_) > 1e-6, np.abs(ref) > 1e-6)
    assert abs(est.intercept_ - case["intercept"]) <= 1e-6 * max(1.0, abs(case["intercept"]))

"""The engine's estimators against fixtures produced by the REFERENCE'S OWN code
(tests/golden/make_golden_reference.py: the reference's unmodified fit -- problem building,
overlap expansion, adaptive loop, fold-back, intercept -- around a KKT-certified solve).
north_star tolerances: coefficients 1e-6 * ||b||_inf, identical support above 1e-6, same number
of adaptive passes."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_reference_fixtures import REF, ref_data, ref_id, ref_kwargs  # noqa: E402


@pytest.mark.parametrize("case", REF["cases"], ids=ref_id)
def test_estimators_match_reference_code_fixtures(case):
    import warnings

    import sparselm_b200.model as M

    X, y, sw = ref_data(case)
    kw = ref_kwargs(case)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)  # groups=None / group_list=None warnings of the reference
        est = getattr(M, case["estimator"])(fit_intercept=case["fit_intercept"], solver_options={"tol": 1e-13},
                                            **kw).fit(X, y, sample_weight=sw)
    assert est.solver_info_["status"] == 0
    ref = np.array(case["coef"])
    scale = max(np.abs(ref).max(), 1e-12)
    assert np.abs(est.coef_ - ref).max() <= 1e-6 * scale
    assert np.array_equal(np.abs(est.coef_) > 1e-6, np.abs(ref) > 1e-6)
    assert abs(est.intercept_ - case["intercept"]) <= 1e-6 * max(1.0, abs(case["intercept"]))
    if case["n_iter"] is not None:
        assert est.n_iter_ == case["n_iter"]

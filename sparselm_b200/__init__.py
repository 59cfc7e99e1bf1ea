"""sparselm_b200 -- B200-native (sm_100a) solver engine behind sparse-lm's
sklearn-compatible estimator API (convex estimators only)."""

__version__ = "0.1.0"

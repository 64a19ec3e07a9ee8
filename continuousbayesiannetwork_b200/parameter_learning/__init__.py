"""Estimator registry (reference cbn/parameter_learning/__init__.py:7-13).  Only the
discrete brute-force estimator is on the B200 hot path; the reference's continuous
estimators (linear/logistic regression, MLP, GP) are out of scope."""
from .brute_force import BruteForce

ESTIMATORS = {
    "brute_force": BruteForce,
}

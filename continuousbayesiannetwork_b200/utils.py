"""Factories keyed by the same config names as the reference (cbn/utils.py:23-38)."""
from typing import Dict

from .base.parameter_learning import BaseParameterLearningEstimator


def choose_probability_estimator(estimator_name: str, config: Dict, **kwargs) -> BaseParameterLearningEstimator:
    from .parameter_learning import ESTIMATORS

    if estimator_name in ESTIMATORS.keys():
        estimator_class = ESTIMATORS[estimator_name](config, **kwargs)
    else:
        raise ValueError(f"Unknown estimator: {estimator_name}")
    return estimator_class


def choose_inference_obj(inference_name: str, config: Dict, **kwargs):
    """The reference's factory is a no-op returning None (cbn/utils.py:35-38); here it builds the engine."""
    from .inference import INFERENCE_OBJS

    if inference_name in INFERENCE_OBJS.keys():
        return INFERENCE_OBJS[inference_name](config, **kwargs)
    raise ValueError(f"Unknown inference object: {inference_name}")

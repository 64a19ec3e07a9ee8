"""Factories keyed by the config names the reference uses (``estimator_name``, ``inference_obj``;
cbn/utils.py:23-38).  An unknown name is a ``ValueError``, as in the reference; its inference factory is a no-op that
returns ``None`` (:35-38), here it builds the engine."""
from typing import Any, Dict, Mapping


def _build(kind: str, registry: Mapping[str, type], name: str, config: Dict[str, Any], kwargs: Dict[str, Any]):
    try:
        cls = registry[name]
    except KeyError:
        raise ValueError(f"Unknown {kind}: {name}") from None
    return cls(config, **kwargs)


def choose_probability_estimator(estimator_name: str, config: Dict[str, Any], **kwargs: Any):
    from .parameter_learning import ESTIMATORS

    return _build("estimator", ESTIMATORS, estimator_name, config, kwargs)


def choose_inference_obj(inference_name: str, config: Dict[str, Any], **kwargs: Any):
    from .inference import INFERENCE_OBJS

    return _build("inference object", INFERENCE_OBJS, inference_name, config, kwargs)

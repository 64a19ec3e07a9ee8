"""Inference engines by config name (``inference_obj: exact``), the registry the reference keeps in
``cbn/inference/__init__.py:3``; ``register`` adds a third-party engine under a new name."""
from typing import Dict, Type

from ..base.inference import BaseInference
from .exact import ExactInference

INFERENCE_OBJS: Dict[str, Type[BaseInference]] = {"exact": ExactInference}


def register(name: str, cls: Type[BaseInference]) -> None:
    if not (isinstance(cls, type) and issubclass(cls, BaseInference)):
        raise TypeError("an inference engine must subclass BaseInference")
    INFERENCE_OBJS[name] = cls


__all__ = ["ExactInference", "INFERENCE_OBJS", "register"]

"""Inference plugin contract.

Drop-in for the reference's ``BaseInference`` (``cbn/base/inference.py:7-23``): the same class name, constructor
keywords and hooks (``_setup_model``, ``_infer``).  One deliberate difference: ``infer`` hands back what ``_infer``
computes -- the reference discards it (:18-19), which is why its plugin slot could never answer a query.
"""
from __future__ import annotations

import abc
from typing import Any, Dict, Optional, Sequence

__all__ = ["BaseInference"]


class BaseInference(abc.ABC):
    """An inference engine bound to one network; ``infer`` answers ``P(target | evidence)`` for a batch of rows."""

    def __init__(self, config: Dict[str, Any], **kwargs: Any):
        self.device = kwargs.get("device", "cuda")
        self.if_log: bool = bool(kwargs.get("log", False))
        if not str(self.device).startswith("cuda"):
            raise RuntimeError(f"{type(self).__name__} runs on a CUDA device only (got device={self.device!r})")

    @abc.abstractmethod
    def _setup_model(self, config: Dict[str, Any], **kwargs: Any) -> None:
        ...

    @abc.abstractmethod
    def _infer(self, target_node: str, evidence: Dict[str, Any], do: Optional[Sequence[str]], **kwargs: Any):
        ...

    def infer(self, target_node: str, evidence: Dict[str, Any], do: Optional[Sequence[str]] = None, **kwargs: Any):
        """Posterior of ``target_node`` per evidence row (``do``: evidence variables that are interventions)."""
        return self._infer(target_node, evidence or {}, do, **kwargs)

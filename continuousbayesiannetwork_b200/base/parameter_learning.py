"""Estimator plugin contract of the discrete path.

Drop-in for the reference's plugin base class (``cbn/base/parameter_learning.py:7-61``): the same class name, the same
public calls (``fit``, ``get_prob``, ``sample``, ``save_model``, ``load_model``) with the same argument meaning, and the
same four hooks a concrete estimator overrides (``_setup_model``, ``_fit``, ``_get_prob``, ``_sample``).  What this
engine adds on top of the contract: tensors are checked for the shapes the docstrings promise before a hook sees them,
and the estimator refuses a non-CUDA device instead of silently running elsewhere (there is no CPU path).
"""
from __future__ import annotations

import abc
from typing import Any, Dict, Optional

import torch

__all__ = ["BaseParameterLearningEstimator"]


class BaseParameterLearningEstimator(abc.ABC):
    """Learns ``P(node | parents)`` from samples and evaluates it.

    Shapes (as in the reference): ``node_data`` is ``[n]``, ``parents_data`` is ``[P, n]`` or ``None``;
    ``point_to_evaluate`` is ``[n_queries, V]`` and ``query`` is ``[n_queries, P, 1]`` or ``None``;
    ``get_prob`` returns ``[n_queries, V]``.
    """

    # ------------------------------------------------------------------ construction
    def __init__(self, config: Dict[str, Any], **kwargs: Any):
        config = config or {}
        self.estimator_name: Optional[str] = config.get("estimator_name")
        self.device = kwargs.get("device", "cuda")
        self.if_log: bool = bool(kwargs.get("log", False))
        if not str(self.device).startswith("cuda"):
            raise RuntimeError(f"{type(self).__name__} runs on a CUDA device only (got device={self.device!r})")

    def __repr__(self) -> str:
        return f"{type(self).__name__}(estimator_name={self.estimator_name!r}, device={self.device!r})"

    # ------------------------------------------------------------------ hooks of a concrete estimator
    @abc.abstractmethod
    def _setup_model(self, config: Dict[str, Any], **kwargs: Any) -> None:
        ...

    @abc.abstractmethod
    def _fit(self, node_data: torch.Tensor, parents_data: Optional[torch.Tensor] = None) -> None:
        ...

    @abc.abstractmethod
    def _get_prob(self, point_to_evaluate: torch.Tensor, query: Optional[torch.Tensor] = None) -> torch.Tensor:
        ...

    @abc.abstractmethod
    def _sample(self, N: int, **kwargs: Any) -> torch.Tensor:
        ...

    # ------------------------------------------------------------------ public calls
    def fit(self, node_data: torch.Tensor, parents_data: Optional[torch.Tensor] = None) -> None:
        """Learn from ``node_data [n]`` and ``parents_data [P, n]`` (``None`` for a root node)."""
        if node_data.dim() != 1:
            raise ValueError(f"node_data must be [n_samples], got {tuple(node_data.shape)}")
        if parents_data is not None and (parents_data.dim() != 2 or parents_data.shape[1] != node_data.shape[0]):
            raise ValueError(f"parents_data must be [n_parents, {node_data.shape[0]}], got {tuple(parents_data.shape)}")
        self._fit(node_data, parents_data)

    def get_prob(self, point_to_evaluate: torch.Tensor, query: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``P(node = point_to_evaluate[q, v] | parents = query[q])`` as ``[n_queries, V]``."""
        return self._get_prob(point_to_evaluate, query)

    def sample(self, N: int, **kwargs: Any) -> torch.Tensor:
        """``N`` draws from the learned table."""
        return self._sample(int(N), **kwargs)

    def save_model(self, path: str) -> None:  # concrete estimators that persist override both
        raise NotImplementedError(f"{type(self).__name__} does not implement save_model")

    def load_model(self, path: str) -> None:
        raise NotImplementedError(f"{type(self).__name__} does not implement load_model")

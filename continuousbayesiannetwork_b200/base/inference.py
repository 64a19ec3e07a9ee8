"""Inference plugin contract -- the reference's ``BaseInference``
(cbn/base/inference.py:7-23), except that ``infer`` returns what ``_infer`` computes
(the reference drops the return value at :18-19, so its slot could never be used)."""
from abc import ABC, abstractmethod
from typing import Dict


class BaseInference(ABC):
    def __init__(self, config: Dict, **kwargs):
        self.device = kwargs.get("device", "cuda")
        self.if_log = kwargs.get("log", False)

    @abstractmethod
    def _setup_model(self, config: Dict, **kwargs):
        raise NotImplementedError

    def infer(self, target_node: str, evidence: Dict, do: Dict = None, **kwargs):
        return self._infer(target_node, evidence, do, **kwargs)

    @abstractmethod
    def _infer(self, target_node: str, evidence: Dict, do: Dict, **kwargs):
        raise NotImplementedError

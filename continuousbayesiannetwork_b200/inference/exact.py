"""Exact inference by Variable Elimination -- the first real occupant of the reference's
empty plugin slot (cbn/inference/exact.py:6-17, registered in cbn/inference/__init__.py:3
and selected by the ``inference_obj: exact`` config key, cbn/conf/inference/exact.yaml:1)."""
from typing import Dict, Optional, Sequence

import torch

from .. import _native as N
from ..base.inference import BaseInference
from ..ve import FusedPlan, QueryPlan, RowPlan, VECompiler


class ExactInference(BaseInference):
    def __init__(self, config: Dict, **kwargs):
        super(ExactInference, self).__init__(config=config, **kwargs)
        self.normalization = "row"
        self.compiler: Optional[VECompiler] = None
        self._fused: Dict[tuple, FusedPlan] = {}
        self._setup_model(config, **kwargs)

    def _setup_model(self, config: Dict, **kwargs):
        self.normalization = (config or {}).get("normalization", "row")
        if self.normalization not in ("row", "global_max"):
            raise ValueError(f"normalization must be 'row' or 'global_max', got {self.normalization!r}")
        self._budget = {k: int(config[k]) for k in ("table_budget_cells", "merge_budget_cells", "row_temp_floats")
                        if config and k in config}
        for k in ("log_space", "rescale", "profile_compile", "row_unit_layout"):
            if config and k in config:
                self._budget[k] = bool(config[k])

    def bind(self, tables):
        """Attach the fitted network tables (called by BayesianNetwork after every fit)."""
        self.tables = tables
        self.compiler = VECompiler(tables, **self._budget)
        self._fused: Dict[tuple, FusedPlan] = {}

    def plan(self, target_node: str, evidence_names: Sequence[str], do: Sequence[str] = ()) -> QueryPlan:
        assert self.compiler is not None, "inference engine is not bound to a fitted network"
        return self.compiler.compile(target_node, list(evidence_names), do=do)

    def fused_plan(self, targets: Sequence[str], evidence_names: Sequence[str]) -> FusedPlan:
        """One launch for several targets that share the evidence list (same target cardinality)."""
        return FusedPlan([self.plan(t, evidence_names) for t in targets])

    def infer_many(self, targets: Sequence[str], evidence: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Posteriors of SEVERAL targets under the same evidence: the evidence columns are uploaded and encoded once,
        and targets of equal cardinality are answered by one fused launch (``cbn_ve_plan_fuse``).  Not part of the
        reference's surface (its ``infer`` takes one target per call); row-normalised posteriors, float32
        [n_queries, card(target)] on the device, keyed by target name."""
        t = self.tables
        evidence = evidence or {}
        names = [n for n in evidence.keys() if n not in targets]
        for n in list(names) + list(targets):
            if n not in t.index:
                raise ValueError(f"{n} is not a node of the network")
        nq = int(evidence[names[0]].shape[0]) if names else 1
        ld = (max(nq, 1) + 15) // 16 * 16
        codes = torch.empty((max(len(names), 1), ld), dtype=torch.uint8, device=t.device)
        for e, n in enumerate(names):
            c = evidence[n]
            if c.shape[0] != nq:
                raise ValueError("n_queries must be equal for all features.")
            t.encode(c.to(t.device, torch.float32, non_blocking=True).reshape(-1), t.index[n], codes[e])
        out: Dict[str, torch.Tensor] = {}
        by_card: Dict[int, list] = {}
        for tg in targets:
            by_card.setdefault(t.cards[t.index[tg]], []).append(tg)
        for _, group in by_card.items():
            plans = [self.plan(tg, names) for tg in group]
            fusable = len(group) > 1 and all(isinstance(p, QueryPlan) for p in plans)
            for i in range(0, len(group), 8 if fusable else 1):
                chunk = group[i: i + 8]
                if fusable and len(chunk) > 1:
                    key = ("fused", tuple(chunk), tuple(names))
                    fused = self._fused.get(key)
                    if fused is None:
                        fused = self._fused[key] = FusedPlan(plans[i: i + 8])
                    for tg, o in zip(chunk, fused.run_codes(codes, nq)):
                        out[tg] = o
                else:
                    out[chunk[0]] = plans[i].run_codes(codes, nq)
        return out

    def infer_map(self, target_node: str, evidence: Dict[str, torch.Tensor]) -> Optional[torch.Tensor]:
        """MAP value of the target per row through the fused kernel (``cbn_ve_run_f32_map``); ``None`` when the plan is
        not a single-target gather plan with at most 8 target values (the caller then takes posterior + argmax)."""
        t = self.tables
        evidence = evidence or {}
        names = [n for n in evidence.keys() if n != target_node]
        for n in names:
            if n not in t.index:
                raise ValueError(f"evidence variable {n} is not a node of the network")
        plan = self.plan(target_node, names)
        if not isinstance(plan, QueryPlan) or plan.card_t > N.GATHER_MAX_CT or len(names) > N.MAX_EVIDENCE_PTRS:
            return None
        nq = int(evidence[names[0]].shape[0]) if names else 1
        cols = []
        for n in names:
            c = evidence[n]
            if c.shape[0] != nq:
                raise ValueError("n_queries must be equal for all features.")
            cols.append(c.to(t.device, torch.float32, non_blocking=True).reshape(-1).contiguous())
        return plan.run_f32_map(cols, nq)

    def _infer(self, target_node: str, evidence: Dict[str, torch.Tensor], do=None, **kwargs) -> torch.Tensor:
        """Posterior ``P(target | evidence)`` per row: float32 [n_queries, card(target)] on the device.

        evidence: name -> tensor [n_queries, 1] (or [n_queries]) of category VALUES (float), as the
        reference's ``infer`` takes them (bayesian_network.py:208-226)."""
        t = self.tables
        evidence = evidence or {}
        names = [n for n in evidence.keys() if n != target_node]
        for n in names:
            if n not in t.index:
                raise ValueError(f"evidence variable {n} is not a node of the network")
        plan = self.plan(target_node, names, do=do or ())
        if names:
            nq = int(evidence[names[0]].shape[0])
            cols = []
            for n in names:
                c = evidence[n]
                if c.shape[0] != nq:
                    raise ValueError("n_queries must be equal for all features.")
                cols.append(c.to(t.device, torch.float32, non_blocking=True).reshape(-1).contiguous())
        else:
            nq, cols = 1, []
        out = plan.run_f32(cols, nq)
        norm = kwargs.get("normalization", self.normalization)
        if norm == "global_max":
            m = torch.zeros(1, dtype=torch.float32, device=t.device)
            s = N.stream_ptr(t.device)
            N.check(N.lib().cbn_batch_max(t.ctx.handle, out.data_ptr(), out.numel(), m.data_ptr(), s), t.ctx.handle)
            N.check(N.lib().cbn_scale_by_inv(t.ctx.handle, out.data_ptr(), out.numel(), m.data_ptr(), s), t.ctx.handle)
        return out

"""Brute-force discrete CPT estimator on the B200 engine.

Same plugin surface as the reference's ``BruteForce``
(cbn/parameter_learning/brute_force.py): ``fit(node_data[n], parents_data[P, n])``,
``get_prob(points[nq, V], query[nq, P, 1]) -> [nq, V]``, ``sample(N)``, and the
``mle_tensor`` attribute ``[M, P+2]``.  Underneath, the family is integer-coded and
counted by ``cbn_count_run`` (no sort), probabilities come from ``cbn_cpt_from_counts``
and lookups from ``cbn_get_prob_f32`` (no broadcast join).
"""
from typing import Dict, Optional

import torch

from ..base.parameter_learning import BaseParameterLearningEstimator
from ..tables import DiscreteTables

_NODE = "__node__"


class BruteForce(BaseParameterLearningEstimator):
    def __init__(self, config: Dict, **kwargs):
        super(BruteForce, self).__init__(config=config, **kwargs)
        self._tables: Optional[DiscreteTables] = None
        self._name: Optional[str] = None
        self._mle: Optional[torch.Tensor] = None
        self._setup_model(config, **kwargs)

    def _setup_model(self, config: Dict = None, **kwargs):
        pass

    # ------------------------------------------------------------------ fit
    def _fit(self, node_data: torch.Tensor, parents_data: torch.Tensor = None):
        """
        :param node_data: shape [n_samples]
        :param parents_data: shape [n_parents_features, n_samples]
        """
        node_data = node_data.reshape(-1)
        n_par = 0 if parents_data is None else int(parents_data.shape[0])
        if parents_data is not None and parents_data.shape[1] != node_data.shape[0]:
            raise ValueError(
                f"parents_data has {parents_data.shape[1]} samples, node_data has {node_data.shape[0]}")
        names = [f"__parent{i:03d}__" for i in range(n_par)] + [_NODE]
        tables = DiscreteTables(names, {_NODE: names[:-1]}, device=self.device)
        cols = {names[i]: parents_data[i] for i in range(n_par)}
        cols[_NODE] = node_data
        tables.fit_columns(cols)
        self.attach(tables, _NODE)

    def attach(self, tables: DiscreteTables, name: str):
        """Use tables fitted by a network-level fused pass (BayesianNetwork._train)."""
        self._tables = tables
        self._name = name
        self._mle = None

    @property
    def tables(self) -> DiscreteTables:
        assert self._tables is not None, "MLE tensor not fitted yet. Call _fit() first."
        return self._tables

    @property
    def mle_tensor(self) -> Optional[torch.Tensor]:
        """[n_unique, n_parents + 2]: unique rows [parents..., node] + joint probability
        (reference brute_force.py:45-53), materialised on demand from the count table."""
        if self._tables is None:
            return None
        if self._mle is None:
            self._mle = self._tables.mle_tensor(self._name)
        return self._mle

    # ------------------------------------------------------------------ query
    def _get_prob(self, point_to_evaluate: torch.Tensor, query: torch.Tensor = None):
        """
        :param point_to_evaluate: [n_queries, n_values]
        :param query:  [n_queries, n_parents, 1] or None
        :return:       [n_queries, n_values] of conditional probabilities
        """
        assert self._tables is not None, "MLE tensor not fitted yet. Call _fit() first."
        if query is not None:
            assert query.dim() == 3 and query.shape[-1] == 1, \
                f"Query must be [n_queries, n_parents, 1]. Got {query.shape}."
            if point_to_evaluate.shape[0] != query.shape[0]:
                raise ValueError(
                    f"'point_to_evaluate' first dimension must match number of queries. "
                    f"Got {point_to_evaluate.shape[0]}, expected {query.shape[0]}.")
        return self._tables.get_prob(self._name, point_to_evaluate, query)

    def _sample(self, N: int, **kwargs):
        """N draws from the empirical joint of the family (reference brute_force.py:246-265)."""
        mle = self.mle_tensor
        assert mle is not None, "MLE tensor not fitted yet. Call _fit() first."
        indices = torch.multinomial(mle[:, -1], N, replacement=True)
        return mle[indices, :-1]

    # ------------------------------------------------------------------ persistence
    def state_dict(self) -> Dict:
        t = self.tables
        i = t.index[self._name]
        vs = t.family_vars(self._name)
        return {
            "domains": [t.domains[v].cpu() for v in vs],
            "counts": t.counts[t.offsets[i]: t.offsets[i] + t.n_cells[i]].cpu(),
            "n_total": t.n_total,
        }

    def save_model(self, path: str):
        torch.save(self.state_dict(), path)

    def load_model(self, path: str):
        sd = torch.load(path)
        n_par = len(sd["domains"]) - 1
        names = [f"__parent{i:03d}__" for i in range(n_par)] + [_NODE]
        tables = DiscreteTables(names, {_NODE: names[:-1]}, device=self.device)
        tables.set_domains(sd["domains"])
        i = tables.index[_NODE]
        tables.counts[tables.offsets[i]: tables.offsets[i] + tables.n_cells[i]] = sd["counts"].to(tables.device)
        # the parents' own marginal tables are not needed for get_prob; they stay empty
        tables.n_total = int(sd["n_total"])
        tables.finalize()
        self.attach(tables, _NODE)

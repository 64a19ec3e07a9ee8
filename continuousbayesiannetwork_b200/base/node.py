"""Node / factor builder -- the reference's ``Node`` surface (cbn/base/node.py:17-381,
plotting excluded) on the B200 tables."""
from typing import Dict, List, Tuple

import torch

from . import BASE_MAX_CARDINALITY, KEY_CONTINUOUS, KEY_DISCRETE, KEY_MAX_CARDINALITY_FOR_DISCRETE
from ..utils import choose_probability_estimator


class Node:
    def __init__(self, node_name: str, estimator_name: str, parameter_learning_config: Dict,
                 parents_names: List[str] = None, **kwargs):
        self.node_name = node_name
        self.parameter_learning_config = parameter_learning_config
        self.parents_names = parents_names if parents_names else []
        self.device = kwargs.get("device", "cuda")
        self.max_cardinality_for_discrete_domain = kwargs.get(KEY_MAX_CARDINALITY_FOR_DISCRETE, BASE_MAX_CARDINALITY)
        self.plot_prob = kwargs.get("plot_prob", False)
        self.fixed_dtype = kwargs.get("fixed_dtype", torch.float32)
        self.estimator = choose_probability_estimator(estimator_name, parameter_learning_config, **kwargs)
        self.info = {}

    # ------------------------------------------------------------------ fit
    def fit(self, node_data: torch.Tensor, parents_data: torch.Tensor = None, **kwargs):
        """
        :param node_data: shape [n_samples]
        :param parents_data: shape [n_parents_features, n_samples]
        """
        if len(self.parents_names) > 0:
            if parents_data is not None:
                if len(self.parents_names) != parents_data.shape[0]:
                    raise ValueError(
                        f"number of parents features in input ({parents_data.shape[0]}) is not equal to number of parents node set ({len(self.parents_names)})")
                # parents are kept sorted by name, rows permuted accordingly (reference node.py:63-73)
                start = self.parents_names
                self.parents_names = sorted(self.parents_names)
                parents_data = parents_data[[start.index(v) for v in self.parents_names]]
            else:
                raise ValueError(
                    f"parents data is empty; should be [{node_data.shape[0], len(self.parents_names)}]")
        else:
            if parents_data is not None:
                raise ValueError("there are no parents for which setting data.")
        self.estimator.fit(node_data, parents_data)
        t = self.estimator.tables
        doms = [t.domains[v] for v in t.family_vars(self.estimator._name)]
        self._set_info(self.parents_names + [self.node_name], doms)

    def attach(self, tables, name: str):
        """Adopt tables fitted by the network-level fused pass."""
        self.parents_names = sorted(self.parents_names)
        self.estimator.attach(tables, name)
        vs = tables.family_vars(name)
        self._set_info([tables.names[v] for v in vs], [tables.domains[v] for v in vs])

    def _set_info(self, names, domains):
        # [min, max, kind, sorted unique values] per variable (reference node.py:85-110; `kind` is never read)
        self.info = {}
        for n, d in zip(names, domains):
            kind = KEY_CONTINUOUS if d.numel() > self.max_cardinality_for_discrete_domain else KEY_DISCRETE
            self.info[n] = [d[0], d[-1], kind, d]

    def sample(self, N: int, **kwargs) -> torch.Tensor:
        return self.estimator.sample(N, **kwargs)

    # ------------------------------------------------------------------ query
    def get_prob(self, query: Dict[str, torch.Tensor], N: int = 1024) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """
        :param query: dict of torch.Tensors, each with shape [n_queries, 1]
        :param N: number of samples if evidence is not provided
        :return: pdf [n_queries, d_1..d_P, n_samples_node], target_node_domains [n_queries, n_samples_node],
                 parents evaluation points [n_queries | 1, n_parents, 1 | N]   (reference node.py:115-204)
        """
        query = dict(query) if query else {}
        if query:
            n_queries = next(iter(query.values())).shape[0]
            for feat, tensor in query.items():
                assert tensor.shape[0] == n_queries, ValueError("n_queries must be equal for all features.")
                assert tensor.dim() == 2, ValueError("Each query tensor must be of dimension 2.")
        else:
            n_queries = 1
        node_query = query.pop(self.node_name, None)
        parents_query, parents_domains = self._setup_parents_query(query, N)
        total = parents_query.shape[2] if parents_query is not None else 0

        if node_query is None:
            target_node_domains = self.sample_domain(self.node_name, N).unsqueeze(0).expand(n_queries, -1)
        else:
            target_node_domains = node_query.to(self.device, self.fixed_dtype)
        n_samples_node = target_node_domains.shape[1]
        parent_dims = [N if total > 1 else 1 for _ in self.parents_names]

        if len(self.parents_names) > 0:
            # one lookup launch for every (query, parent combination): rows = n_queries * total
            pq = parents_query
            if pq.shape[0] != n_queries:
                pq = pq.expand(n_queries, -1, -1)
            flat_q = pq.permute(0, 2, 1).reshape(n_queries * total, len(self.parents_names), 1)
            pts = target_node_domains.unsqueeze(1).expand(n_queries, total, n_samples_node).reshape(-1, n_samples_node)
            pdfs = self.estimator.get_prob(pts.contiguous(), flat_q.contiguous()).view(n_queries, total, n_samples_node)
        else:
            pdfs = self.estimator.get_prob(target_node_domains.contiguous())
        pdfs = pdfs.view(*([n_queries] + parent_dims + [n_samples_node]))
        return pdfs, target_node_domains, parents_domains

    def _setup_parents_query(self, query: Dict[str, torch.Tensor], N: int):
        """[n_queries, n_parents, combinations] evaluation grid (reference node.py:206-284): all parents
        observed -> one combination; otherwise EVERY parent gets N points (observed ones repeated)."""
        query_features = sorted(list(query.keys()))
        P = len(self.parents_names)
        if len(query_features) > 0:
            nq = query[query_features[0]].shape[0]
            assert all(f in self.parents_names for f in query_features), ValueError(
                "You have specified parent features that don't exist")
            if query_features == self.parents_names:
                new_query = torch.stack([query[p].to(self.device, self.fixed_dtype).reshape(nq) for p in self.parents_names],
                                        dim=1).unsqueeze(-1)
                return new_query, new_query
            pts = torch.empty((nq, P, N), device=self.device, dtype=self.fixed_dtype)
            for i, p in enumerate(self.parents_names):
                if p in query_features:
                    pts[:, i, :] = query[p].to(self.device, self.fixed_dtype).expand(-1, N)
                else:
                    pts[:, i, :] = self.sample_domain(p, N).unsqueeze(0).expand(nq, -1)
            return self._batched_meshgrid_combinations(pts), pts
        if P > 0:
            pts = torch.empty((1, P, N), device=self.device, dtype=self.fixed_dtype)
            for i, p in enumerate(self.parents_names):
                pts[0, i, :] = self.sample_domain(p, N)
            return self._batched_meshgrid_combinations(pts), pts
        return None, None

    def sample_domain(self, node: str, N: int = 1024) -> torch.Tensor:
        """N evaluation points of a variable, sorted.  Same contract as the reference's helper (node.py:286-333): a
        rounded-linspace subsample of the fitted domain when N is smaller than it, the domain itself when equal, and the
        domain plus N - card extra points inside [min, max] otherwise.  The extra points are never-observed values (every
        lookup on them is zero, so they do not change any posterior); the reference draws them at random, here they are
        placed deterministically by bisecting the currently widest gap of the domain, so grids are reproducible."""
        lo, hi, _, dom = self.info[node]
        card = int(dom.shape[0])
        if N <= card:
            if N == card:
                return dom
            pick = torch.linspace(0, card - 1, N).round().long().to(dom.device)
            return dom[pick]
        import heapq

        pts = [float(v) for v in dom.tolist()]
        if card == 1 or float(hi) <= float(lo):
            # a single observed value leaves no gap to split: continue above it in unit steps
            pts += [pts[-1] + 1.0 + k for k in range(N - card)]
            return torch.tensor(pts, dtype=dom.dtype, device=dom.device)
        gaps = [(-(b - a), a, b) for a, b in zip(pts, pts[1:])]
        heapq.heapify(gaps)
        have = set(pts)
        while len(pts) < N:
            _, a, b = heapq.heappop(gaps)
            mid = float(torch.tensor(0.5 * (a + b), dtype=dom.dtype))       # round to the column's dtype before comparing
            if mid in have or not (a < mid < b):
                continue                                                    # gap too narrow for this dtype: drop it
            have.add(mid)
            pts.append(mid)
            heapq.heappush(gaps, (-(mid - a), a, mid))
            heapq.heappush(gaps, (-(b - mid), mid, b))
            if not gaps:
                break
        while len(pts) < N:                                                 # every gap exhausted (pathological domains)
            pts.append(max(pts) + 1.0)
        return torch.tensor(sorted(pts), dtype=dom.dtype, device=dom.device)

    def _batched_meshgrid_combinations(self, input_tensor: torch.Tensor) -> torch.Tensor:
        """[n_queries, n_parents, N] -> [n_queries, n_parents, N^n_parents], 'ij' order (reference
        node.py:335-375), built by broadcasting instead of a Python loop over queries."""
        nq, P, N = input_tensor.shape
        outs = []
        for i in range(P):
            shape = [nq] + [1] * P
            shape[1 + i] = N
            outs.append(input_tensor[:, i, :].reshape(shape).expand([nq] + [N] * P).reshape(nq, -1))
        return torch.stack(outs, dim=1)

    def save_node(self, path: str):
        self.estimator.save_model(path)

    def load_node(self, path: str):
        self.estimator.load_model(path)
        t = self.estimator.tables
        doms = [t.domains[v] for v in t.family_vars(self.estimator._name)]
        self._set_info(sorted(self.parents_names) + [self.node_name], doms)

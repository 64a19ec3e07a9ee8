"""CPU oracle for the cbn hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement of the reference algorithm
(Giovannibriglia/ContinuousBayesianNetwork, mounted at /root/reference when
the fixtures were generated).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
there only as the checker or as the timed CPU baseline -- never as a product
path.  Nothing under ``continuousbayesiannetwork_b200/`` imports this file.

Parity pinning: the reference's own tests hold no golden vector for this path
(both test files are dead), so the restatement is pinned against
outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` checks every function below against them).

Three tiers:

* O1  restated reference arithmetic -- ``fit_mle``, ``get_prob``, ``node_domains``,
      ``infer_star`` follow the reference line by line in behaviour (citations
      below) and use the same PyTorch CPU operators the reference uses, because
      that is where the reference's fp32 arithmetic lives.
* O2  textbook Variable Elimination in fp32 (``ve_posterior``) -- the reference
      has no correct multi-layer query path; this is the
      "reference-style PyTorch path" for Asia/Alarm/synthetic DAGs.
* O3  fp64 full enumeration (``enumerate_posterior``) -- ground truth on small
      networks.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

EPS = 1e-10  # cbn/parameter_learning/brute_force.py:240


# --------------------------------------------------------------------------- O1
def fit_mle(node_data: torch.Tensor, parents_data: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sparse joint table of one family.

    Follows ``BruteForce._fit`` (cbn/parameter_learning/brute_force.py:17-53):
    rows ``[pa_1..pa_P, x]`` are de-duplicated with a lexicographic
    ``torch.unique(dim=0)`` and each unique row gets
    ``prob = fp32(count) / fp32(n)``.

    node_data f32[n]; parents_data f32[P, n] (already in sorted-parent order,
    as ``Node.fit`` hands them over, cbn/base/node.py:63-73).
    Returns f32[M, P+2].
    """
    x = node_data.reshape(-1, 1).to(torch.float32)
    if parents_data is not None:
        rows = torch.cat([parents_data.to(torch.float32).T, x], dim=1)
    else:
        rows = x
    uniq, counts = torch.unique(rows, dim=0, return_counts=True)
    probs = counts.to(torch.float32) / counts.sum()
    return torch.cat([uniq, probs.reshape(-1, 1)], dim=1)


def mle_counts(mle: torch.Tensor, n: int) -> np.ndarray:
    """Integer counts behind an ``mle_tensor`` (``round(prob * n)``; exact for n < 2**24)."""
    return np.rint(mle[:, -1].double().numpy() * n).astype(np.int64)


def node_domains(node_data: torch.Tensor, parents_data: Optional[torch.Tensor]) -> List[torch.Tensor]:
    """Sorted unique values per variable, ``[parents..., node]``.

    Follows the domain bookkeeping of ``Node.fit`` (cbn/base/node.py:85-110):
    ``torch.unique`` of the node column and of every parent row.
    """
    out = []
    if parents_data is not None:
        for i in range(parents_data.shape[0]):
            out.append(torch.unique(parents_data[i].to(torch.float32)))
    out.append(torch.unique(node_data.to(torch.float32)))
    return out


def get_prob(
    mle: torch.Tensor,
    point_to_evaluate: torch.Tensor,
    query: Optional[torch.Tensor] = None,
    chunk_rows: int = 4096,
) -> torch.Tensor:
    """``P(x = v | pa = q)`` for V candidate values per row.

    Follows ``BruteForce._get_prob`` (cbn/parameter_learning/brute_force.py:172-244):
    a broadcast equality join of every ``[q, v]`` against the unique rows,
    ``joint / (parent + 1e-10)``; the marginal branch (no parents) is the
    masked-sum loop at :192-201.  The join is chunked over query rows only
    (the reference materialises the whole ``Q*V x M x (P+1)`` boolean), which
    does not change any per-row arithmetic.

    mle f32[M, P+2]; point_to_evaluate f32[Q, V]; query f32[Q, P, 1] or None.
    """
    mle_data = mle[:, :-1]
    mle_probs = mle[:, -1]
    pts = point_to_evaluate.to(torch.float32)
    if query is None:
        node_values = mle_data[:, -1]
        match = pts.reshape(-1, 1) == node_values.reshape(1, -1)          # [Q*V, M]
        pdf = torch.stack([mle_probs[m].sum() if bool(m.any()) else torch.tensor(0.0) for m in match])
        return pdf.reshape(pts.shape).to(torch.float32)

    assert query.dim() == 3 and query.shape[-1] == 1
    nq, npar, _ = query.shape
    assert pts.shape[0] == nq
    nv = pts.shape[1]
    out = torch.empty((nq, nv), dtype=torch.float32)
    for s in range(0, nq, chunk_rows):
        e = min(nq, s + chunk_rows)
        pq = query[s:e].squeeze(-1).to(torch.float32)                    # [q, P]
        full = torch.cat([pq.unsqueeze(1).expand(-1, nv, -1), pts[s:e].unsqueeze(-1)], dim=-1)
        flat = full.reshape(-1, npar + 1)
        jm = (flat[:, None, :] == mle_data[None, :, :]).all(dim=-1)
        joint = (jm * mle_probs).sum(dim=-1)
        pm = (flat[:, None, :-1] == mle_data[None, :, :-1]).all(dim=-1)
        parent = (pm * mle_probs).sum(dim=-1)
        out[s:e] = (joint / (parent + EPS)).reshape(e - s, nv)
    return out


def infer_star(
    mles: Dict[str, torch.Tensor],
    domains: Dict[str, torch.Tensor],
    parents_of_target: Sequence[str],
    target: str,
    evidence: Dict[str, torch.Tensor],
    n_max: int,
    root_ancestors: Sequence[str] = (),
) -> Tuple[torch.Tensor, torch.Tensor]:
    """``BayesianNetwork.infer`` on a depth-1 DAG with every parent of the target observed.

    Follows cbn/base/bayesian_network.py:208-305 restricted to the only shape on
    which the reference is a posterior: factors of the root
    ancestors collapse to the scalar ``mean_v P(root = v)`` (:282-292, they
    cancel at :296), the target's factor is ``P(x | pa = evidence)`` evaluated on
    its ``N_max``-point domain, and the batch is divided by ONE global max (:296).
    ``n_max`` must equal ``card(target)`` (cbn/base/node.py:298-300).
    """
    dom_t = domains[target]
    assert n_max == dom_t.shape[0], "restated only for N_max == card(target)"
    parents = sorted(parents_of_target)
    nq = evidence[parents[0]].shape[0]
    out = torch.ones((nq, n_max), dtype=torch.float32)
    for r in root_ancestors:
        d = domains[r]
        if n_max < d.shape[0]:
            # sub-sampled domain (cbn/base/node.py:291-296)
            idx = torch.linspace(0, d.shape[0] - 1, n_max).round().long()
            pts = d[idx].reshape(1, -1)
        else:
            # N_max > card pads with random zero-probability points (node.py:302-333):
            # they add zeros to the sum, the mean's divisor stays N_max.
            pts = d.reshape(1, -1)
        p = get_prob(mles[r], pts, None)
        out = out * (p.sum(dim=1) / float(n_max))
    q = torch.stack([evidence[p].reshape(nq) for p in parents], dim=1).unsqueeze(-1).to(torch.float32)
    pts = dom_t.reshape(1, -1).expand(nq, -1)
    out = out * get_prob(mles[target], pts, q)
    out = out / out.max()
    return out, pts.clone()


# --------------------------------------------------------------------------- dense CPTs
def dense_counts(codes: np.ndarray, fam_vars: Sequence[int], cards: Sequence[int]) -> np.ndarray:
    """Dense int64 count table of one family from integer codes.

    codes uint8[n_vars, n]; fam_vars = [parents (sorted-name order)..., node];
    table layout = row-major over fam_vars (node fastest), i.e. the
    lexicographic order of ``torch.unique(dim=0)`` (brute_force.py:42).
    """
    shape = [int(cards[v]) for v in fam_vars]
    idx = np.zeros(codes.shape[1], dtype=np.int64)
    for v in fam_vars:
        idx = idx * int(cards[v]) + codes[v].astype(np.int64)
    return np.bincount(idx, minlength=int(np.prod(shape))).astype(np.int64).reshape(shape)


def cpt_from_counts(counts: np.ndarray, n_total: int) -> Tuple[np.ndarray, np.ndarray]:
    """``joint = fp32(c)/fp32(n)`` (brute_force.py:43) and
    ``cond = joint / (sum_x joint + 1e-10)`` (brute_force.py:228-241), both fp32."""
    joint = (counts.astype(np.float32) / np.float32(n_total)).astype(np.float32)
    parent = joint.sum(axis=-1, keepdims=True, dtype=np.float32)
    cond = (joint / (parent + np.float32(EPS))).astype(np.float32)
    return joint, cond


# --------------------------------------------------------------------------- O2 / O3
class DiscreteNet:
    """Plain container: cards[v], parents[v] (ordered as the CPT axes), cpts[v] with
    shape [card(pa_1).. card(pa_P), card(v)]."""

    def __init__(self, cards, parents, cpts):
        self.cards = [int(c) for c in cards]
        self.parents = [list(p) for p in parents]
        self.cpts = [np.asarray(c) for c in cpts]
        self.n = len(self.cards)

    def ancestors(self, seeds):
        seen, stack = set(), list(seeds)
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            stack.extend(self.parents[v])
        return seen


def _einsum(ops, subs, out_sub, use_log):
    """einsum over integer axis labels, relabelled locally (<= 52 distinct labels)."""
    labels = sorted({l for s in subs for l in s} | set(out_sub))
    m = {l: i for i, l in enumerate(labels)}
    args = []
    for o, s in zip(ops, subs):
        args += [o, [m[l] for l in s]]
    args.append([m[l] for l in out_sub])
    return torch.einsum(*args)


def ve_posterior(
    net: DiscreteNet,
    target: int,
    evidence_vars: Sequence[int],
    evidence_codes: np.ndarray,
    dtype=torch.float32,
    log_space: bool = False,
) -> np.ndarray:
    """Textbook batched Variable Elimination, ``P(target | evidence)`` per row.

    evidence_codes int[n_rows, len(evidence_vars)].  Every CPT that mentions an
    evidence variable is sliced per row (gaining a leading row axis ``-1``),
    hidden variables are summed out one at a time in min-size order, the
    remaining factors are multiplied and the result normalised over the target.
    Rows whose evidence has probability zero return zeros (the reference's
    convention for unseen parent configurations, brute_force.py:240-241).
    ``log_space`` runs products as sums and sum-outs as logsumexp.
    """
    ROW = -1
    ev = {int(v): torch.as_tensor(np.asarray(evidence_codes)[:, i].astype(np.int64)) for i, v in enumerate(evidence_vars)}
    n_rows = int(np.asarray(evidence_codes).shape[0])
    keep = net.ancestors([target] + list(ev))
    factors: List[Tuple[List[int], torch.Tensor]] = []
    for v in sorted(keep):
        scope = net.parents[v] + [v]
        t = torch.as_tensor(np.asarray(net.cpts[v])).to(dtype)
        if log_space:
            t = torch.log(t)
        has_row = False
        # index evidence axes per row
        idx, new_scope = [], []
        ev_axes = [a for a in scope if a in ev]
        if ev_axes:
            # move evidence axes to the front, then advanced-index them with the row codes
            perm = [scope.index(a) for a in ev_axes] + [i for i, a in enumerate(scope) if a not in ev]
            t = t.permute(perm)
            t = t[tuple(ev[a] for a in ev_axes)]  # -> [n_rows, rest...]
            new_scope = [ROW] + [a for a in scope if a not in ev]
            has_row = True
        else:
            new_scope = list(scope)
        factors.append((new_scope, t))
    hidden = [v for v in keep if v not in ev and v != target]

    def size_after(v):
        s = set()
        for sc, _ in factors:
            if v in sc:
                s |= set(sc)
        s.discard(v); s.discard(ROW)
        return int(np.prod([net.cards[a] for a in s])) if s else 1

    while hidden:
        v = min(hidden, key=size_after)
        hidden.remove(v)
        touching = [(sc, t) for sc, t in factors if v in sc]
        factors = [(sc, t) for sc, t in factors if v not in sc]
        out_scope = []
        for sc, _ in touching:
            for a in sc:
                if a != v and a not in out_scope:
                    out_scope.append(a)
        out_scope.sort(key=lambda a: (a != ROW, a))
        if log_space:
            full = out_scope + [v]
            acc = None
            for sc, t in touching:
                # broadcast-add
                shape = [t.shape[sc.index(a)] if a in sc else 1 for a in full]
                tt = t.permute([sc.index(a) for a in full if a in sc]).reshape(shape)
                acc = tt if acc is None else acc + tt
            res = torch.logsumexp(acc, dim=-1)
            res = torch.where(torch.isnan(res), torch.full_like(res, -float("inf")), res)
        else:
            res = _einsum([t for _, t in touching], [sc for sc, _ in touching], out_scope, False)
        factors.append((out_scope, res))
    # multiply what is left: scopes subset of {ROW, target}
    out = torch.zeros((n_rows, net.cards[target]), dtype=dtype) if log_space else torch.ones((n_rows, net.cards[target]), dtype=dtype)
    for sc, t in factors:
        shape = [t.shape[sc.index(a)] if a in sc else 1 for a in (ROW, target)]
        tt = t.permute([sc.index(a) for a in (ROW, target) if a in sc]).reshape(shape)
        out = out + tt if log_space else out * tt
    if log_space:
        m = out.max(dim=1, keepdim=True).values
        m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
        out = torch.exp(out - m)
    z = out.sum(dim=1, keepdim=True)
    post = torch.where(z > 0, out / torch.where(z > 0, z, torch.ones_like(z)), torch.zeros_like(out))
    return post.numpy()


def enumerate_posterior(
    net: DiscreteNet, target: int, evidence_vars: Sequence[int], evidence_codes: np.ndarray
) -> np.ndarray:
    """fp64 ground truth by summing the full joint (small networks only)."""
    evidence_vars = [int(v) for v in evidence_vars]
    keep = sorted(net.ancestors([target] + evidence_vars))
    pos = {v: i for i, v in enumerate(keep)}
    cards = [net.cards[v] for v in keep]
    total = int(np.prod(cards))
    assert total <= 1 << 24, "enumeration oracle is for small networks"
    grid = np.indices(cards).reshape(len(keep), -1)           # [k, total]
    joint = np.ones(total, dtype=np.float64)
    for v in keep:
        idx = tuple(grid[pos[p]] for p in net.parents[v]) + (grid[pos[v]],)
        joint *= np.asarray(net.cpts[v], dtype=np.float64)[idx]
    ec = np.asarray(evidence_codes)
    uniq, inv = np.unique(ec, axis=0, return_inverse=True)
    res = np.zeros((uniq.shape[0], net.cards[target]), dtype=np.float64)
    for u, cfg in enumerate(uniq):
        mask = np.ones(total, dtype=bool)
        for v, c in zip(evidence_vars, cfg):
            mask &= grid[pos[v]] == c
        w = np.bincount(grid[pos[target]][mask], weights=joint[mask], minlength=net.cards[target])
        z = w.sum()
        res[u] = w / z if z > 0 else 0.0
    return res[inv.reshape(-1)]

"""Integer-coded entry points of the engine (no pandas / float categories): what the
benchmark configurations and the multi-GPU driver use.  The float-valued, name-keyed
surface of the reference lives in ``base/bayesian_network.py``; both sit on the same
``DiscreteTables`` and kernels."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N
from .inference.exact import ExactInference
from .synth import NetSpec
from .tables import DiscreteTables


def tables_from_spec(spec: NetSpec, device="cuda") -> DiscreteTables:
    t = DiscreteTables(spec.names, spec.parents_by_name(), device=device)
    t.set_cards(spec.cards)
    return t


def bind_inference(tables: DiscreteTables, **config) -> ExactInference:
    cfg = {"inference_obj": "exact"}
    cfg.update(config)
    inf = ExactInference(cfg, device=str(tables.device))
    inf.bind(tables)
    return inf


def sample_network(spec: NetSpec, seed: int, first: int, n: int, device="cuda",
                   tables: Optional[DiscreteTables] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Ancestral samples ``first .. first+n`` of the network as uint8 codes [n_vars, ld] on the device.
    Counter-based: any split of the sample range over calls / GPUs gives the same data."""
    t = tables if tables is not None else tables_from_spec(spec, device)
    cdf = torch.zeros(t.total_cells, dtype=torch.float32, device=t.device)
    for i, c in enumerate(spec.cdfs()):
        cdf[t.offsets[i]: t.offsets[i] + t.n_cells[i]] = torch.from_numpy(np.ascontiguousarray(c).reshape(-1)).to(t.device)
    codes = out if out is not None else t.new_code_matrix(n)
    order = (C.c_int32 * spec.n)(*spec.topological_order())
    N.check(N.lib().cbn_sample_forward(t.ctx.handle, spec.n, order, t.fams, cdf.data_ptr(), int(seed), int(first), int(n),
                                       codes.data_ptr(), codes.stride(0), N.stream_ptr(t.device)), t.ctx.handle)
    return codes


def fit_network_from_codes(spec: NetSpec, codes: torch.Tensor, n: int, device="cuda",
                           **inference_config) -> Tuple[DiscreteTables, ExactInference]:
    """Count every family of ``spec`` over ``codes`` and derive the CPTs (single GPU)."""
    t = tables_from_spec(spec, device)
    t.count(codes, n)
    t.finalize()
    return t, bind_inference(t, **inference_config)


def install_cpts(spec: NetSpec, device="cuda", **inference_config) -> Tuple[DiscreteTables, ExactInference]:
    """Use the ground-truth CPTs of ``spec`` (no fitting)."""
    t = tables_from_spec(spec, device)
    t.set_cond_tables(spec.cpts)
    return t, bind_inference(t, **inference_config)

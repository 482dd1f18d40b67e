"""Name-keyed deterministic weights (TEST INFRASTRUCTURE).

Golden fixtures must be reproducible on the GPU box without shipping multi-MB
state_dicts: every tensor of a `state_dict` is refilled from a generator seeded by
crc32(key) ^ seed, so the reference model (in `make_golden.py`), the oracle and the product
modules (in tests) all see bit-identical weights given only key names and shapes.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict

import torch

_KEEP_SUFFIXES = ("offset", "atomic_mass", "eps", "num_batches_tracked", "initial_atomref")


def fill_state_dict(sd: Dict[str, torch.Tensor], seed: int = 1) -> Dict[str, torch.Tensor]:
    out = {}
    for key in sd:
        t = sd[key]
        if key.endswith(_KEEP_SUFFIXES) or not t.is_floating_point():
            out[key] = t.clone()
            continue
        # SchNet registers the filter MLP twice (`schnet.py:141-148,175`): alias keys share values
        canon = key.replace(".conv.nn.", ".mlp.")
        g = torch.Generator().manual_seed((zlib.crc32(canon.encode()) ^ seed) & 0x7FFFFFFF)
        if key.endswith(".W"):  # GaussianFourierProjection: N(0,1)
            v = torch.randn(t.shape, generator=g)
        elif t.dim() >= 2:  # xavier-uniform-like
            bound = math.sqrt(6.0 / (t.size(-1) + t.size(-2)))
            v = (torch.rand(t.shape, generator=g) * 2 - 1) * bound
        elif key.endswith("running_var") or key.endswith(".weight"):  # norm scales / variances
            v = torch.rand(t.shape, generator=g) + 0.5
        else:  # biases, running_mean
            v = (torch.rand(t.shape, generator=g) * 2 - 1) * 0.1
        out[key] = v.to(t.dtype)
    return out

"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np
import torch

from oracle import ddpm_oracle as O

TRANSFORMER_CASES = {
    # name: (ModelConfig kwargs, batch)
    "tiny": (dict(num_layers=1, num_heads=8, num_mlp_layers=1, channels=42), 2),
    "base_c42": (dict(num_layers=6, num_heads=8, num_mlp_layers=2, channels=42), 4),
    "base_c146": (dict(num_layers=6, num_heads=8, num_mlp_layers=2, channels=146), 3),
    "base_c512": (dict(num_layers=2, num_heads=8, num_mlp_layers=2, channels=512), 5),
    "large_c42": (dict(num_layers=8, num_heads=16, num_mlp_layers=3, channels=42), 4),
    "heads4": (dict(num_layers=2, num_heads=4, num_mlp_layers=1, channels=42), 3),     # head dim 32 (mma k-steps = 4)
    "heads32": (dict(num_layers=1, num_heads=32, num_mlp_layers=1, channels=42), 3),   # head dim 4 (SIMT attention)
}


def oracle_kwargs(cfg):
    if cfg.arch == "DenseDDPM":
        return dict(num_layers=cfg.num_layers, mlp_dims=cfg.mlp_dims)
    return dict(num_layers=cfg.num_layers, num_heads=cfg.num_heads, num_mlp_layers=cfg.num_mlp_layers,
                mlp_dims=cfg.mlp_dims)


def params_torch(engine, flat, dtype=torch.float32):
    return {k: torch.from_numpy(v).to(dtype) for k, v in engine.flat_to_dict(flat).items()}


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_inputs(seed, batch, shape):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (batch, *shape)).astype(np.float32)
    t = rng.uniform(0.05, 1.0, (batch,)).astype(np.float32)
    return x, t

"""Parameter names / shapes of the reference score networks and flax-default initialisation, in pure numpy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): lets the CPU oracle and bench.py's reference arm build a model
without importing the product package (so no product shared library is mapped into a reference-arm process).
The order and naming mirror the flax module tree of models/ncsn.py:122-179 / models/shared.py:58-75 as the product
arena names them (tests/test_oracle_layout.py asserts both layouts are identical).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

E, FILM_EMB, FILM_HID = 128, 128, 512


def param_shapes(arch: str = "TransformerDDPM", num_layers: int = 6, num_heads: int = 8, num_mlp_layers: int = 2,
                 mlp_dims: int = 2048, channels: int = 42, **_ignored) -> List[Tuple[str, Tuple[int, ...]]]:
    C, M = channels, mlp_dims
    out: List[Tuple[str, Tuple[int, ...]]] = []
    add = lambda n, *s: out.append((n, tuple(s)))
    if arch in ("TransformerDDPM", "TransformerDDPM4"):
        add("in.kernel", C, E); add("in.bias", E)                      # models/ncsn.py:155
        for l in range(num_layers):                                     # models/ncsn.py:158-168
            p = f"l{l}."
            add(p + "ln1.scale", E); add(p + "ln1.bias", E)
            add(p + "attn.qkv.kernel", E, 3 * E); add(p + "attn.qkv.bias", 3 * E)
            add(p + "attn.out.kernel", E, E); add(p + "attn.out.bias", E)
            add(p + "ln2.scale", E); add(p + "ln2.bias", E)
            add(p + "ffn1.kernel", E, M); add(p + "ffn1.bias", M)
            add(p + "ffn2.kernel", M, E); add(p + "ffn2.bias", E)
        add("post_ln.scale", E); add("post_ln.bias", E)                 # models/ncsn.py:170-171
        add("post.kernel", E, M); add("post.bias", M)
        K = num_mlp_layers
    elif arch in ("DenseDDPM", "DenseNCSN"):                            # (DenseNCSN: models/ncsn.py:83-98, same tree)
        add("in.kernel", C, M); add("in.bias", M)                       # models/ncsn.py:129
        K = num_layers
    else:
        raise ValueError(f"unknown architecture {arch}")
    for k in range(K):                                                  # models/ncsn.py:173-175, shared.py:58-75
        p = f"k{k}."
        add(p + "film.d1.kernel", FILM_EMB, FILM_HID); add(p + "film.d1.bias", FILM_HID)
        add(p + "film.d2.kernel", FILM_HID, FILM_HID); add(p + "film.d2.bias", FILM_HID)
        add(p + "film.ss.kernel", FILM_HID, 2 * M); add(p + "film.ss.bias", 2 * M)
        add(p + "res.ln_a.scale", M); add(p + "res.ln_a.bias", M)
        add(p + "res.a.kernel", M, M); add(p + "res.a.bias", M)
        add(p + "res.ln_b.scale", M); add(p + "res.ln_b.bias", M)
        add(p + "res.b.kernel", M, M); add(p + "res.b.bias", M)
    add("out_ln.scale", M); add("out_ln.bias", M)                       # models/ncsn.py:177-178
    add("out.kernel", M, C); add("out.bias", C)
    return out


def num_params(**kw) -> int:
    return int(sum(int(np.prod(s)) for _, s in param_shapes(**kw)))


def init_params(seed: int = 0, perturb: float = 0.0, **kw) -> Dict[str, np.ndarray]:
    """flax.nn 0.3.0 defaults: Dense kernel lecun_normal (truncated normal +-2 sigma, stddev sqrt(1/fan_in)/0.8796...),
    bias zeros; LayerNorm scale ones / bias zeros.  Same draw order as the product's Engine.init_params, so identical
    seeds give identical arrays."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in param_shapes(**kw):
        n = int(np.prod(shape))
        if name.endswith(".kernel"):
            std = np.sqrt(1.0 / shape[0]) / 0.87962566103423978
            v = rng.standard_normal(n * 2)
            v = v[np.abs(v) <= 2.0][:n]
            while v.size < n:  # pragma: no cover
                extra = rng.standard_normal(n)
                v = np.concatenate([v, extra[np.abs(extra) <= 2.0]])[:n]
            val = (v * std).astype(np.float32)
        elif name.endswith(".scale"):
            val = np.ones((n,), np.float32)
        else:
            val = np.zeros((n,), np.float32)
        if perturb and not name.endswith(".kernel"):
            val = val + rng.normal(0.0, perturb, n).astype(np.float32)
        out[name] = val.reshape(shape)
    return out


def flops_fwd_per_sample(arch: str = "TransformerDDPM", num_layers: int = 6, num_heads: int = 8, num_mlp_layers: int = 2,
                         mlp_dims: int = 2048, channels: int = 42, seq_len: int = 32, **_ignored) -> float:
    """Algorithmic forward FLOPs per sample."""
    S, C, M = seq_len, channels, mlp_dims
    if arch in ("DenseDDPM", "DenseNCSN"):
        K = num_layers
        mac_tok, S = C * M + K * 2 * M * M + M * C, 1
    else:
        K = num_mlp_layers
        mac_tok = C * E + num_layers * (4 * E * E + 2 * S * E + 2 * E * M) + E * M + K * 2 * M * M + M * C
    mac_film = K * (128 * 512 + 512 * 512 + 2 * 512 * M)
    return 2.0 * (S * mac_tok + mac_film)

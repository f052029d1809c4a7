"""CPU restatement (torch, fp32 or fp64) of the reference DDPM hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED by the
reference (no golden vectors exist upstream); pinned by KATs in tests/.

Each function cites the reference file:line it restates (paths relative to
the upstream tree magenta/symbolic-music-diffusion @ 469204d).

Parameters are a flat ``dict[str, Tensor]`` using the arena names of
``smd_b200.params`` (weights are ``(in, out)`` row-major exactly as
``flax.nn.Dense`` stores them: ``y = x @ W + b``).

``emulate_bf16=True`` rounds both GEMM operands to bfloat16 at exactly the
points where the CUDA path feeds the tensor cores (fp32 accumulate), so the
kernels can be checked tightly (logic) as well as against true fp32 (accuracy).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# third-party primitives (flax.nn 0.3.0 / jax.nn) restated -- SURVEY Appendix B.1
# ----------------------------------------------------------------------------
def _q(x: Tensor, emulate_bf16: bool) -> Tensor:
    if not emulate_bf16:
        return x
    return x.to(torch.bfloat16).to(x.dtype)


def dense(x: Tensor, kernel: Tensor, bias: Tensor, emulate_bf16: bool = False) -> Tensor:
    """flax.nn.Dense: y = x @ kernel + bias (kernel is (in, out))."""
    return _q(x, emulate_bf16) @ _q(kernel, emulate_bf16) + bias


def layer_norm(x: Tensor, scale: Tensor, bias: Tensor, eps: float = 1e-6) -> Tensor:
    """flax.nn.LayerNorm 0.3.0: var = E[x^2] - E[x]^2, y = (x-mean)*rsqrt(var+eps)*scale+bias."""
    mean = x.mean(dim=-1, keepdim=True)
    mean2 = (x * x).mean(dim=-1, keepdim=True)
    var = mean2 - mean * mean
    mul = torch.rsqrt(var + eps) * scale
    return (x - mean) * mul + bias


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def gelu_tanh(x: Tensor) -> Tensor:
    """jax.nn.gelu(approximate=True) as used by flax.nn.gelu 0.3.0."""
    c = math.sqrt(2.0 / math.pi)
    return 0.5 * x * (1.0 + torch.tanh(c * (x + 0.044715 * x * x * x)))


def self_attention(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int,
                   emulate_bf16: bool = False) -> Tensor:
    """flax.nn.SelfAttention (MultiHeadDotProductAttention, inputs_kv = inputs_q).

    Called at models/ncsn.py:161.  q is scaled by 1/sqrt(depth) BEFORE the dot;
    weights = exp(logits - logsumexp(logits)); no mask, dropout 0.
    qkv.kernel is (E, 3E) = [q | k | v] columns, head-major inside each E block.
    """
    B, S, E = x.shape
    dh = E // num_heads
    qkv = dense(x, p[prefix + "qkv.kernel"], p[prefix + "qkv.bias"], emulate_bf16)
    q, k, v = qkv.split(E, dim=-1)
    q = q.reshape(B, S, num_heads, dh) / math.sqrt(dh)
    k = k.reshape(B, S, num_heads, dh)
    v = v.reshape(B, S, num_heads, dh)
    logits = torch.einsum("bqhd,bkhd->bhqk", q, k)
    w = torch.exp(logits - torch.logsumexp(logits, dim=-1, keepdim=True))
    o = torch.einsum("bhqk,bkhd->bqhd", w, v).reshape(B, S, E)
    return dense(o, p[prefix + "out.kernel"], p[prefix + "out.bias"], emulate_bf16)


# ----------------------------------------------------------------------------
# models/shared.py, models/ncsn.py
# ----------------------------------------------------------------------------
def _sinusoid_freqs(half_dim: int, dtype) -> Tensor:
    # shared.py:41-43 / ncsn.py:33-35: emb = log(10000)/(half_dim-1); exp(arange*-emb)
    # jnp computes this in float32 regardless; keep that rounding for fp32, exact for fp64.
    if dtype == torch.float64:
        emb = math.log(10000.0) / float(half_dim - 1)
        return torch.exp(torch.arange(half_dim, dtype=torch.float64) * -emb)
    emb = np.float32(np.log(np.float32(10000.0))) / np.float32(half_dim - 1)
    f = np.exp(np.arange(half_dim).astype(np.float32) * -emb).astype(np.float32)
    return torch.from_numpy(f)


def transformer_positional_encoding(seq_len: int, channels: int, dtype=torch.float32) -> Tensor:
    """models/shared.py:33-48 (TransformerPositionalEncoding)."""
    half = channels // 2
    f = _sinusoid_freqs(half, dtype)
    ts = torch.arange(seq_len, dtype=dtype)
    arg = ts[:, None] * f[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def noise_encoding(noise: Tensor, channels: int) -> Tensor:
    """models/ncsn.py:25-41 (NoiseEncoding). noise: (B,) -> (B, channels).

    Product order matters in fp32: (5000 * noise)[:, None] * freq[None, :].
    """
    half = channels // 2
    f = _sinusoid_freqs(half, noise.dtype)
    arg = (5000 * noise)[:, None] * f[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def dense_film(t: Tensor, p: Dict[str, Tensor], prefix: str, emb_ch: int = 128):
    """models/ncsn.py:44-61 (DenseFiLM). t: (B,) -> scale, shift (B, out).

    ss.kernel is (4*emb_ch, 2*out) = [scale | shift] columns (two flax Dense).
    FiLM GEMMs are kept in full precision on both paths (never bf16).
    """
    e = noise_encoding(t, emb_ch)
    e = dense(e, p[prefix + "d1.kernel"], p[prefix + "d1.bias"])
    e = swish(e)
    e = dense(e, p[prefix + "d2.kernel"], p[prefix + "d2.bias"])
    ss = dense(e, p[prefix + "ss.kernel"], p[prefix + "ss.bias"])
    out = ss.shape[-1] // 2
    return ss[:, :out], ss[:, out:]


def dense_res_block(x: Tensor, scale: Tensor, shift: Tensor, p: Dict[str, Tensor], prefix: str,
                    emulate_bf16: bool = False) -> Tensor:
    """models/shared.py:58-75 (DenseResBlock with FeaturewiseAffine :51-55)."""
    o = layer_norm(x, p[prefix + "ln_a.scale"], p[prefix + "ln_a.bias"])
    o = scale * o + shift
    o = swish(o)
    o = dense(o, p[prefix + "a.kernel"], p[prefix + "a.bias"], emulate_bf16)
    if emulate_bf16:
        # the CUDA path stores this intermediate as bf16 (it only feeds the next LayerNorm), but takes the
        # LayerNorm statistics from the fp32 accumulators
        mean = o.mean(dim=-1, keepdim=True)
        var = (o * o).mean(dim=-1, keepdim=True) - mean * mean
        oq = _q(o, True)
        o = (oq - mean) * (torch.rsqrt(var + 1e-6) * p[prefix + "ln_b.scale"]) + p[prefix + "ln_b.bias"]
    else:
        o = layer_norm(o, p[prefix + "ln_b.scale"], p[prefix + "ln_b.bias"])
    o = scale * o + shift
    o = swish(o)
    o = dense(o, p[prefix + "b.kernel"], p[prefix + "b.bias"], emulate_bf16)
    return o + x


def transformer_ddpm(p: Dict[str, Tensor], inputs: Tensor, t: Tensor, num_layers: int = 6,
                     num_heads: int = 8, num_mlp_layers: int = 2, mlp_dims: int = 2048,
                     emulate_bf16: bool = False, trace: Optional[dict] = None) -> Tensor:
    """models/ncsn.py:138-179 (TransformerDDPM.apply). inputs (B,S,C), t (B,1,1) or (B,).

    ``trace`` (tests only) collects the intermediate tensors under the CUDA workspace names."""
    def rec(name, val):
        if trace is not None:
            trace[name] = val.detach().clone()
    B, S, C = inputs.shape
    E = 128
    dt = inputs.dtype
    temb = transformer_positional_encoding(S, E, dt)[None]
    # NB: the C->128 input projection is SIMT fp32 on the CUDA path (never bf16).
    x = dense(inputs, p["in.kernel"], p["in.bias"])
    x = x + temb
    rec("t.h0", x)
    for l in range(num_layers):
        pre = f"l{l}."
        sc = x
        a = layer_norm(x, p[pre + "ln1.scale"], p[pre + "ln1.bias"])
        rec(f"t.a1_{l}", a)
        a = self_attention(a, p, pre + "attn.", num_heads, emulate_bf16)
        x = a + sc
        rec(f"t.h{2 * l + 1}", x)
        sc2 = x
        m = layer_norm(x, p[pre + "ln2.scale"], p[pre + "ln2.bias"])
        rec(f"t.a2_{l}", m)
        m = dense(m, p[pre + "ffn1.kernel"], p[pre + "ffn1.bias"], emulate_bf16)
        rec(f"t.hpre{l}", m)
        m = gelu_tanh(m)
        rec(f"t.hid{l}", m)
        m = dense(m, p[pre + "ffn2.kernel"], p[pre + "ffn2.bias"], emulate_bf16)
        x = m + sc2
        rec(f"t.h{2 * l + 2}", x)
    x = layer_norm(x, p["post_ln.scale"], p["post_ln.bias"])
    rec("t.a_post", x)
    x = dense(x, p["post.kernel"], p["post.bias"], emulate_bf16)
    rec("t.u0", x)
    tt = t.reshape(B)
    for k in range(num_mlp_layers):
        pre = f"k{k}."
        scale, shift = dense_film(tt, p, pre + "film.")
        rec(f"ss{k}", torch.cat([scale, shift], dim=-1))
        x = dense_res_block(x, scale[:, None, :], shift[:, None, :], p, pre + "res.", emulate_bf16)
        rec(f"t.u{k + 1}", x)
    x = layer_norm(x, p["out_ln.scale"], p["out_ln.bias"])
    rec("t.act_out", x)
    x = dense(x, p["out.kernel"], p["out.bias"], emulate_bf16)
    return x


def dense_ddpm(p: Dict[str, Tensor], inputs: Tensor, t: Tensor, num_layers: int = 3,
               mlp_dims: int = 2048, emulate_bf16: bool = False, **_ignored) -> Tensor:
    """models/ncsn.py:122-135 (DenseDDPM.apply). inputs (B,C), t (B,1) or (B,).

    Accepts-and-ignores num_heads / num_mlp_layers.
    """
    B = inputs.shape[0]
    x = dense(inputs, p["in.kernel"], p["in.bias"], emulate_bf16)
    tt = t.reshape(B)
    for k in range(num_layers):
        pre = f"k{k}."
        scale, shift = dense_film(tt, p, pre + "film.")
        x = dense_res_block(x, scale, shift, p, pre + "res.", emulate_bf16)
    x = layer_norm(x, p["out_ln.scale"], p["out_ln.bias"])
    x = dense(x, p["out.kernel"], p["out.bias"], emulate_bf16)
    return x


def dense_ncsn(p: Dict[str, Tensor], inputs: Tensor, sigmas: Tensor, num_layers: int = 3,
               mlp_dims: int = 2048, emulate_bf16: bool = False, **_ignored) -> Tensor:
    """models/ncsn.py:83-98 (DenseNCSN.apply) with its undefined `t` read as `sigmas` (broken as released; SURVEY
    section 0): the DenseDDPM stack conditioned on sigma, output divided by sigma.  inputs (B,C), sigmas (B,1) or (B,)."""
    B = inputs.shape[0]
    return dense_ddpm(p, inputs, sigmas, num_layers=num_layers, mlp_dims=mlp_dims, emulate_bf16=emulate_bf16) / \
        sigmas.reshape(B, 1)


def model_apply(arch: str, p, inputs, t, **kw):
    if arch in ("TransformerDDPM", "TransformerDDPM4"):
        return transformer_ddpm(p, inputs, t, **kw)
    if arch == "DenseDDPM":
        return dense_ddpm(p, inputs, t, **kw)
    if arch == "DenseNCSN":
        return dense_ncsn(p, inputs, t, **kw)
    raise ValueError(f"unknown architecture {arch}")


# ----------------------------------------------------------------------------
# utils/ebm_utils.py:62-86  create_noise_schedule  (fp32 numpy, like jnp)
# ----------------------------------------------------------------------------
def _linspace_f32(start, stop, num: int) -> np.ndarray:
    """jnp.linspace as of jax 0.2.8: start + iota(num) * ((stop - start) / (num - 1)), all float32.

    (From memory of that version -- unverifiable offline; np.linspace differs by <= 1 ulp.)
    """
    f32 = np.float32
    start, stop = f32(start), f32(stop)
    if num == 1:
        return np.asarray([start], f32)
    delta = f32(stop - start) / f32(num - 1)
    return (start + np.arange(num, dtype=f32) * delta).astype(f32)


def create_noise_schedule(sigma_begin=1.0, sigma_end=1e-2, L=10, schedule="geometric") -> np.ndarray:
    if schedule == "geometric":
        s = np.exp(_linspace_f32(np.log(np.float32(sigma_begin)), np.log(np.float32(sigma_end)), L)).astype(np.float32)
    elif schedule == "linear":
        s = _linspace_f32(sigma_begin, sigma_end, L)
    elif schedule == "fibonacci":
        v = [1e-6, 2e-6]
        for _ in range(L - 2):
            v.append(v[-1] + v[-2])
        s = np.asarray(v, dtype=np.float32)
    else:
        raise ValueError(f"Unsupported schedule: {schedule}")
    return s


# ----------------------------------------------------------------------------
# utils/losses.py:22-30, 250-308
# ----------------------------------------------------------------------------
def reduce_fn(x: Tensor, mode: Optional[str]):
    if mode == "none" or mode is None:
        return x
    if mode == "sum":
        return x.sum()
    if mode == "mean":
        return x.mean()
    raise ValueError("Unsupported reduction option.")


def alphas_prod_with_one(betas: np.ndarray) -> np.ndarray:
    """losses.py:277-281: concat([1], cumprod(1 - betas)) in float32."""
    b = np.asarray(betas, dtype=np.float32)
    return np.concatenate([np.ones((1,), np.float32), np.cumprod(np.float32(1.0) - b, dtype=np.float32)])


def uniform_minmax(u01: np.ndarray, minval: np.ndarray, maxval: np.ndarray) -> np.ndarray:
    """jax 0.2.8 random.uniform tail: max(minval, u*(maxval-minval)+minval)  (SURVEY D8)."""
    u01 = np.asarray(u01, np.float32)
    return np.maximum(minval, u01 * (maxval - minval) + minval).astype(np.float32)


def diffusion_loss_tensors(apply_fn, batch: Tensor, used_alphas: Tensor, eps: Tensor,
                           reduction: str = "mean"):
    """losses.py:288-308 given the sampled tensors (tensor-level parity contract).

    used_alphas: (B,) ; eps: batch.shape.  Returns (loss, pred).
    """
    B = batch.shape[0]
    ua = used_alphas.reshape(B, *([1] * (batch.dim() - 1)))
    perturbed = torch.sqrt(ua) * batch + torch.sqrt(1 - ua) * eps
    pred = apply_fn(perturbed, torch.sqrt(ua))
    loss = (eps - pred) ** 2
    loss = loss.mean(dim=tuple(range(1, loss.dim())))
    return reduce_fn(loss, reduction), pred


def diffusion_loss_draws(rng_key, batch_shape, betas, continuous_noise=True):
    """losses.py:270-294 random draws via the jax-0.2.8 threefry restatement.

    Returns labels (B,) int32, used_alphas (B,) f32, eps batch_shape f32.
    """
    from . import threefry as tf
    T = len(betas)
    rng, label_rng, sample_rng = tf.split(rng_key, 3)
    B = batch_shape[0]
    labels = tf.randint(label_rng, (B,), int(continuous_noise), T + int(continuous_noise))
    ap = alphas_prod_with_one(betas)
    rng, noise_rng = tf.split(rng, 2)
    u = tf.uniform01(noise_rng, (B,))
    used = uniform_minmax(u, ap[labels - 1], ap[labels])
    eps = tf.normal(sample_rng, tuple(batch_shape))
    return labels, used, eps


# ----------------------------------------------------------------------------
# utils/losses.py:129-179  denoising_score_matching_loss  (NCSN family)
# ----------------------------------------------------------------------------
def dsm_loss_tensors(apply_fn, batch: Tensor, used_sigmas: Tensor, eps: Tensor, reduction: str = "mean"):
    """losses.py:161-179 given the sampled tensors.  used_sigmas (B,), eps: batch.shape.  Returns (loss, scores)."""
    B = batch.shape[0]
    us = used_sigmas.reshape(B, *([1] * (batch.dim() - 1)))
    noise = eps * us
    perturbed = batch + noise
    target = -1 / (us ** 2) * noise
    scores = apply_fn(perturbed, us)
    loss = 0.5 * ((scores.reshape(B, -1) - target.reshape(B, -1)) ** 2).sum(dim=-1) * us.reshape(B) ** 2
    return reduce_fn(loss, reduction), scores


def dsm_draws(rng_key, batch_shape, sigmas, continuous_noise=False):
    """losses.py:146-164 via the jax-0.2.8 threefry restatement.  Returns labels, used_sigmas (B,), eps (unit normal;
    the reference multiplies it by sigma at :163)."""
    from . import threefry as tf
    sig = np.asarray(sigmas, np.float32)
    rng, label_rng, sample_rng = tf.split(rng_key, 3)
    B = batch_shape[0]
    labels = tf.randint(label_rng, (B,), int(continuous_noise), len(sig))
    if continuous_noise:
        rng, noise_rng = tf.split(rng, 2)
        used = uniform_minmax(tf.uniform01(noise_rng, (B,)), sig[labels - 1], sig[labels])
    else:
        used = sig[labels]
    eps = tf.normal(sample_rng, tuple(batch_shape))
    return labels, used.astype(np.float32), eps


# ----------------------------------------------------------------------------
# utils/ebm_utils.py:89-198, 201-271  annealed / consistent Langevin dynamics
# ----------------------------------------------------------------------------
def _axis1_norm(a: Tensor) -> Tensor:   # ebm_utils.py:166-170: sqrt(sum(a^2, axis=1) + 1e-10).mean()
    return torch.sqrt((a * a).sum(dim=1) + 1e-10).mean()


def annealed_langevin_dynamics(apply_fn, sigmas, init: Tensor, epsilon, T: int, denoise: bool, noise_fn,
                               infill=False, infill_samples=None, infill_masks=None):
    """ebm_utils.py:89-198.  noise_fn(sigma_index, step) -> (z, z_infill).  Returns state, collection
    (101 + denoise, ...), metrics (4, L, T)."""
    sig = np.asarray(sigmas, np.float32)
    L = len(sig)
    dt = init.dtype
    if not infill:
        infill_samples, infill_masks = torch.zeros_like(init), torch.zeros_like(init)
    idx_tab = np.linspace(1, L * T, 100).astype(np.int32)
    collection = torch.zeros((101 + int(bool(denoise)),) + tuple(init.shape), dtype=dt)
    collection[0] = init * (1 - infill_masks) + infill_samples * infill_masks
    mets = torch.zeros((4, L, T), dtype=dt)
    state = init
    for si in range(L):
        sigma = torch.tensor(sig[si], dtype=dt)
        alpha = torch.tensor(np.float32(epsilon) * (sig[si] / sig[-1]) ** 2, dtype=dt)
        for i in range(T):
            z, zi = noise_fn(si, i)
            y = infill_samples + sigma * (zi if zi is not None else torch.zeros_like(init))
            grad = apply_fn(state, sigma * torch.ones((state.shape[0],) + (1,) * (state.dim() - 1), dtype=dt))
            noise = torch.sqrt(2 * alpha) * z
            nxt = state + alpha * grad + noise
            nxt = nxt * (1 - infill_masks) + y * infill_masks
            image_idx = si * T + i + 1
            mask = idx_tab == image_idx
            if mask.any():
                # (with fewer than 100 total steps linspace repeats indices and the summed slot leaves the buffer: XLA's
                # scatter drops such an update)
                slot = int(np.sum(np.arange(len(idx_tab)) * mask) + 1)
                if slot < collection.shape[0]:
                    collection[slot] = nxt
            mets[0, si, i] = _axis1_norm(grad); mets[1, si, i] = _axis1_norm(alpha * grad)
            mets[2, si, i] = alpha; mets[3, si, i] = _axis1_norm(noise)
            state = nxt
    if denoise:
        s_last = torch.tensor(sig[-1], dtype=dt)
        state = state + s_last ** 2 * apply_fn(state, s_last * torch.ones((state.shape[0],) + (1,) * (state.dim() - 1), dtype=dt))
        collection[-1] = state
    return state, collection, mets


def consistent_langevin_dynamics(apply_fn, sigmas, init: Tensor, epsilon, denoise: bool, noise_fn):
    """ebm_utils.py:201-271.  noise_fn(i) -> z.  Returns state, metrics (4, L, 1)."""
    sig = np.asarray(sigmas, np.float32)
    L = len(sig)
    dt = init.dtype
    beta = torch.tensor(np.sqrt(np.float32(1) - (np.float32(1) - np.float32(epsilon) / sig[-1] ** 2) ** 2), dtype=dt)
    mets = torch.zeros((4, L, 1), dtype=dt)
    state = init
    ones = lambda: torch.ones((state.shape[0],) + (1,) * (state.dim() - 1), dtype=dt)
    for i in range(L):
        sigma = torch.tensor(sig[i], dtype=dt)
        nxt_sigma = torch.tensor(sig[i + 1] if i < L - 1 else 0.0, dtype=dt)
        alpha = torch.tensor(np.float32(epsilon) * (sig[i] / sig[-1]) ** 2, dtype=dt)
        grad = apply_fn(state, sigma * ones())
        noise = beta * nxt_sigma * noise_fn(i)
        mets[0, i, 0] = _axis1_norm(grad); mets[1, i, 0] = _axis1_norm(alpha * grad)
        mets[2, i, 0] = alpha; mets[3, i, 0] = _axis1_norm(noise)
        state = state + alpha * grad + noise
    if denoise:
        s_last = torch.tensor(sig[-1], dtype=dt)
        state = state + s_last ** 2 * apply_fn(state, s_last * ones())
    return state, mets


# ----------------------------------------------------------------------------
# utils/ebm_utils.py:274-405  diffusion_dynamics
# ----------------------------------------------------------------------------
def reverse_coefficients(betas: np.ndarray) -> Dict[str, np.ndarray]:
    """Per-step scalars of ebm_utils.py:315-318, 332-357, 363-364 in float32, same op order."""
    f32 = np.float32
    betas = np.asarray(betas, f32)
    alphas = (f32(1) - betas).astype(f32)
    alphas_prod = np.cumprod(alphas, dtype=f32)
    alphas_prod_prev = np.concatenate([np.ones((1,), f32), alphas_prod[:-1]])
    sqrt_recip = np.sqrt(f32(1) / alphas_prod).astype(f32)
    sqrt_m1 = (np.sqrt(f32(1) - alphas_prod) * sqrt_recip).astype(f32)
    mu1 = (betas * np.sqrt(alphas_prod_prev) / (f32(1) - alphas_prod)).astype(f32)
    mu2 = ((f32(1) - alphas_prod_prev) * np.sqrt(alphas) / (f32(1) - alphas_prod)).astype(f32)
    var = (betas * (f32(1) - alphas_prod_prev) / (f32(1) - alphas_prod)).astype(f32)
    log_var = np.log(np.maximum(var, f32(1e-20))).astype(f32)
    sigma = np.exp(f32(0.5) * log_var).astype(f32)
    return dict(beta=betas, alpha=alphas, alpha_prod=alphas_prod, alpha_prod_prev=alphas_prod_prev,
                sqrt_recip_alpha_prod=sqrt_recip, sqrt_alpha_prod_m1=sqrt_m1,
                sqrt_alpha_prod=np.sqrt(alphas_prod).astype(f32),
                sqrt_one_minus_alpha_prod=np.sqrt(f32(1) - alphas_prod).astype(f32),
                mu1=mu1, mu2=mu2, sigma=sigma)


def collection_slots(T: int, collection_steps: int = 40) -> np.ndarray:
    """ebm_utils.py:320-325, 387-394: slot written after the step that handles index t (or -1)."""
    idx_tab = np.linspace(1, T, collection_steps).astype(np.int32)
    slots = np.full((T,), -1, np.int32)
    for t in range(T):
        image_idx = T - t + 1
        mask = np.isin(idx_tab, image_idx)
        if mask.any():
            slots[t] = int(np.sum(np.arange(len(idx_tab)) * mask) + 1)
    return slots


def reverse_step(apply_fn, state: Tensor, t: int, coef: Dict[str, np.ndarray], z: Tensor,
                 infill_samples: Optional[Tensor] = None, infill_masks: Optional[Tensor] = None,
                 infill_noise: Optional[Tensor] = None):
    """One body of lax.scan in ebm_utils.py:327-397 with the noise tensors supplied.

    Returns next_state, eps_recon, metrics (grad_norm, step_norm, alpha_prod, noise_norm).
    """
    dt = state.dtype
    c = {k: torch.tensor(v[t], dtype=dt) for k, v in coef.items()}
    noise = (z if t > 0 else torch.zeros_like(state)) * c["sigma"]
    cond = c["sqrt_alpha_prod"] * torch.ones((state.shape[0],) + (1,) * (state.dim() - 1), dtype=dt)
    eps_recon = apply_fn(state, cond)
    recon = c["sqrt_recip_alpha_prod"] * state - c["sqrt_alpha_prod_m1"] * eps_recon
    recon = torch.clamp(recon, -1.0, 1.0)
    next_state = c["mu1"] * recon + c["mu2"] * state + noise
    if infill_masks is not None:
        if t > 0:
            y = c["sqrt_alpha_prod"] * infill_samples + c["sqrt_one_minus_alpha_prod"] * infill_noise
        else:
            y = infill_samples
        next_state = next_state * (1 - infill_masks) + y * infill_masks
    step = state - next_state

    def nrm(a):  # ebm_utils.py:381-383: sqrt(sum(a^2, axis=1) + 1e-10).mean()
        return torch.sqrt((a * a).sum(dim=1) + 1e-10).mean()

    metrics = (nrm(eps_recon), nrm(step), c["alpha_prod"], nrm(noise))
    return next_state, eps_recon, metrics


def diffusion_dynamics(apply_fn, betas: np.ndarray, init: Tensor, noise_fn, infill=False,
                       infill_samples=None, infill_masks=None, steps: Optional[int] = None):
    """ebm_utils.py:274-405.  noise_fn(step_index, t) -> (z, infill_noise).

    ``steps`` (oracle-only) truncates the chain after that many reverse steps.
    Returns state, collection (41, *init.shape), ld_metrics (4, T, 1).
    """
    T = len(betas)
    coef = reverse_coefficients(betas)
    slots = collection_slots(T)
    state = init           # ebm_utils.py:398: the scan starts from the raw init ...
    collection = torch.zeros((41,) + tuple(init.shape), dtype=init.dtype)
    # ... and the merged start only goes to collection[0] (ebm_utils.py:321-323)
    collection[0] = init * (1 - infill_masks) + infill_samples * infill_masks if infill else init
    mets = torch.zeros((4, T, 1), dtype=init.dtype)
    n = T if steps is None else steps
    for i, t in enumerate(range(T - 1, T - 1 - n, -1)):
        z, inz = noise_fn(i, t)
        state, _, m = reverse_step(apply_fn, state, t, coef, z,
                                   infill_samples if infill else None,
                                   infill_masks if infill else None, inz)
        for j in range(4):
            mets[j, i, 0] = m[j]
        if slots[t] >= 0:
            collection[slots[t]] = state
    return state, collection, mets


# ----------------------------------------------------------------------------
# train_ncsn.py:260-288 train_step tail; flax.optim.Adam; jax clip_grads; EMA; LR schedule
# ----------------------------------------------------------------------------
def l2_norm(grads: Dict[str, Tensor]) -> Tensor:
    return torch.sqrt(sum((g * g).sum() for g in grads.values()))


def clip_grads(grads: Dict[str, Tensor], max_norm: float) -> Dict[str, Tensor]:
    """jax.experimental.optimizers.clip_grads: g if norm < max else g * (max / norm)."""
    n = l2_norm(grads)
    if n < max_norm:
        return dict(grads)
    return {k: g * (max_norm / n) for k, g in grads.items()}


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              beta1=0.9, beta2=0.999, eps=1e-8):
    """flax.optim.Adam 0.3.0 apply_param_gradient (weight_decay 0). step = count BEFORE this update."""
    m = (1 - beta1) * g + beta1 * m
    v = (1 - beta2) * g * g + beta2 * v
    t = step + 1.0
    m_hat = m / (1 - beta1 ** t)
    v_hat = v / (1 - beta2 ** t)
    p = p - lr * m_hat / (torch.sqrt(v_hat) + eps)
    return p, m, v


def ema_update(p_ema: Tensor, p: Tensor, mu: float = 0.999) -> Tensor:
    """utils/train_utils.py:73-78: p_ema*mu + p*(1-mu)."""
    return p_ema * mu + p * (1 - mu)


def stepped_lr(base_lr: float, step: int, interval: int = 10000, gamma: float = 0.98,
               n: int = 1000) -> float:
    """flax lr_schedule.create_stepped_learning_rate_schedule as called at train_ncsn.py:340-342."""
    boundaries = np.array([round(i * interval) for i in range(n)])
    values = np.array([1.0] + [gamma ** i for i in range(n)]) * base_lr
    return float(values[int(np.sum(boundaries < step))])


def train_step(arch, p, m, v, step, batch, used_alphas, eps, lr, grad_clip=1.0, model_kw=None):
    """train_ncsn.py:260-288 with tensor-level noise inputs.  Returns new (p,m,v), loss, grad_norm(post-clip), raw grads."""
    model_kw = model_kw or {}
    leaves = {k: t.detach().clone().requires_grad_(True) for k, t in p.items()}
    loss, _ = diffusion_loss_tensors(lambda x, t: model_apply(arch, leaves, x, t, **model_kw),
                                     batch, used_alphas, eps, "mean")
    loss.backward()
    grads = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in leaves.items()}
    clipped = clip_grads(grads, grad_clip)
    gnorm = l2_norm(clipped)
    np_, nm, nv = {}, {}, {}
    for k in p:
        np_[k], nm[k], nv[k] = adam_step(p[k].detach(), clipped[k], m[k], v[k], step, lr)
    return (np_, nm, nv), loss.detach(), gnorm, grads

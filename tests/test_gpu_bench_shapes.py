"""Parity at the sizes BASELINE.json benchmarks (not only at toy batches): the CUDA path through the C ABI vs the
true fp32 CPU oracle on the same seeded inputs.

  cfg1  ddpm-mel-1seq-512      DenseDDPM L6, B=8, C=512            loss + gradients
  cfg2  ddpm-mel-32seq-512     base L6/H8/K2, B=128, C=42          forward, loss, gradients   (4096 tokens: split-K
                                                                    choices, weight-gradient stream, 3-stream overlap)
  cfg3  same model             N=1000 samples, one reverse step    (32000 tokens: fused FFN auto-path, FiLM table,
                                                                    CUDA graph replay, in-kernel threefry noise)
  cfg4  ddpm-mel-32seq-512-large  L8/H16/K3, B=128/GPU              forward, loss, gradients
  cfg5  ddpm-multi-32seq-512   base, C=146, N=1000                  one reverse step
  c512  base, C=512 (no slice) B=128                                forward + loss

Every test appends the MEASURED errors to gpurun_out/parity_r02.json (copied to profiles/ after a GPU run); the
asserted bounds are 2x the values measured on B200 (profiles/r02_parity_measured.json) -- bf16 tensor-core operands,
fp32 accumulation, against an all-fp32 reference: SURVEY section 7 "Precision vs parity".
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import threefry as tf
from tests.util import oracle_kwargs, params_torch, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE = dict(num_layers=6, num_heads=8, num_mlp_layers=2)
LARGE = dict(num_layers=8, num_heads=16, num_mlp_layers=3)

# Stated tolerances vs the true fp32 oracle = 2x the worst value measured on B200 (profiles/r02_parity_measured.json):
#   eps_hat rel-L2 3.8e-3 .. 6.0e-3, max-abs 1.4e-2 .. 3.0e-2; |d loss| / loss 1.7e-5 .. 7.7e-4 (worst single example
#   2.1e-3); gradient: worst tensor rel-L2 6.9e-3 .. 8.9e-3, norm ratio within 1e-3, 1 - cosine < 1e-4;
#   state after one reverse step 1.1e-5; 200-step chain from t=999 4.6e-4, 300-step chain down to t=0 4.7e-3.
# (SURVEY section 7 proposed rel-L2 5e-3 / max-abs 3e-2 / |d loss| 2e-3 for bf16 operands: met to within 20%.)
TOL_FWD_REL_L2 = 1.2e-2
TOL_FWD_MAX_ABS = 6e-2
TOL_LOSS_REL = 2e-3
TOL_LOSS_REL_EXAMPLE = 5e-3
TOL_GRAD_COS = 0.9998
TOL_GRAD_NORM = 3e-3
TOL_GRAD_TENSOR = 2e-2
TOL_STATE = 3e-5
TOL_PATHS = 2e-6
TOL_CHAIN = {999: 1e-3, 299: 1e-2}


def record(name, **vals):
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_r02.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = {k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in vals.items()}
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass
    print(f"[parity] {name}: " + " ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items()))


def _train_engine(arch, kw, batch, seed=1):
    from smd_b200 import Engine, ModelConfig
    eng = Engine(ModelConfig(arch=arch, **kw), max_batch=batch, cta_group=2, training=True)
    flat = eng.init_params(seed=seed, perturb=0.02)
    eng.set_params(flat)
    eng.init_train_state()
    return eng, flat


def _draws(batch, shape, seed=0):
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-1, 1, (batch, *shape)).astype(np.float32)
    eps = rng.standard_normal((batch, *shape)).astype(np.float32)
    ap = O.alphas_prod_with_one(O.create_noise_schedule(1e-6, 0.01, 1000, "linear"))
    labels = rng.integers(1, 1001, size=batch)
    return x0, ap[labels - 1].astype(np.float32), eps


def _check_loss_and_grads(name, arch, kw, batch, shape):
    eng, flat = _train_engine(arch, kw, batch)
    x0, used, eps = _draws(batch, shape)
    dx0, dused, deps = (torch.from_numpy(a).cuda() for a in (x0, used, eps))
    loss_dev, pred = eng.ddpm_loss(dx0, dused, deps, want_pred=True)
    eng.compute_grads(dx0, dused, deps)
    torch.cuda.synchronize()
    got = eng.flat_to_dict(eng.grads)
    loss_tr = float(eng.loss_sum) / batch

    p = {k: v.clone().requires_grad_(True) for k, v in params_torch(eng, flat).items()}
    okw = oracle_kwargs(eng.cfg)
    per_ex, ref_pred = O.diffusion_loss_tensors(lambda a, c: O.model_apply(arch, p, a, c, **okw), torch.from_numpy(x0),
                                                torch.from_numpy(used), torch.from_numpy(eps), "none")
    loss_ref = per_ex.mean()
    loss_ref.backward()
    ref = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    loss_ref = float(loss_ref)

    fwd = rel_l2(pred, ref_pred)
    mabs = float((pred.cpu() - ref_pred.detach()).abs().max())
    dl = abs(float(loss_dev.mean()) - loss_ref) / loss_ref
    dl_tr = abs(loss_tr - loss_ref) / loss_ref
    dl_ex = float(((loss_dev.cpu() - per_ex.detach()).abs() / per_ex.detach()).max())
    total = sum(float((g.double() ** 2).sum()) for g in ref.values())
    dot = nn_got = 0.0
    worst, worst_name = 0.0, ""
    for k, g in ref.items():
        gg = torch.from_numpy(got[k])
        dot += float((gg.double() * g.double()).sum())
        nn_got += float((gg.double() ** 2).sum())
        if float((g.double() ** 2).sum()) >= 1e-4 * total:
            e = rel_l2(gg, g)
            if e > worst:
                worst, worst_name = e, k
    cos = dot / (np.sqrt(nn_got) * np.sqrt(total))
    nratio = float(np.sqrt(nn_got / total))
    record(name, fwd_rel_l2=fwd, fwd_max_abs=mabs, dloss_rel=dl, dloss_rel_train_path=dl_tr, dloss_rel_worst_example=dl_ex,
           grad_cos=cos, grad_one_minus_cos=1.0 - cos, grad_norm_ratio=nratio, grad_worst_tensor_rel_l2=worst, grad_worst_tensor=worst_name,
           tokens=batch * (shape[0] if len(shape) == 2 else 1), loss=loss_ref)
    assert fwd < TOL_FWD_REL_L2 and mabs < TOL_FWD_MAX_ABS
    assert dl < TOL_LOSS_REL and dl_tr < TOL_LOSS_REL and dl_ex < TOL_LOSS_REL_EXAMPLE
    assert cos > TOL_GRAD_COS and abs(nratio - 1.0) < TOL_GRAD_NORM
    assert worst < TOL_GRAD_TENSOR, (worst, worst_name)


def test_cfg2_train_batch128_forward_loss_grads(lib):
    _check_loss_and_grads("cfg2_train_b128_c42", "TransformerDDPM", dict(channels=42, **BASE), 128, (32, 42))


def test_cfg4_large_batch128_forward_loss_grads(lib):
    _check_loss_and_grads("cfg4_large_b128_c42", "TransformerDDPM", dict(channels=42, **LARGE), 128, (32, 42))


def test_cfg1_dense_batch8_loss_grads(lib):
    _check_loss_and_grads("cfg1_dense_b8_c512", "DenseDDPM", dict(num_layers=6, channels=512), 8, (512,))


def test_c512_noslice_batch128_forward_loss(lib):
    from smd_b200 import Engine, ModelConfig
    kw = dict(channels=512, **BASE)
    eng = Engine(ModelConfig(**kw), max_batch=128, cta_group=2)
    flat = eng.init_params(seed=1, perturb=0.02)
    eng.set_params(flat)
    x0, used, eps = _draws(128, (32, 512))
    loss, pred = eng.ddpm_loss(*(torch.from_numpy(a).cuda() for a in (x0, used, eps)), want_pred=True)
    p = params_torch(eng, flat)
    okw = oracle_kwargs(eng.cfg)
    with torch.no_grad():
        per_ex, ref_pred = O.diffusion_loss_tensors(lambda a, c: O.transformer_ddpm(p, a, c, **okw), torch.from_numpy(x0),
                                                    torch.from_numpy(used), torch.from_numpy(eps), "none")
    fwd = rel_l2(pred, ref_pred)
    mabs = float((pred.cpu() - ref_pred).abs().max())
    dl = abs(float(loss.mean()) - float(per_ex.mean())) / float(per_ex.mean())
    record("c512_noslice_b128", fwd_rel_l2=fwd, fwd_max_abs=mabs, dloss_rel=dl)
    assert fwd < TOL_FWD_REL_L2 and mabs < TOL_FWD_MAX_ABS and dl < TOL_LOSS_REL


def _noise_keys(key, steps):
    """jax key schedule of diffusion_dynamics (utils/ebm_utils.py:329,342,360): 3 splits per scan step."""
    rkey = np.array(key, np.uint32)
    out = []
    for _ in range(steps):
        rkey, _k = tf.split(rkey, 2)
        rkey, infill_k = tf.split(rkey, 2)
        rkey, noise_k = tf.split(rkey, 2)
        out.append(noise_k)
    return out


@pytest.mark.parametrize("name,channels", [("cfg3_sample_n1000_c42", 42), ("cfg5_sample_n1000_c146", 146)])
def test_sampling_n1000_reverse_step_graph_path(lib, name, channels):
    """One reverse step over 1000 samples exactly as sample_ncsn drives it: smd_ddpm_sample with the CUDA graph, the
    per-schedule FiLM table, the fused FFN kernel (auto-engaged at >= 8192 tokens) and in-kernel threefry noise."""
    from smd_b200 import Engine, ModelConfig
    N, key = 1000, (0, 11)
    eng = Engine(ModelConfig(channels=channels, **BASE), max_batch=N, cta_group=2)
    flat = eng.init_params(seed=1, perturb=0.02)
    eng.set_params(flat)
    betas = O.create_noise_schedule(1e-6, 0.01, 1000, "linear")
    eng.sampler_setup(betas, key=key)
    rng = np.random.default_rng(5)
    init = torch.from_numpy(rng.standard_normal((N, 32, channels)).astype(np.float32))
    x = init.clone().cuda()
    mets = torch.zeros((4, 1000), device="cuda")
    eng.sample(x, steps=1, metrics=mets, use_graph=True)
    torch.cuda.synchronize()
    # the model output of the same step through the plain (non-graph) entry point, for the eps_hat comparison
    eh = torch.empty((N, 32, channels), device="cuda")
    z = torch.from_numpy(tf.normal(_noise_keys(key, 1)[0], (N, 32, channels)))
    x2 = eng.reverse_step(init.cuda(), 999, z=z.cuda(), eps_hat=eh)
    torch.cuda.synchronize()

    p = params_torch(eng, flat)
    okw = oracle_kwargs(eng.cfg)
    coef = O.reverse_coefficients(betas)
    with torch.no_grad():
        ref_next, ref_eps, ref_m = O.reverse_step(lambda a, c: O.transformer_ddpm(p, a, c, **okw), init, 999, coef, z)
    e_eps = rel_l2(eh, ref_eps)
    e_abs = float((eh.cpu() - ref_eps).abs().max())
    e_graph = rel_l2(x, ref_next)
    e_plain = rel_l2(x2, ref_next)
    e_paths = rel_l2(x, x2)     # graph + device RNG vs supplied noise: same arithmetic, normals agree to ~2e-5
    m_err = float(np.max(np.abs(mets[:, 0].cpu().numpy() - np.array([float(v) for v in ref_m])) /
                         np.maximum(np.abs(np.array([float(v) for v in ref_m])), 1e-6)))
    record(name, eps_hat_rel_l2=e_eps, eps_hat_max_abs=e_abs, x_next_rel_l2_graph=e_graph, x_next_rel_l2_plain=e_plain,
           graph_vs_plain_rel_l2=e_paths, metrics_rel=m_err, tokens=N * 32)
    assert e_eps < TOL_FWD_REL_L2 and e_abs < TOL_FWD_MAX_ABS
    # x' = mu1 * clip(x/sqrt(abar) - sqrt(1-abar)/sqrt(abar) eps_hat) + mu2 x + sigma z with mu1(t=999) ~ 8e-4:
    # the eps_hat error is damped by mu1 * 12.2 ~ 1e-2 before it reaches the state
    assert e_graph < TOL_STATE and e_plain < TOL_STATE and e_paths < TOL_PATHS
    assert m_err < 5e-3


@pytest.mark.parametrize("t0", [999, 299])
def test_chain_200_steps_error_growth(lib, t0):
    """>= 200 consecutive reverse steps (supplied noise, N=8) against the fp32 oracle chain: the bf16 error must not
    grow beyond the stated bound (the reconstruction x/sqrt(abar) - ... amplifies eps_hat error ~12x at t ~ 999 but
    mu1 damps it; near t = 0 mu1 -> 1 and sqrt(1 - abar) -> 0)."""
    from smd_b200 import Engine, ModelConfig
    N, steps = 8, 200 if t0 == 999 else 300
    eng = Engine(ModelConfig(channels=42, **BASE), max_batch=N, cta_group=2)
    flat = eng.init_params(seed=1, perturb=0.02)
    eng.set_params(flat)
    betas = O.create_noise_schedule(1e-6, 0.01, 1000, "linear")
    eng.sampler_setup(betas, key=(0, 3))
    rng = np.random.default_rng(t0)
    scale = 1.0 if t0 == 999 else 0.6
    init = torch.from_numpy((scale * rng.standard_normal((N, 32, 42))).astype(np.float32))
    zs = [torch.from_numpy(rng.standard_normal((N, 32, 42)).astype(np.float32)) for _ in range(steps)]
    p = params_torch(eng, flat)
    okw = oracle_kwargs(eng.cfg)
    coef = O.reverse_coefficients(betas)
    x = init.clone().cuda()
    ref = init.clone()
    errs = []
    with torch.no_grad():
        for i in range(steps):
            t = t0 - i
            x = eng.reverse_step(x, t, z=zs[i].cuda())
            ref = O.reverse_step(lambda a, c: O.transformer_ddpm(p, a, c, **okw), ref, t, coef, zs[i])[0]
            if (i + 1) % 25 == 0 or i == steps - 1:
                errs.append(rel_l2(x, ref))
    record(f"chain_{steps}_steps_from_t{t0}", final_rel_l2=errs[-1], max_rel_l2=max(errs),
           trajectory=[float(e) for e in errs])
    assert max(errs) < TOL_CHAIN[t0], errs

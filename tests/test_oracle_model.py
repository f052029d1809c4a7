"""Known-answer and self-consistency pins for the CPU restatement (the reference ships no golden vectors)."""
import math

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from smd_b200 import Engine, ModelConfig
from tests.util import oracle_kwargs, params_torch, rel_l2


def _engine(**kw):
    return Engine(ModelConfig(**kw), max_batch=4)


def test_param_counts_match_reference_report():
    # what utils/train_utils.py:121-131 (report_model) would log for the shipped configs
    assert _engine().num_params == 25_579_946
    assert _engine(num_layers=8, num_heads=16, num_mlp_layers=3).num_params == 37_596_842
    assert Engine(ModelConfig(arch="DenseDDPM", channels=512), 4).num_params == 67_088_896


def test_zero_weights_give_final_bias():
    eng = _engine(num_layers=1, num_mlp_layers=1)
    d = {k: np.zeros_like(v) for k, v in eng.flat_to_dict(eng.init_params(0)).items()}
    d["out.bias"] = np.arange(42, dtype=np.float32) * 0.01
    p = {k: torch.from_numpy(v) for k, v in d.items()}
    y = O.transformer_ddpm(p, torch.randn(2, 32, 42), torch.tensor([0.3, 0.9]), **oracle_kwargs(eng.cfg))
    assert torch.allclose(y, torch.from_numpy(d["out.bias"]).expand(2, 32, 42), atol=0)


def test_fp32_vs_fp64_consistency():
    eng = _engine(num_layers=2, num_mlp_layers=1)
    flat = eng.init_params(3, perturb=0.02)
    x = torch.randn(3, 32, 42, dtype=torch.float64)
    t = torch.tensor([0.2, 0.5, 0.99], dtype=torch.float64)
    kw = oracle_kwargs(eng.cfg)
    y64 = O.transformer_ddpm(params_torch(eng, flat, torch.float64), x, t, **kw)
    y32 = O.transformer_ddpm(params_torch(eng, flat), x.float(), t.float(), **kw)
    assert rel_l2(y32, y64) < 2e-5
    ybf = O.transformer_ddpm(params_torch(eng, flat), x.float(), t.float(), emulate_bf16=True, **kw)
    assert rel_l2(ybf, y64) < 3e-2  # the stated bf16-operand tolerance of the CUDA path


def test_dense_ddpm_ignores_transformer_kwargs():
    eng = Engine(ModelConfig(arch="DenseDDPM", channels=512, num_layers=2), 4)
    p = params_torch(eng, eng.init_params(0, 0.02))
    x = torch.randn(4, 512)
    t = torch.rand(4, 1)
    a = O.dense_ddpm(p, x, t, num_layers=2)
    b = O.dense_ddpm(p, x, t, num_layers=2, num_heads=8, num_mlp_layers=2)
    assert torch.equal(a, b)


def test_positional_and_noise_encoding_values():
    pe = O.transformer_positional_encoding(32, 128)
    assert pe.shape == (32, 128)
    assert torch.all(pe[0, :64] == 0) and torch.all(pe[0, 64:] == 1)
    assert abs(float(pe[1, 0]) - math.sin(1.0)) < 1e-6 and abs(float(pe[1, 63]) - math.sin(1e-4)) < 1e-9
    ne = O.noise_encoding(torch.tensor([1.0]), 128)
    assert abs(float(ne[0, 0]) - math.sin(5000.0)) < 1e-3  # fp32 argument rounding, large argument
    assert abs(float(ne[0, 64 + 63]) - math.cos(0.5)) < 1e-5


def test_linear_schedule_closed_forms():
    betas = O.create_noise_schedule(1e-6, 0.01, 1000, "linear")
    assert betas.dtype == np.float32 and betas.shape == (1000,)
    assert betas[0] == np.float32(1e-6) and abs(betas[-1] - 0.01) < 1e-8
    c = O.reverse_coefficients(betas)
    assert abs(c["alpha_prod"][-1] - math.exp(np.sum(np.log1p(-betas.astype(np.float64))))) < 1e-6
    # in fp32 mu1(t=0) is NOT 1 (cancellation in 1 - alpha_bar_0)
    assert abs(c["mu1"][0] - 0.98690) < 2e-3
    assert c["mu2"][0] == 0.0 and c["sigma"][0] == np.float32(np.exp(np.float32(0.5) * np.log(np.float32(1e-20))))
    with pytest.raises(ValueError):
        O.create_noise_schedule(1, 2, 3, "cosine")
    fib = O.create_noise_schedule(L=6, schedule="fibonacci")
    np.testing.assert_allclose(fib, [1e-6, 2e-6, 3e-6, 5e-6, 8e-6, 13e-6], rtol=1e-6)


def test_reverse_step_t0_has_no_noise_and_clips():
    betas = O.create_noise_schedule(1e-6, 0.01, 1000, "linear")
    coef = O.reverse_coefficients(betas)
    x = torch.randn(3, 32, 5) * 3
    nxt, eps, m = O.reverse_step(lambda a, c: torch.zeros_like(a), x, 0, coef, torch.randn_like(x))
    expect = torch.tensor(coef["mu1"][0]) * torch.clamp(torch.tensor(coef["sqrt_recip_alpha_prod"][0]) * x, -1, 1)
    assert torch.allclose(nxt, expect, atol=1e-6)
    assert abs(float(m[3]) - 1e-5) < 1e-8  # noise norm = sqrt(0 + 1e-10)


def test_collection_slots_quirks():
    s = O.collection_slots(1000)
    assert s[999] == -1 or s[999] >= 1
    written = sorted(set(int(v) for v in s if v >= 0))
    assert 1 not in written          # slot 1 (image_idx == 1) is never reached: image_idx starts at 2
    assert s[0] == -1                # the final state (image_idx = T + 1) is never snapshotted
    assert max(written) == 40


def test_loss_and_optimizer_pieces():
    ap = O.alphas_prod_with_one(O.create_noise_schedule(1e-6, 0.01, 1000, "linear"))
    assert ap[0] == 1.0 and ap.shape == (1001,)
    u = O.uniform_minmax(np.array([0.3, 0.9], np.float32), ap[[4, 9]], ap[[5, 10]])
    np.testing.assert_array_equal(u, ap[[4, 9]])  # D8
    assert O.stepped_lr(1e-3, 0) == pytest.approx(1e-3)
    assert O.stepped_lr(1e-3, 10000) == pytest.approx(1e-3)
    assert O.stepped_lr(1e-3, 10001) == pytest.approx(1e-3 * 0.98)
    assert O.stepped_lr(1e-3, 25000) == pytest.approx(1e-3 * 0.98 ** 2)
    g = {"a": torch.tensor([3.0, 4.0])}
    assert float(O.l2_norm(O.clip_grads(g, 1.0))) == pytest.approx(1.0)
    assert torch.equal(O.clip_grads(g, 10.0)["a"], g["a"])
    p, m, v = O.adam_step(torch.tensor([1.0]), torch.tensor([0.5]), torch.zeros(1), torch.zeros(1), 0, 1e-3)
    assert float(p) == pytest.approx(1.0 - 1e-3 * 0.5 / (0.5 + 1e-8), rel=1e-6)
    assert float(O.ema_update(torch.tensor(1.0), torch.tensor(0.0), 0.999)) == pytest.approx(0.999)


def test_golden_fixture_matches_oracle():
    """tests/golden/transformer_tiny.npz was produced by scripts/make_golden.py from this oracle in float64."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "transformer_tiny.npz")
    g = np.load(path)
    eng = Engine(ModelConfig(num_layers=int(g["num_layers"]), num_heads=int(g["num_heads"]),
                             num_mlp_layers=int(g["num_mlp_layers"]), channels=int(g["channels"])), 4)
    flat = eng.init_params(int(g["param_seed"]), perturb=float(g["perturb"]))
    y = O.transformer_ddpm(params_torch(eng, flat, torch.float64), torch.from_numpy(g["x"]).double(),
                           torch.from_numpy(g["t"]).double(), **oracle_kwargs(eng.cfg))
    assert rel_l2(y, torch.from_numpy(g["y64"])) < 1e-12


def test_dsm_oracle_semantics():
    """utils/losses.py:146-179: with a decreasing schedule the `continuous` uniform(minval=sigmas[l-1], maxval=sigmas[l])
    has minval > maxval and collapses to sigmas[l-1] (jax 0.2.8 max(minval, ...)); a perfect score network
    (score = -noise / sigma^2) has zero loss; the loss of the zero network is 0.5 * |eps|^2."""
    from oracle import threefry as tf
    sigmas = O.create_noise_schedule(1.0, 0.01, 15, "geometric")
    assert sigmas[0] > sigmas[-1]
    key = tf.prng_key(1)
    labels, used, eps = O.dsm_draws(key, (32, 8), sigmas, continuous_noise=True)
    assert labels.min() >= 1 and labels.max() <= 14
    np.testing.assert_array_equal(used, sigmas[labels - 1])
    labels0, used0, _ = O.dsm_draws(key, (32, 8), sigmas, continuous_noise=False)
    assert labels0.min() >= 0 and labels0.max() <= 14
    np.testing.assert_array_equal(used0, sigmas[labels0])
    x = torch.zeros(32, 8)
    e = torch.from_numpy(eps)
    us = torch.from_numpy(used)
    perfect = lambda a, s: -(a - x) / (s ** 2)
    loss, _ = O.dsm_loss_tensors(perfect, x, us, e, "none")
    assert float(loss.abs().max()) < 1e-8
    zero = lambda a, s: torch.zeros_like(a)
    loss0, _ = O.dsm_loss_tensors(zero, x, us, e, "none")
    np.testing.assert_allclose(loss0.numpy(), 0.5 * (e ** 2).sum(-1).numpy(), rtol=1e-5)


def test_langevin_oracles_with_a_zero_score():
    """With score == 0 and no noise both samplers leave the state untouched; the collection / metrics shapes are those of
    utils/ebm_utils.py:123-129,196-198 and :265-271."""
    sig = O.create_noise_schedule(1.0, 0.05, 4, "geometric")
    init = torch.ones(3, 5)
    z0 = lambda *a: (torch.zeros(3, 5), torch.zeros(3, 5))
    st, coll, m = O.annealed_langevin_dynamics(lambda a, s: torch.zeros_like(a), sig, init, 1e-4, 3, True, z0)
    assert torch.equal(st, init) and coll.shape == (102, 3, 5) and m.shape == (4, 4, 3)
    assert torch.equal(coll[0], init) and torch.equal(coll[-1], init)
    np.testing.assert_allclose(m[2, :, 0].numpy(), 1e-4 * (sig / sig[-1]) ** 2, rtol=1e-6)
    st, m = O.consistent_langevin_dynamics(lambda a, s: torch.zeros_like(a), sig, init, 1e-4, True, lambda i: torch.zeros(3, 5))
    assert torch.equal(st, init) and m.shape == (4, 4, 1)

"""NCSN family: DenseNCSN score network, denoising score matching (loss, draws, gradients) and the
annealed / consistent Langevin samplers -- CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.
Tolerances as for the DDPM path (bf16 tensor-core operands): the division by sigma scales signal and error alike."""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import threefry as tf
from tests.util import params_torch, rel_l2

pytestmark = pytest.mark.gpu
KW = dict(num_layers=2, mlp_dims=2048)
C = 64


def _engine(batch, training=False):
    from smd_b200 import Engine, ModelConfig
    eng = Engine(ModelConfig(arch="DenseNCSN", channels=C, **KW), max_batch=batch, cta_group=2, training=training)
    flat = eng.init_params(seed=4, perturb=0.02)
    eng.set_params(flat)
    return eng, flat


def test_dense_ncsn_forward_is_the_ddpm_stack_divided_by_sigma(lib):
    eng, flat = _engine(8)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (8, C)).astype(np.float32)
    sig = rng.uniform(0.05, 1.0, (8,)).astype(np.float32)
    y = eng.forward(torch.from_numpy(x).cuda(), torch.from_numpy(sig).cuda())
    p = params_torch(eng, flat)
    ref_bf = O.dense_ncsn(p, torch.from_numpy(x), torch.from_numpy(sig), emulate_bf16=True, **KW)
    ref32 = O.dense_ncsn(p, torch.from_numpy(x), torch.from_numpy(sig), **KW)
    assert rel_l2(y, ref_bf) < 1e-2 and rel_l2(y, ref32) < 1.2e-2


@pytest.mark.parametrize("continuous", [False, True])
def test_dsm_draws_match_jax_restatement(lib, continuous):
    eng, _ = _engine(64)
    sigmas = O.create_noise_schedule(1.0, 0.01, 15, "geometric")
    eng.dsm_setup(sigmas)
    key = tf.prng_key(3)
    used, eps, lab = eng.dsm_draws((int(key[0]), int(key[1])), 64, want_labels=True, continuous_noise=continuous)
    rl, ru, re = O.dsm_draws(key, (64, C), sigmas, continuous)
    np.testing.assert_array_equal(lab.cpu().numpy(), rl)
    np.testing.assert_array_equal(used.cpu().numpy(), ru)
    np.testing.assert_allclose(eps.cpu().numpy(), re, rtol=2e-5, atol=2e-6)
    # rows 16..31 of the same global draw (data-parallel slice)
    u2, e2 = eng.dsm_draws((int(key[0]), int(key[1])), 16, global_batch=64, first_row=16, continuous_noise=continuous)
    assert torch.equal(u2, used[16:32]) and torch.equal(e2, eps[16:32])


def test_dsm_loss_and_gradients(lib):
    eng, flat = _engine(8, training=True)
    eng.init_train_state()
    rng = np.random.default_rng(1)
    x0 = rng.uniform(-1, 1, (8, C)).astype(np.float32)
    eps = rng.standard_normal((8, C)).astype(np.float32)
    sig = O.create_noise_schedule(1.0, 0.01, 15, "geometric")[rng.integers(0, 15, 8)].astype(np.float32)
    dev = lambda a: torch.from_numpy(a).cuda()
    loss, scores = eng.dsm_loss(dev(x0), dev(sig), dev(eps), want_pred=True)
    eng.compute_dsm_grads(dev(x0), dev(sig), dev(eps))
    torch.cuda.synchronize()
    p = {k: v.clone().requires_grad_(True) for k, v in params_torch(eng, flat).items()}
    per_ex, ref_scores = O.dsm_loss_tensors(lambda a, s: O.dense_ncsn(p, a, s, **KW), torch.from_numpy(x0),
                                            torch.from_numpy(sig), torch.from_numpy(eps), "none")
    per_ex.mean().backward()
    assert rel_l2(scores, ref_scores) < 1.2e-2
    np.testing.assert_allclose(loss.cpu().numpy(), per_ex.detach().numpy(), rtol=2e-2)
    assert abs(float(eng.loss_mean) - float(per_ex.mean())) < 1e-2 * float(per_ex.mean())
    got = eng.flat_to_dict(eng.grads)
    tot = sum(float((v.grad ** 2).sum()) for v in p.values())
    dot = sum(float((torch.from_numpy(got[k]) * v.grad).sum()) for k, v in p.items())
    nn_ = sum(float((torch.from_numpy(got[k]) ** 2).sum()) for k in p)
    assert dot / np.sqrt(nn_ * tot) > 0.9995 and abs(np.sqrt(nn_ / tot) - 1) < 1e-2
    for k, v in p.items():
        if float((v.grad ** 2).sum()) >= 1e-4 * tot:
            assert rel_l2(torch.from_numpy(got[k]), v.grad) < 3e-2, k


def _model(eng, flat):
    from smd_b200 import ncsn, nn
    module = ncsn.DenseNCSN.partial(**KW)
    arena = nn.ParamArena(module, (C,), flat)
    return nn.Model(module, arena)


def test_annealed_langevin_short_chain(lib):
    """3 noise levels x 4 steps with in-kernel threefry noise (key schedule of ebm_utils.py:139), infill mask on."""
    from smd_b200 import ebm_utils
    eng, flat = _engine(4)
    model = _model(eng, flat)
    p = params_torch(eng, flat)
    sigmas = np.asarray([1.0, 0.3, 0.05], np.float32)
    rng = np.random.default_rng(2)
    init = torch.from_numpy(rng.uniform(-1.7, 1.7, (4, C)).astype(np.float32))
    samples = torch.from_numpy(rng.uniform(-1, 1, (4, C)).astype(np.float32))
    masks = torch.zeros(4, C); masks[:, :16] = 1
    key = tf.prng_key(11)
    state, coll, mets = ebm_utils.annealed_langevin_dynamics(key, model, sigmas, init, 2e-5, 4, True, True, samples, masks)
    keys = {}
    k = key
    for si in range(3):
        for i in range(4):
            k, sk, ik = tf.split(k, 3)
            keys[(si, i)] = (sk, ik)
    noise_fn = lambda si, i: (torch.from_numpy(tf.normal(keys[(si, i)][0], (4, C))),
                              torch.from_numpy(tf.normal(keys[(si, i)][1], (4, C))))
    ref_state, ref_coll, ref_m = O.annealed_langevin_dynamics(
        lambda a, s: O.dense_ncsn(p, a, s, emulate_bf16=True, **KW), sigmas, init, 2e-5, 4, True, noise_fn, True, samples, masks)
    assert rel_l2(state, ref_state) < 1e-2
    assert coll.shape == (102, 4, C) and rel_l2(coll[0], ref_coll[0]) < 1e-6 and rel_l2(coll[-1], ref_coll[-1]) < 1e-2
    np.testing.assert_allclose(mets.cpu().numpy(), ref_m.numpy(), rtol=2e-2, atol=1e-7)
    assert mets.shape == (4, 3, 4)
    # the masked entries hold y = samples + sigma_last * z' after the last step, then the denoising step moves them
    assert torch.isfinite(state).all()


def test_consistent_langevin_chain(lib):
    from smd_b200 import ebm_utils
    eng, flat = _engine(4)
    model = _model(eng, flat)
    p = params_torch(eng, flat)
    sigmas = O.create_noise_schedule(1.0, 0.05, 6, "geometric")
    init = torch.from_numpy(np.random.default_rng(5).uniform(-1.7, 1.7, (4, C)).astype(np.float32))
    key = tf.prng_key(12)
    state, coll, mets = ebm_utils.consistent_langevin_dynamics(key, model, sigmas, init, 1e-4, 1, True)
    ks, k = [], key
    for _ in range(6):
        k, sk = tf.split(k, 2)
        ks.append(sk)
    ref_state, ref_m = O.consistent_langevin_dynamics(lambda a, s: O.dense_ncsn(p, a, s, emulate_bf16=True, **KW), sigmas,
                                                      init, 1e-4, True, lambda i: torch.from_numpy(tf.normal(ks[i], (4, C))))
    assert coll is None and mets.shape == (4, 6, 1)
    assert rel_l2(state, ref_state) < 1e-2
    np.testing.assert_allclose(mets.cpu().numpy(), ref_m.numpy(), rtol=2e-2, atol=1e-7)

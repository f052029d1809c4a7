"""Reverse-diffusion sampler parity (utils/ebm_utils.py:274-405) and device threefry parity."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import threefry as tf
from tests.util import TRANSFORMER_CASES, oracle_kwargs, params_torch, rel_l2

pytestmark = pytest.mark.gpu


def _setup(batch=4, case="tiny", key=(0, 7)):
    from smd_b200 import Engine, ModelConfig
    kw, _ = TRANSFORMER_CASES[case]
    eng = Engine(ModelConfig(**kw), max_batch=batch, cta_group=2)
    flat = eng.init_params(seed=1, perturb=0.02)
    eng.set_params(flat)
    betas = O.create_noise_schedule(1e-6, 0.01, 1000, "linear")
    eng.sampler_setup(betas, key=key)
    p = params_torch(eng, flat)
    okw = oracle_kwargs(eng.cfg)
    apply_bf = lambda a, c: O.transformer_ddpm(p, a, c, emulate_bf16=True, **okw)
    return eng, betas, apply_bf


@pytest.mark.parametrize("n", [1, 2, 1000, 4097, 32 * 42 * 5])
def test_device_threefry_normal_matches_jax_restatement(lib, n):
    from smd_b200 import lib as L
    key = (C.c_uint32 * 2)(0, 42)
    out = torch.empty((n,), device="cuda")
    L.check(lib.smd_threefry_normal(key, out.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    ref = tf.normal(tf.prng_key(42), (n,))
    # logf / polynomial rounding may differ from XLA's by a few ulp
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-5, atol=2e-6)


def test_device_threefry_known_jax_value(lib):
    from smd_b200 import lib as L
    key = (C.c_uint32 * 2)(0, 42)
    out = torch.empty((3,), device="cuda")
    L.check(lib.smd_threefry_normal(key, out.data_ptr(), 3, torch.cuda.current_stream().cuda_stream))
    np.testing.assert_allclose(out.cpu().numpy(), [0.18693547, -1.2806505, -1.5593132], rtol=1e-5)


@pytest.mark.parametrize("t", [999, 500, 1, 0])
def test_reverse_step_supplied_noise(lib, t):
    eng, betas, apply_bf = _setup()
    rng = np.random.default_rng(t)
    x = torch.from_numpy(rng.standard_normal((4, 32, 42)).astype(np.float32))
    z = torch.from_numpy(rng.standard_normal((4, 32, 42)).astype(np.float32))
    mets = torch.zeros((4, 1000), device="cuda")
    eh = torch.empty((4, 32, 42), device="cuda")
    nxt = eng.reverse_step(x.cuda(), t, z=z.cuda(), eps_hat=eh, metrics=mets)
    coef = O.reverse_coefficients(betas)
    ref_next, ref_eps, ref_m = O.reverse_step(apply_bf, x, t, coef, z)
    assert rel_l2(eh, ref_eps) < 1e-2
    # the x/sqrt(abar) - ... reconstruction amplifies eps_hat error by sqrt(1-abar)/sqrt(abar) (~12 at t=999)
    assert rel_l2(nxt, ref_next) < 1e-2
    col = mets[:, 999 - t].cpu().numpy()
    np.testing.assert_allclose(col, [float(m) for m in ref_m], rtol=5e-3, atol=1e-6)


def test_reverse_step_infill(lib):
    eng, betas, apply_bf = _setup()
    rng = np.random.default_rng(9)
    mk = lambda: torch.from_numpy(rng.standard_normal((4, 32, 42)).astype(np.float32))
    x, z, iz, ix = mk(), mk(), mk(), mk().clamp(-1, 1)
    mask = torch.zeros(4, 32, 42)
    mask[:, :16] = 1.0
    nxt = eng.reverse_step(x.cuda(), 300, z=z.cuda(), infill_x=ix.cuda(), infill_mask=mask.cuda(), infill_z=iz.cuda())
    coef = O.reverse_coefficients(betas)
    ref_next, _, _ = O.reverse_step(apply_bf, x, 300, coef, z, ix, mask, iz)
    assert rel_l2(nxt, ref_next) < 1e-2


@pytest.mark.parametrize("use_graph", [False, True])
def test_short_chain_with_device_rng(lib, use_graph):
    """First 6 steps of the chain with in-kernel threefry noise (jax key schedule of ebm_utils.py:329,342,360)."""
    key = (0, 7)
    eng, betas, apply_bf = _setup(key=key)
    rng = np.random.default_rng(0)
    init = torch.from_numpy(rng.standard_normal((4, 32, 42)).astype(np.float32))
    x = init.clone().cuda()
    coll = torch.zeros((41, 4, 32, 42), device="cuda")
    mets = torch.zeros((4, 1000), device="cuda")
    eng.sample(x, steps=6, collection=coll, metrics=mets, use_graph=use_graph)
    torch.cuda.synchronize()

    rkey = np.array(key, np.uint32)
    keys = []
    for _ in range(6):
        rkey, _unused = tf.split(rkey, 2)
        rkey, infill_k = tf.split(rkey, 2)
        rkey, noise_k = tf.split(rkey, 2)
        keys.append(noise_k)

    def noise_fn(i, t):
        return torch.from_numpy(tf.normal(keys[i], (4, 32, 42))), None

    ref_state, ref_coll, ref_m = O.diffusion_dynamics(apply_bf, betas, init, noise_fn, steps=6)
    assert rel_l2(x, ref_state) < 1e-2
    np.testing.assert_allclose(mets[:, :6].cpu().numpy(), ref_m[:, :6, 0].numpy(), rtol=2e-2, atol=1e-6)
    # slot 2.. are written when image_idx hits linspace(1,1000,40).int32: step t=998 -> image_idx 3? check table
    slots = O.collection_slots(1000)
    for i, t in enumerate(range(999, 993, -1)):
        if slots[t] >= 0:
            assert rel_l2(coll[slots[t]], ref_coll[slots[t]]) < 1e-2


@pytest.mark.parametrize("n,lo,hi", [(7, 0.0, 1.0), (4096, -1.0, 1.0), (1000 * 32 * 42 + 1, 0.25, 0.75)])
def test_device_threefry_uniform_matches_jax_restatement(lib, n, lo, hi):
    """jax.random.uniform (0.2.8) -- the infill chain's initial state (sample_ncsn.py:230) -- bit-exact."""
    from oracle import threefry as T
    from smd_b200 import jrandom
    key = jrandom.PRNGKey(11)
    got = jrandom.uniform(key, (n,), lo, hi).cpu().numpy()
    ref = T.uniform(np.asarray(key, np.uint32), (n,), lo, hi)
    assert got.dtype == np.float32 and np.array_equal(got, ref)

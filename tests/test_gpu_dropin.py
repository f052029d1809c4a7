"""End-to-end through the reference-shaped Python surface (train_ncsn / sample_ncsn functions, nn.Model, optim,
checkpoints) on the GPU with a small flag set: the path a user of the reference would run."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_model_protocol_and_objective(lib):
    from smd_b200 import ebm_utils, jrandom as random, ncsn, nn
    from smd_b200.losses import diffusion_loss
    module = ncsn.TransformerDDPM.partial(num_layers=1, num_heads=8, num_mlp_layers=1, mlp_dims=2048)
    _, params = module.init_by_shape(random.PRNGKey(0), [((4, 32, 42), np.float32), ((4, 1, 1), np.float32)])
    model = nn.Model(module, params)
    assert set(model.params) >= {"in", "l0", "post", "k0", "out"} and model.params["in"]["kernel"].shape == (42, 128)
    x = np.random.default_rng(0).uniform(-1, 1, (4, 32, 42)).astype(np.float32)
    t = np.full((4, 1, 1), 0.5, np.float32)
    y = model(x, t)
    p = {k: v.cpu() for k, v in model.arena.as_dict().items()}
    ref = O.transformer_ddpm(p, torch.from_numpy(x), torch.from_numpy(t), num_layers=1, num_heads=8, num_mlp_layers=1)
    assert rel_l2(y, ref) < 3e-2
    betas = ebm_utils.create_noise_schedule(1e-6, 0.01, 1000, "linear")
    key = random.PRNGKey(3)
    loss = diffusion_loss(x, model, betas, key, True, "mean")
    labels, used, eps = O.diffusion_loss_draws(key, x.shape, betas, True)
    oloss, _ = O.diffusion_loss_tensors(lambda a, c: O.transformer_ddpm(p, a, c, num_layers=1, num_heads=8,
                                                                        num_mlp_layers=1),
                                        torch.from_numpy(x), torch.from_numpy(used), torch.from_numpy(eps), "mean")
    assert abs(float(loss) - float(oloss)) < 2e-2 * float(oloss)
    with pytest.raises(ValueError):
        diffusion_loss(x, model, betas, key, True, "median")
    # replace(params=...) shares / swaps arenas like nn.Model.replace
    m2 = model.replace(params=model.arena.clone())
    assert m2.arena is not model.arena and torch.equal(m2.arena.flat, model.arena.flat)


def test_train_and_sample_cli_synthetic(tmp_path):
    """python -m smd_b200.train_ncsn / sample_ncsn with a ddpm flagfile, synthetic data, a few steps."""
    cfg = tmp_path / "ddpm-tiny.cfg"
    cfg.write_text(textwrap.dedent(f"""\
        --loss=ddpm
        --sampling=ddpm
        --schedule_type=linear
        --sigma_begin=1e-6
        --sigma_end=0.01
        --num_sigmas=50
        --continuous_noise
        --problem=vae
        --ema=False
        --nosnapshot_sampling
        --architecture=TransformerDDPM
        --num_layers=1
        --num_mlp_layers=1
        --data_shape=32,42
        --batch_size=8
        --learning_rate=1e-3
        --max_steps=6
        --snapshot_freq=3
        --logging_freq=2
        --synthetic
        --synthetic_examples=64
        --model_dir={tmp_path / 'run'}
        """))
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "smd_b200.train_ncsn", f"--flagfile={cfg}"], capture_output=True,
                       text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    from smd_b200 import checkpoints
    names = checkpoints.list_checkpoints(str(tmp_path / "run"))
    # numbered by sampling_step like train_ncsn.py:382,395-399 (snapshots at steps 3 and 6; --max_steps ends the run)
    assert names == ["checkpoint_0", "checkpoint_1"], names
    r = subprocess.run([sys.executable, "-m", "smd_b200.sample_ncsn", f"--flagfile={cfg}", "--sample_size=16"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    import pickle
    out = tmp_path / "run" / "samples" / "ncsn"
    gen = pickle.load(open(out / "generated.pkl", "rb"))
    coll = pickle.load(open(out / "collection.pkl", "rb"))
    real = pickle.load(open(out / "real.pkl", "rb"))
    assert gen.shape == (16, 32, 42) and real.shape == (16, 32, 42) and coll.shape == (41, 16, 32, 42)
    assert np.isfinite(gen).all()
    # --infill (sample_ncsn.py:408-427): the first and last 8 latents of every sequence are held fixed
    r = subprocess.run([sys.executable, "-m", "smd_b200.sample_ncsn", f"--flagfile={cfg}", "--sample_size=8", "--infill",
                        "--sampling_dir=infill"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = tmp_path / "run" / "infill" / "ncsn"
    gen = pickle.load(open(out / "generated.pkl", "rb"))
    real = pickle.load(open(out / "real.pkl", "rb"))
    assert gen.shape == (8, 32, 42) and np.isfinite(gen).all()
    np.testing.assert_allclose(gen[:, :8], real[:, :8], rtol=0, atol=1e-5)
    np.testing.assert_allclose(gen[:, -8:], real[:, -8:], rtol=0, atol=1e-5)
    assert np.abs(gen[:, 8:-8] - real[:, 8:-8]).max() > 1e-3
    # --interpolate (sample_ncsn.py:429-438): 9 latent mixtures between each example and its neighbour
    r = subprocess.run([sys.executable, "-m", "smd_b200.sample_ncsn", f"--flagfile={cfg}", "--sample_size=4",
                        "--interpolate", "--sampling_dir=interp"], capture_output=True, text=True, timeout=600, env=env,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = tmp_path / "run" / "interp" / "ncsn"
    gen = pickle.load(open(out / "generated.pkl", "rb"))
    coll = pickle.load(open(out / "collection.pkl", "rb"))
    assert gen.shape == (9, 4, 32, 42) and coll.shape == (9, 41, 4, 32, 42) and np.isfinite(gen).all()
    assert np.abs(gen[0] - gen[8]).max() > 1e-3          # the two end points decode different latents
    # neighbouring mixtures stay closer to each other than the end points do (same sampling key for every chain)
    assert np.abs(gen[4] - gen[5]).mean() < np.abs(gen[0] - gen[8]).mean()

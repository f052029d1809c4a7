"""CPU tests of the host-side mirror: flag surface / flagfiles, LR schedule, early stopping, checkpoints,
schedule parity with the oracle, and the data-parallel plumbing over gloo with world_size 2."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# The reference's flagfiles, restated (configs/ddpm-base.cfg:1-15 + configs/ddpm-mel-32seq-512.cfg:1-10): the
# GPU box has no /root/reference, so the text lives here; key/values are the upstream ones.
DDPM_BASE = """\
--loss=ddpm
--learning_rate=1e-3
--batch_size=64
--sigma_begin=1e-6
--sigma_end=0.01
--num_sigmas=1000
--schedule_type=linear
--ema=False
--sampling=ddpm
--continuous_noise
--normalize
--problem=vae
--nosnapshot_sampling
--epochs=100
--max_steps=500000
"""
DDPM_MEL = """\
--flagfile={base}
--architecture=TransformerDDPM
--num_layers=6
--num_heads=8
--num_mlp_layers=2
--mlp_dims=2048
--data_shape=32,512
--dataset=/tmp/nonexistent
--slice_ckpt=./checkpoints/slice-mel-512.pkl
--model_dir=./save/ddpm-mel
"""


def test_flagfiles_parse_like_the_reference(tmp_path):
    base = tmp_path / "ddpm-base.cfg"
    base.write_text(DDPM_BASE)
    mel = tmp_path / "ddpm-mel-32seq-512.cfg"
    mel.write_text(DDPM_MEL.format(base=base))
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        from absl import flags
        from smd_b200 import sample_ncsn
        F = flags.FLAGS
        F(['prog', '--flagfile={mel}', '--sample_size=250'])
        assert F.loss == 'ddpm' and F.sampling == 'ddpm' and F.schedule_type == 'linear'
        assert F.num_sigmas == 1000 and F.ema is False and F.snapshot_sampling is False
        assert F.continuous_noise is True and F.batch_size == 64 and F.max_steps == 500000
        assert F.data_shape == ['32', '512'] and F.num_heads == 8 and F.sample_size == 250
        assert abs(F.learning_rate - 1e-3) < 1e-12 and abs(F.sigma_begin - 1e-6) < 1e-15
        # defaults of flags the cfg does not set (train_ncsn.py:48-128)
        assert F.grad_clip == 1.0 and F.lr_gamma == 0.98 and F.lr_schedule_interval == 10000 and F.mu == 0.999
        assert F.seed == 0 and F.sample_seed == 1 and F.checkpoints_to_keep == 50
        from smd_b200 import train_ncsn
        assert train_ncsn.lr_at(1) == 1e-3 and train_ncsn.lr_at(10000) == 1e-3
        assert abs(train_ncsn.lr_at(10001) - 1e-3 * 0.98) < 1e-15 and abs(train_ncsn.lr_at(25000) - 1e-3 * 0.98 ** 2) < 1e-15
        print('ok')
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_schedule_matches_oracle_and_rejects_unknown():
    from oracle import ddpm_oracle as O
    from smd_b200 import ebm_utils
    for kind, (a, b, n) in {"linear": (1e-6, 0.01, 1000), "geometric": (1.0, 0.01, 15), "fibonacci": (0, 0, 12)}.items():
        np.testing.assert_array_equal(ebm_utils.create_noise_schedule(a or 1, b or 1, n, kind),
                                      O.create_noise_schedule(a or 1, b or 1, n, kind))
    with pytest.raises(ValueError):
        ebm_utils.create_noise_schedule(1, 2, 3, "cosine")


def test_unknown_dispatch_values_raise_like_the_reference():
    from smd_b200 import train_ncsn
    with pytest.raises(ValueError):
        train_ncsn.sample(None, None, np.array([0, 1], np.uint32), (32, 42), sampling="hmc")
    from smd_b200.losses import reduce_fn
    import torch
    with pytest.raises(ValueError):
        reduce_fn(torch.ones(2), "median")


def test_early_stopping_and_collate():
    from smd_b200 import ebm_utils, train_utils
    es = train_utils.EarlyStopping(patience=1)
    improved, es = es.update(1.0)
    assert improved and es.best_metric == 1.0
    improved, es = es.update(1.5)
    assert not improved and not es.should_stop and es.patience_count == 1
    improved, es = es.update(1.2)
    assert not improved and es.should_stop
    m = ebm_utils.collate_sampling_metrics(np.arange(4 * 3 * 1, dtype=np.float32).reshape(4, 3, 1))
    assert len(m) == 3 and m[1][0] == {"slope": 1.0, "step": 4.0, "alpha": 7.0, "noise": 10.0}


DP_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
import torch.distributed as dist
from smd_b200 import parallel
parallel.init_from_env("gloo")
w, r = parallel.world_size(), parallel.rank()
assert w == 2
rng = np.random.default_rng(0)
X = rng.standard_normal((8, 5)).astype(np.float32)          # global batch, same on every rank
theta = torch.from_numpy(rng.standard_normal(5).astype(np.float32))
def grad_sum(x):   # gradient of sum_i (x_i . theta)^2 over the rows given (un-normalised)
    x = torch.from_numpy(x)
    return (2 * (x @ theta)[:, None] * x).sum(0)
local = parallel.shard_rows(X)
assert local.shape[0] == parallel.shard_size(8) == 4
np.testing.assert_array_equal(local, X[r * 4:(r + 1) * 4])
g = grad_sum(local) / 8.0                                   # pre-scaled by 1/global_batch (smd_ddpm_grads contract)
loss_sum = torch.tensor([float(((torch.from_numpy(local) @ theta) ** 2).sum())])
parallel.all_reduce_sum_(g); parallel.all_reduce_sum_(loss_sum)
ref = grad_sum(X) / 8.0
assert torch.allclose(g, ref, atol=1e-5), (g, ref)
assert abs(float(loss_sum) / 8.0 - float(((torch.from_numpy(X) @ theta) ** 2).mean())) < 1e-4
rows = parallel.gather_rows(torch.full((3, 2), float(r)))
assert rows.shape == (6, 2) and float(rows[0, 0]) == 0.0 and float(rows[5, 0]) == 1.0
try:
    parallel.shard_size(7)
    raise SystemExit("expected ValueError")
except ValueError:
    pass
parallel.shutdown()
sys.stdout.write(f"rank{r}-ok\n"); sys.stdout.flush()
"""


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(DP_WORKER.replace("{root!r}", repr(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "rank0-ok" in r.stdout and "rank1-ok" in r.stdout


def test_bench_reference_arm_exits_cleanly_on_nonzero_rank():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""

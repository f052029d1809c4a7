"""oracle/layout.py (pure numpy, used by the CPU oracle and bench.py's reference arm) names, orders and initialises
the parameters exactly like the product's arena (smd_tensor_info / Engine.init_params) -- without importing it."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import layout as LY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    dict(arch="TransformerDDPM", num_layers=6, num_heads=8, num_mlp_layers=2, mlp_dims=2048, channels=42),
    dict(arch="TransformerDDPM", num_layers=8, num_heads=16, num_mlp_layers=3, mlp_dims=2048, channels=42),
    dict(arch="TransformerDDPM4", num_layers=6, num_heads=8, num_mlp_layers=2, mlp_dims=2048, channels=146),
    dict(arch="DenseDDPM", num_layers=6, mlp_dims=2048, channels=512),
]


@pytest.mark.parametrize("kw", CASES)
def test_layout_matches_product_arena(lib, kw):
    from smd_b200 import Engine, ModelConfig
    eng = Engine(ModelConfig(**kw), max_batch=2)
    assert [(n, tuple(s)) for n, _, s in eng.layout] == LY.param_shapes(**kw)
    assert LY.num_params(**kw) == eng.num_params
    assert abs(LY.flops_fwd_per_sample(**kw) - eng.cfg.flops_fwd_per_sample()) < 1.0


def test_reported_parameter_counts():
    assert LY.num_params(**CASES[0]) == 25_579_946
    assert LY.num_params(**CASES[1]) == 37_596_842
    assert LY.num_params(**CASES[3]) == 67_088_896


def test_init_matches_product_init(lib):
    from smd_b200 import Engine, ModelConfig
    kw = dict(arch="TransformerDDPM", num_layers=1, num_heads=8, num_mlp_layers=1, mlp_dims=256, channels=42)
    eng = Engine(ModelConfig(**kw), max_batch=2)
    for perturb in (0.0, 0.02):
        ref = eng.flat_to_dict(eng.init_params(seed=3, perturb=perturb))
        got = LY.init_params(seed=3, perturb=perturb, **kw)
        assert list(ref) == list(got)
        for k in ref:
            np.testing.assert_array_equal(ref[k], got[k])


def test_reference_arm_runs_without_product_code():
    """bench.py --impl reference: oracle only (asserted inside: no smd_b200 module, no libsmd mapping), same workload
    string and global batch as the GPU arm."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                        "--batch", "4"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["config"]["workload"].startswith("train ddpm-mel-32seq-512.cfg") and line["config"]["sample_batch"] == 4
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0

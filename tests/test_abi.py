"""The C-ABI library loads on a CPU-only box and exports every symbol include/smd.h declares."""
import ctypes
import os
import re

import pytest

from smd_b200 import Engine, ModelConfig
from smd_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "smd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 20
    raw = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"libsmd.so does not export {n}"
    assert set(L.EXPORTED_SYMBOLS) == set(names), set(L.EXPORTED_SYMBOLS) ^ set(names)
    assert lib.smd_version() >= 100


def test_plan_layout_and_errors(lib):
    eng = Engine(ModelConfig(channels=146), max_batch=8)
    names = [n for n, _, _ in eng.layout]
    assert names[0] == "in.kernel" and names[-1] == "out.bias"
    assert all(off % 4 == 0 for _, off, _ in eng.layout)          # 16-byte aligned tensors
    assert eng.workspace_bytes > 0
    with pytest.raises(ValueError):
        Engine(ModelConfig(arch="ToyDDPM"), 4)                     # unknown architecture (train_ncsn.py:194)
    with pytest.raises(ValueError):
        Engine(ModelConfig(seq_len=16), 4)                         # unsupported on the CUDA path
    with pytest.raises(ValueError):
        Engine(ModelConfig(mlp_dims=1000), 4)


def test_transformer_ddpm4_alias(lib):
    a = Engine(ModelConfig(arch="TransformerDDPM4", channels=146), 4)   # configs/ddpm-multi-32seq-512.cfg:1
    b = Engine(ModelConfig(arch="TransformerDDPM", channels=146), 4)
    assert a.layout == b.layout


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    eng = Engine(ModelConfig(num_layers=1, num_mlp_layers=1), 2)
    with pytest.raises(L.SmdError):
        eng.set_params(eng.init_params(0))


def test_host_threefry_split_matches_oracle(lib):
    import numpy as np
    from oracle import threefry as tf
    key = (ctypes.c_uint32 * 2)(0, 0)
    out = (ctypes.c_uint32 * 6)()
    assert lib.smd_threefry_split(key, 3, out) == 0
    np.testing.assert_array_equal(np.array(list(out), np.uint32).reshape(3, 2), tf.split(tf.prng_key(0), 3))


def test_tail_gradient_slice_is_the_film_tail_and_output_layer(lib):
    """smd_grads_tail_range: the contiguous arena slice that is final after the tail backward (data-parallel overlap)."""
    for cfg in (ModelConfig(channels=42), ModelConfig(arch="DenseDDPM", num_layers=3, channels=512)):
        eng = Engine(cfg, max_batch=4, training=True)
        first, count = eng.grads_tail_range()
        offs = {n: off for n, off, _ in eng.layout}
        assert first == offs["k0.film.d1.kernel"] and first + count == eng.arena_floats
        inside = [n for n, off, _ in eng.layout if off >= first]
        assert all(n.startswith(("k", "out_ln.", "out.")) for n in inside) and "out.bias" in inside
        assert not any(n.startswith(("k", "out")) for n, off, _ in eng.layout if off < first)
        assert count > 0.8 * eng.arena_floats           # ~85% of the parameters live in the FiLM'd tail


def test_plain_c_client(lib, tmp_path):
    """include/smd.h compiles as C11 and a C program drives the plan / layout entry points of libsmd.so."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    exe = tmp_path / "abi_client"
    libdir = os.path.dirname(L.LIB_PATH)
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c_client", "abi_client.c"), "-o", str(exe),
                        "-L", libdir, "-l:libsmd.so", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "params=25579946" in r.stdout and "tensors=" in r.stdout

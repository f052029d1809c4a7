"""Stage-by-stage comparison of the CUDA forward (training-save mode keeps every intermediate) against the
bf16-emulating oracle.  Prints rel-L2 per workspace region so a wrong kernel is located in one GPU call."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ddpm_oracle as O  # noqa: E402
from smd_b200 import Engine, ModelConfig  # noqa: E402
from smd_b200 import lib as L  # noqa: E402
from tests.util import oracle_kwargs, params_torch, rel_l2  # noqa: E402


def fetch(eng, name, shape, dtype):
    """Copy a named workspace region into a fresh torch tensor (device -> device, via cuda-python)."""
    from cuda import cudart
    ptr = C.c_void_p()
    nbytes = C.c_size_t()
    L.check(eng.lib.smd_debug_buffer(eng._plan, name.encode(), C.byref(ptr), C.byref(nbytes)))
    n = int(np.prod(shape))
    esz = 2 if dtype == torch.bfloat16 else 4
    assert n * esz <= nbytes.value, (name, n * esz, nbytes.value)
    out = torch.empty(n, dtype=dtype, device="cuda")
    torch.cuda.synchronize()
    (err,) = cudart.cudaMemcpy(out.data_ptr(), ptr.value, n * esz, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert int(err) == 0, err
    return out.reshape(shape).float().cpu()


def main():
    kw = dict(num_layers=2, num_heads=8, num_mlp_layers=2, channels=42)
    batch = 4
    for cg in (1, 2):
        eng = Engine(ModelConfig(**kw), max_batch=batch, cta_group=cg, training=True)
        flat = eng.init_params(seed=1, perturb=0.02)
        eng.set_params(flat)
        rng = np.random.default_rng(7)
        x = rng.uniform(-1, 1, (batch, 32, 42)).astype(np.float32)
        t = rng.uniform(0.05, 1.0, (batch,)).astype(np.float32)
        xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
        y = torch.empty_like(xd)
        L.check(eng.lib.smd_debug_forward_save(eng._plan, eng.params.data_ptr(), xd.data_ptr(), td.data_ptr(), batch,
                                               y.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        trace = {}
        p = params_torch(eng, flat)
        ref = O.transformer_ddpm(p, torch.from_numpy(x), torch.from_numpy(t), emulate_bf16=True, trace=trace,
                                 **oracle_kwargs(eng.cfg))
        M, Md = batch * 32, 2048
        print(f"--- cta_group {cg}: final rel-L2 {rel_l2(y, ref):.3e}")
        shapes = {"t.h": ((M, 128), torch.float32), "t.a1_": ((M, 128), torch.bfloat16),
                  "t.a2_": ((M, 128), torch.bfloat16), "t.hpre": ((M, Md), torch.bfloat16),
                  "t.hid": ((M, Md), torch.bfloat16), "t.a_post": ((M, 128), torch.bfloat16),
                  "t.u": ((M, Md), torch.float32), "t.act_out": ((M, Md), torch.bfloat16)}
        for name, val in trace.items():
            if name.startswith("ss"):
                k = int(name[2:])
                got = fetch(eng, "ss", (eng.cfg.num_mlp_layers, batch, 2 * Md), torch.float32)[k]
                print(f"  {name:10s} rel-L2 {rel_l2(got, val):.3e}")
                continue
            key = max((s for s in shapes if name.startswith(s)), key=len)
            shape, dt = shapes[key]
            got = fetch(eng, name, shape, dt)
            print(f"  {name:10s} rel-L2 {rel_l2(got, val.reshape(shape)):.3e}")


if __name__ == "__main__":
    main()

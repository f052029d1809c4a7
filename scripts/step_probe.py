"""Where does the train step go?  CUDA-event timings of the pieces (forward-only loss, gradients, clip+Adam, whole
step) for the headline model and for variants with one trunk layer / one res-block, plus the whole step replayed from
a CUDA graph (torch.cuda.graph capture of the library calls).  Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smd_b200 import Engine, ModelConfig  # noqa: E402


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def probe(L, K, B=128, graph=True):
    cfg = ModelConfig(num_layers=L, num_heads=8, num_mlp_layers=K, channels=42)
    eng = Engine(cfg, max_batch=B, cta_group=2, training=True)
    eng.set_params(eng.init_params(seed=1))
    eng.init_train_state()
    eng.objective_setup(np.linspace(1e-6, 0.01, 1000, dtype=np.float32))
    x = torch.rand(B, 32, 42, device="cuda") * 2 - 1
    used, eps = eng.draws((0, 1), B)
    out = {"L": L, "K": K, "B": B}
    out["fwd_loss_ms"] = timeit(lambda: eng.ddpm_loss(x, used, eps))
    out["grads_ms"] = timeit(lambda: eng.compute_grads(x, used, eps))
    out["adam_ms"] = timeit(lambda: eng.apply_grads(1e-3))
    l0 = eng.launch_count()
    eng.train_step(x, used, eps, lr=1e-3)
    out["launches_per_step"] = eng.launch_count() - l0
    out["step_ms"] = timeit(lambda: eng.train_step(x, used, eps, lr=1e-3))
    if graph:
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                eng.train_step(x, used, eps, lr=1e-3)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s, capture_error_mode="relaxed"):
                    eng.compute_grads(x, used, eps)
                    eng.apply_grads(1e-3)
                torch.cuda.synchronize()
                out["step_graph_ms"] = timeit(g.replay)
        except Exception as e:  # noqa: BLE001
            out["step_graph_error"] = repr(e)[:300]
    return out


if __name__ == "__main__":
    torch.cuda.set_stream(torch.cuda.Stream())
    res = [probe(6, 2), probe(1, 2, graph=False), probe(6, 1, graph=False), probe(1, 1, graph=False)]
    print(json.dumps(res, indent=1))

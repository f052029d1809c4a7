"""Timings of the BASELINE.json configs that bench.py does not print (single GPU, synthetic data, random-init weights):
  * cfg3 as the user runs it: the FULL 1000-step reverse chain over 1000 samples (smd_ddpm_sample, CUDA-graph replay);
  * cfg4's per-GPU work: ddpm-mel-32seq-512-large.cfg (L8 H16 K3) train step at batch 128;
  * cfg5: ddpm-multi-32seq-512.cfg (C = 146) reverse step over 1000 samples.
Writes one JSON object to stdout.  Usage (on a GPU box): python scripts/extra_configs.py > gpurun_out/extra.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smd_b200 import Engine, ModelConfig  # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    torch.cuda.set_stream(torch.cuda.Stream())
    betas = np.linspace(1e-6, 0.01, 1000, dtype=np.float32)
    out = {}
    # ---- full chain, base model
    cfg = ModelConfig(num_layers=6, num_heads=8, num_mlp_layers=2, channels=42)
    eng = Engine(cfg, max_batch=1000, cta_group=2)
    eng.set_params(eng.init_params(seed=1))
    eng.sampler_setup(betas, key=(0, 5))
    x = torch.randn(1000, 32, 42, device="cuda")
    ms = timed(lambda: eng.sample(x.clone(), steps=1000, use_graph=True), 2, 1)
    out["cfg3_full_chain_1000x1000"] = {"ms": ms, "sample_steps_per_s": 1000 * 1000 / (ms / 1e3),
                                        "tflops": cfg.flops_fwd_per_sample() * 1e6 / (ms / 1e3) / 1e12}
    del eng
    # ---- multi-track latents, C = 146
    cfg = ModelConfig(num_layers=6, num_heads=8, num_mlp_layers=2, channels=146)
    eng = Engine(cfg, max_batch=1000, cta_group=2)
    eng.set_params(eng.init_params(seed=1))
    eng.sampler_setup(betas, key=(0, 5))
    x = torch.randn(1000, 32, 146, device="cuda")
    ms = timed(lambda: eng.sample(x, steps=1, use_graph=True), 20, 5)
    out["cfg5_multi_c146_reverse_step_1000"] = {"ms": ms, "sample_steps_per_s": 1000 / (ms / 1e3),
                                                "tflops": cfg.flops_fwd_per_sample() * 1000 / (ms / 1e3) / 1e12}
    del eng
    # ---- large model, train step batch 128 (the per-GPU work of cfg4)
    cfg = ModelConfig(num_layers=8, num_heads=16, num_mlp_layers=3, channels=42)
    eng = Engine(cfg, max_batch=128, cta_group=2, training=True)
    eng.set_params(eng.init_params(seed=1))
    eng.init_train_state(ema=False)
    eng.objective_setup(betas)
    x0 = torch.rand(128, 32, 42, device="cuda") * 2 - 1
    used, eps = eng.draws((0, 17), 128)
    ms = timed(lambda: eng.train_step(x0, used, eps, lr=1e-3), 20, 5)
    out["cfg4_large_train_step_batch128"] = {"ms": ms, "sample_steps_per_s": 128 / (ms / 1e3),
                                             "tflops": 3 * cfg.flops_fwd_per_sample() * 128 / (ms / 1e3) / 1e12,
                                             "params": int(eng.arena_floats)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

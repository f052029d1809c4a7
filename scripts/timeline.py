"""Kernel timeline of one train step / one sampling step from CUPTI (torch.profiler): every kernel with its stream,
start and duration, so overlap between the library's streams is visible (ncu serialises launches, this does not).
Usage: python scripts/timeline.py {train|sample} out.json"""
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smd_b200 import Engine, ModelConfig  # noqa: E402


def main():
    mode, out_path = sys.argv[1], sys.argv[2]
    torch.cuda.set_stream(torch.cuda.Stream())
    cfg = ModelConfig(num_layers=6, num_heads=8, num_mlp_layers=2, channels=42)
    betas = np.linspace(1e-6, 0.01, 1000, dtype=np.float32)
    if mode == "train":
        B = 128
        eng = Engine(cfg, max_batch=B, cta_group=2, training=True)
        eng.set_params(eng.init_params(seed=1))
        eng.init_train_state()
        eng.objective_setup(betas)
        x = torch.rand(B, 32, 42, device="cuda") * 2 - 1
        used, eps = eng.draws((0, 1), B)
        step = lambda: eng.train_step(x, used, eps, lr=1e-3)
    else:
        B = 1000
        eng = Engine(cfg, max_batch=B, cta_group=2, training=False)
        eng.set_params(eng.init_params(seed=1))
        eng.sampler_setup(betas, (0, 7))
        x = torch.randn(B, 32, 42, device="cuda")
        t = [999]

        def step():
            eng.reverse_step(x, t[0], x_next=x)
            t[0] -= 1
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    trace = out_path + ".trace.json"
    prof.export_chrome_trace(trace)
    tr = json.load(open(trace))
    rows = [{"name": t_["name"][:90], "ts": t_["ts"], "dur": t_["dur"], "stream": t_.get("args", {}).get("stream"),
             "cat": t_.get("cat")}
            for t_ in tr.get("traceEvents", []) if t_.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    rows.sort(key=lambda r: r["ts"])
    t0 = rows[0]["ts"] if rows else 0
    for r in rows:
        r["ts"] = round(r["ts"] - t0, 3)
    json.dump(rows, open(out_path, "w"), indent=0)
    os.remove(trace)
    print(mode, "kernels", len(rows), "span_us", rows[-1]["ts"] + rows[-1]["dur"] if rows else 0)


if __name__ == "__main__":
    main()

"""Benchmark of the DDPM hot path (BASELINE.json metric: denoising steps/sec on (B,32,512)-derived latents).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path (oracle)

Workloads (config.workload):
  train   configs[1]: ddpm-mel-32seq-512.cfg (TransformerDDPM L6/H8/K2/M2048, C=42 after slice-mel-512),
          batch 128 per GPU, one optimizer step = q_sample + forward + backward + NCCL all-reduce + clip + Adam.
  sample  configs[2]: same model, one reverse-diffusion step over 1000 samples per GPU (model call + update).
A "step" is one pass of that hot path over one batch; value = samples x steps / s over all GPUs (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoising steps/sec (Bx32x512 latents)"
UNIT = "sample-steps/s"


def model_config(workload: str):
    from smd_b200 import ModelConfig
    return ModelConfig(arch="TransformerDDPM", num_layers=6, num_heads=8, num_mlp_layers=2, mlp_dims=2048,
                       seq_len=32, channels=42)


def synthetic_batch(batch: int, seed: int):
    """(B,32,512) N(0,1) 'MusicVAE' latents -> slice 42 dims -> min/max normalise to [-1,1] (input_pipeline.py:36-48)."""
    rng = np.random.default_rng(seed)
    raw = rng.standard_normal((batch, 32, 512)).astype(np.float32)
    idx = np.sort(np.random.default_rng(1234).choice(512, 42, replace=False))
    x = np.ascontiguousarray(raw[..., idx])
    lo, hi = x.min(), x.max()
    return np.ascontiguousarray((2.0 * (x - lo) / (hi - lo) - 1.0).astype(np.float32))


class ClockSampler(threading.Thread):
    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.idx = device_index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------- CPU arm
def host_threads() -> int:
    """Threads the CPU arm may use: the scheduler affinity mask, clipped by a cgroup CPU quota if there is one and by
    32 (a batch-16 sample of this model stops scaling -- and with an over-reported core count collapses -- beyond)."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(np.ceil(int(quota) / int(period)))))
    except (OSError, ValueError):
        pass
    return max(1, min(n, 32))


class CpuTrainStep:
    """Oracle (torch CPU fp32 restatement of train_ncsn.py:260-288) on a bounded sample of the train workload."""

    def __init__(self, sample_batch: int, threads: int):
        from oracle import ddpm_oracle as O
        from smd_b200 import Engine
        self.O = O
        torch.set_num_threads(threads)
        cfg = model_config("train")
        eng = Engine(cfg, max_batch=sample_batch)     # layout only; never touches the GPU
        flat = eng.init_params(seed=1)
        self.p = {k: torch.from_numpy(v) for k, v in eng.flat_to_dict(flat).items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(t) for k, t in self.p.items()}
        self.kw = dict(num_layers=6, num_heads=8, num_mlp_layers=2, mlp_dims=2048)
        self.n = 0
        self.set_batch(sample_batch)

    def set_batch(self, sample_batch: int):
        O = self.O
        self.batch = sample_batch
        self.x0 = torch.from_numpy(synthetic_batch(sample_batch, 0))
        rng = np.random.default_rng(2)
        self.eps = torch.from_numpy(rng.standard_normal(tuple(self.x0.shape)).astype(np.float32))
        ap = O.alphas_prod_with_one(O.create_noise_schedule(1e-6, 0.01, 1000, "linear"))
        self.used = torch.from_numpy(ap[rng.integers(1, 1001, sample_batch) - 1])

    def step(self) -> float:
        t0 = time.perf_counter()
        (self.p, self.m, self.v), _, _, _ = self.O.train_step("TransformerDDPM", self.p, self.m, self.v, self.n, self.x0,
                                                              self.used, self.eps, 1e-3, model_kw=self.kw)
        self.n += 1
        return time.perf_counter() - t0


def cpu_train_step_rate(sample_batch: int, min_seconds: float, max_steps: int, threads: int):
    """(sample-steps/s, steps, seconds) of the CPU restatement: one untimed warm-up step, then steps until
    `min_seconds` have passed (at most `max_steps`); if the warm-up shows a step would blow the budget the sample
    shrinks (batch 16 -> 4 -> 1)."""
    job = CpuTrainStep(sample_batch, threads)
    w = job.step()
    while w > max(3.0, min_seconds) and job.batch > 1:
        job.set_batch(max(1, job.batch // 4))
        w = job.step()
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < min_seconds):
        job.step()
        n += 1
    dt = time.perf_counter() - t0
    return job.batch * n / dt, n, dt, job.batch


def cpu_sample_step_rate(n_samples: int, min_seconds: float, max_steps: int, threads: int):
    """Oracle reverse-diffusion steps (ebm_utils.py:327-397) on a bounded sample of the sampling workload."""
    from oracle import ddpm_oracle as O
    from smd_b200 import Engine
    torch.set_num_threads(threads)
    cfg = model_config("sample")
    eng = Engine(cfg, max_batch=n_samples)     # layout only; never touches the GPU
    p = {k: torch.from_numpy(v) for k, v in eng.flat_to_dict(eng.init_params(seed=1)).items()}
    kw = dict(num_layers=6, num_heads=8, num_mlp_layers=2, mlp_dims=2048)
    coef = O.reverse_coefficients(O.create_noise_schedule(1e-6, 0.01, 1000, "linear"))
    rng = np.random.default_rng(3)
    state = torch.from_numpy(rng.standard_normal((n_samples, 32, 42)).astype(np.float32))
    z = torch.from_numpy(rng.standard_normal((n_samples, 32, 42)).astype(np.float32))
    apply_fn = lambda a, c: O.transformer_ddpm(p, a, c.reshape(-1), **kw)
    with torch.no_grad():
        state = O.reverse_step(apply_fn, state, 999, coef, z)[0]      # warm-up
        t0 = time.perf_counter()
        n = 0
        while n < max_steps and (n == 0 or time.perf_counter() - t0 < min_seconds):
            state = O.reverse_step(apply_fn, state, 998 - n, coef, z)[0]
            n += 1
    dt = time.perf_counter() - t0
    return n_samples * n / dt, n, dt, n_samples


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    budget = 150.0                      # seconds for the W + K steps: each step is a bounded sample of the workload
    job = CpuTrainStep(16, threads)
    w = job.step()                      # sizing probe (also pays the first-call costs); not reported
    per_step = budget / max(1, args.warmup + args.steps)
    while w > per_step and job.batch > 1:
        job.set_batch(max(1, job.batch // 2))
        w = job.step()
    for _ in range(args.warmup):
        job.step()
    times = [job.step() for _ in range(args.steps)]
    sample_batch = job.batch
    ms = 1e3 * float(np.mean(times))
    value = sample_batch / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "train ddpm-mel-32seq-512.cfg (TransformerDDPM L6 H8 K2 M2048 C42)",
                       "global_batch": sample_batch, "note": "CPU restatement of the reference path (JAX 0.2.8/flax "
                       f"0.3.0 not installable); each step is a batch-{sample_batch} sample of the batch-128 train step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"batch {sample_batch} optimizer steps (fwd+bwd+clip+Adam), torch CPU fp32, "
                                       f"{threads} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def time_dominant_gemm(eng, M_tokens: int, cta_group: int, iters: int = 20):
    """Average duration of the dominant kernel (2048x2048 res-block GEMM over M tokens) alone, L2 flushed."""
    from smd_b200 import lib as L
    lib = eng.lib
    A = torch.randn(M_tokens, 2048, device="cuda").to(torch.bfloat16)
    B = torch.randn(2048, 2048, device="cuda").to(torch.bfloat16)
    out = torch.empty(M_tokens, 2048, device="cuda")
    bias = torch.zeros(2048, device="cuda")
    stats = torch.zeros(M_tokens, 2, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    total = 0.0
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.smd_gemm_bf16(A.data_ptr(), B.data_ptr(), M_tokens, 2048, 2048, 0, 0, 256, cta_group,
                                  bias.data_ptr(), None, 0, out.data_ptr(), None, stats.data_ptr(), None, None, st))
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            total += e0.elapsed_time(e1)
    return total / iters  # ms


def run_gpu(args):
    import torch.distributed as dist
    from smd_b200 import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cfg = model_config(args.workload)
    betas = np.linspace(1e-6, 0.01, 1000, dtype=np.float32)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev = torch.device(f"cuda:{local}")
    tpeak, tsust, hbm, peak_src = peaks()

    if args.workload == "train":
        B = args.batch or 128
        eng = Engine(cfg, max_batch=B, cta_group=args.cta_group, training=True)
        eng.set_params(eng.init_params(seed=1))
        eng.init_train_state(ema=False)
        eng.objective_setup(betas)
        x_host = torch.from_numpy(synthetic_batch(B, 100 + rank)).pin_memory()
        x_dev = x_host.to(dev, non_blocking=True)
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
        used, eps = eng.draws((0, 17 + rank), B)

        def step_resident(i):
            eng.train_step(x_dev, used, eps, lr=1e-3, world_size=world)

        def step_e2e(i):
            xb = x_host.to(dev, non_blocking=True)                      # H2D of this step's batch (pinned)
            u, e = eng.draws((i, 17 + rank), B)                          # device threefry draws (losses.py:270-294)
            loss, _ = eng.train_step(xb, u, e, lr=1e-3, world_size=world)
            loss_host.copy_(loss, non_blocking=True)                     # D2H of the step's loss
            torch.cuda.current_stream().synchronize()

        units = B
        flops_step = 3.0 * cfg.flops_fwd_per_sample() * B
        h2d, d2h = x_host.numel() * 4, 4
        wl = f"train ddpm-mel-32seq-512.cfg (TransformerDDPM L6 H8 K2 M2048 C42), batch {B}/GPU"
        m_tokens = B * 32
    else:
        N = args.batch or 1000
        eng = Engine(cfg, max_batch=N, cta_group=args.cta_group, training=False)
        eng.set_params(eng.init_params(seed=1))
        eng.sampler_setup(betas, key=(0, 5 + rank))
        x_host = torch.from_numpy(np.random.default_rng(rank).standard_normal((N, 32, 42)).astype(np.float32)).pin_memory()
        x_dev = x_host.to(dev)
        out_host = torch.empty((N, 32, 42), dtype=torch.float32).pin_memory()

        def step_resident(i):
            eng.sample(x_dev, steps=1, use_graph=True)

        def step_e2e(i):
            xb = x_host.to(dev, non_blocking=True)
            eng.sample(xb, steps=1, use_graph=False)
            out_host.copy_(xb, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        units = N
        flops_step = cfg.flops_fwd_per_sample() * N
        h2d = d2h = x_host.numel() * 4
        wl = f"sample ddpm-mel-32seq-512.cfg (TransformerDDPM L6 H8 K2 M2048 C42), {N} samples/GPU, 1 reverse step"
        m_tokens = N * 32

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = eng.launch_count()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps, eng.launch_count() - launches0

    clocks = ClockSampler(local)
    clocks.start()
    ms_step, launches = timed(step_resident, args.steps, args.warmup)
    clocks.stop_flag.set()
    clocks.join(timeout=2)
    ms_e2e, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2))

    value = units * world / (ms_step / 1e3)
    e2e_value = units * world / (ms_e2e / 1e3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl, "global_batch": units * world, "parallelism": f"dp{world}",
                       "cta_group": args.cta_group, "steps_per_s": 1e3 / ms_step,
                       "step_tflops": flops_step * world / (ms_step / 1e3) / 1e12,
                       "step_frac_of_sustained_peak": flops_step / (ms_step / 1e3) / 1e12 / tsust,
                       "l2": "per-step working set (params+grads+Adam ~400 MB, activations ~1 GB) >> 126 MB L2; no flush",
                       "precision": "bf16 tensor-core operands, fp32 accumulate / master weights / LN / softmax / Adam"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches)}
    if rank == 0:
        g_ms = time_dominant_gemm(eng, m_tokens, args.cta_group)
        gflop = 2.0 * m_tokens * 2048 * 2048
        ach = gflop / (g_ms / 1e3) / 1e12
        traffic = None      # DRAM bytes per launch from the committed `ncu --set full` capture of this same kernel/shape
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_dominant_gemm_ncu.json"))).get(str(m_tokens))
            if cap and args.cta_group == 2:
                traffic = cap["traffic_bytes"]
        except (OSError, ValueError, KeyError):
            traffic = None
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                            "traffic": traffic, "algorithmic_bytes": m_tokens * 2048 * 6 + 2048 * 2048 * 2 + m_tokens * 8 + 8192,
                            "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({peak_src}, burst: kernel timed alone)",
                            "kernel": f"gemm_bf16_tcgen05_kernel<{args.cta_group}> [{m_tokens}x2048x2048] res-block GEMM "
                                      "+ bias + row-stat epilogue", "ms_per_launch": g_ms}
        if not args.no_cpu:
            threads = host_threads()
            if args.workload == "train":
                rate, n, dt, sb = cpu_train_step_rate(16, 10.0, 8, threads)
                what = f"{n} optimizer steps at batch {sb} (of the batch-128 train step)"
            else:
                rate, n, dt, sb = cpu_sample_step_rate(32, 10.0, 8, threads)
                what = f"{n} reverse-diffusion steps over {sb} samples (of the 1000-sample step)"
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{what} in {dt:.1f}s, torch CPU fp32 restatement of the reference path "
                                              f"(JAX unavailable), {threads} threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "sample"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--cta-group", dest="cta_group", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

"""Noise schedule and the DDPM reverse sampler with the reference signatures (utils/ebm_utils.py)."""
from __future__ import annotations

import numpy as np
import torch


def _linspace_f32(start, stop, num):
    f32 = np.float32
    start, stop = f32(start), f32(stop)
    if num == 1:
        return np.asarray([start], f32)
    delta = f32(stop - start) / f32(num - 1)
    return (start + np.arange(num, dtype=f32) * delta).astype(f32)


def create_noise_schedule(sigma_begin=1, sigma_end=1e-2, L=10, schedule="geometric"):
    """utils/ebm_utils.py:62-86 (float32, like jnp)."""
    if schedule == "geometric":
        return np.exp(_linspace_f32(np.log(np.float32(sigma_begin)), np.log(np.float32(sigma_end)), L)).astype(np.float32)
    if schedule == "linear":
        return _linspace_f32(sigma_begin, sigma_end, L)
    if schedule == "fibonacci":
        s = [1e-6, 2e-6]
        for _ in range(L - 2):
            s.append(s[-1] + s[-2])
        return np.asarray(s, np.float32)
    raise ValueError(f"Unsupported schedule: {schedule}")


def diffusion_dynamics(rng, model, betas, init, epsilon, T, denoise, infill=False, infill_samples=None,
                       infill_masks=None, shard=None):
    """utils/ebm_utils.py:274-405.  Returns (state, collection (41, N, *shape), ld_metrics (4, len(betas), 1)).

    epsilon / T / denoise are null parameters upstream too.  The whole chain runs on the GPU (CUDA-graph replay of
    one reverse step); the noise of every step comes from `rng` with the reference's three splits per step."""
    from .nn import _as_device_f32
    del epsilon, T, denoise
    x = _as_device_f32(init).clone()
    n = x.shape[0]
    eng = model.engine(n)
    betas = np.asarray(betas, np.float32)
    eng.sampler_setup(betas, key=(int(rng[0]), int(rng[1])))
    # shard = (first_row, total_rows): `init` holds this rank's rows of a global batch; the per-step noise is the same
    # slice of the global threefry stream and the batch-mean metrics are reduced over ranks below
    eng.set_sampler_shard(*(shard if shard is not None else (0, 0)))
    ix = im = None
    collection = torch.zeros((41,) + tuple(x.shape), dtype=torch.float32, device=x.device)
    if infill:
        ix = _as_device_f32(infill_samples)
        im = _as_device_f32(infill_masks)
        # utils/ebm_utils.py:321-323,398: the merged start only goes to collection[0]; the scan itself starts from the
        # raw `init` (the first step's blend overwrites the masked entries anyway)
        collection[0] = x * (1 - im) + ix * im
    else:
        collection[0] = x
    metrics = torch.zeros((4, len(betas)), dtype=torch.float32, device=x.device)
    eng.sample(x, steps=len(betas), infill_x=ix, infill_mask=im, collection=collection, metrics=metrics, use_graph=True)
    if shard is not None:
        # utils/ebm_utils.py:380-384: the metrics are means over the batch -> one weighted (4, T) all-reduce at the end
        from . import parallel
        if parallel.world_size() > 1:
            w = float(n) / float(shard[1])
            alpha_row = metrics[2].clone()              # alpha_prod is not a batch mean
            metrics.mul_(w)
            parallel.all_reduce_sum_(metrics)
            metrics[2] = alpha_row
    return x, collection, metrics.unsqueeze(2)


def collate_sampling_metrics(ld_metrics):
    """utils/ebm_utils.py:408-428."""
    ld = ld_metrics.detach().cpu().numpy() if isinstance(ld_metrics, torch.Tensor) else np.asarray(ld_metrics)
    _, num_sigmas, num_steps = ld.shape
    out = [[] for _ in range(num_sigmas)]
    for i in range(num_sigmas):
        grad_norm, step_norm, alpha, noise_norm = ld[:, i, :]
        for j in range(num_steps):
            out[i].append({"slope": grad_norm[j], "step": step_norm[j], "alpha": alpha[j], "noise": noise_norm[j]})
    return out

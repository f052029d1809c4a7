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


def _collection_slot_fn(total_steps, collection_steps):
    idx = np.linspace(1, total_steps, collection_steps).astype(np.int32)      # utils/ebm_utils.py:127-129

    def slot(image_idx):
        mask = idx == image_idx
        return int(np.sum(np.arange(len(idx)) * mask) + 1) if mask.any() else -1
    return slot


def annealed_langevin_dynamics(rng, model, sigmas, init, epsilon, T, denoise, infill=False, infill_samples=None,
                               infill_masks=None):
    """utils/ebm_utils.py:89-198 (Song & Ermon): for every noise level T steps x += alpha * score + sqrt(2 alpha) z with
    alpha = epsilon (sigma / sigma_L)^2.  Returns (state, collection (101 + denoise, N, ...), metrics (4, L, T)).

    The network call and the update kernel run on the GPU (smd_forward, smd_langevin_step); the host only walks the
    (sigma, step) loop and the key schedule `rng, step_rng, infill_rng = split(rng, 3)` of the scan body."""
    from . import jrandom as random
    from .nn import _as_device_f32
    x = _as_device_f32(init).clone()
    n = x.shape[0]
    eng = model.engine(n)
    sig = np.asarray(sigmas, np.float32)
    assert len(sig) >= 2
    ix = _as_device_f32(infill_samples) if infill else None
    im = _as_device_f32(infill_masks) if infill else None
    L = len(sig)
    collection = torch.zeros((101 + int(bool(denoise)),) + tuple(x.shape), dtype=torch.float32, device=x.device)
    collection[0] = x * (1 - im) + ix * im if infill else x
    slot_of = _collection_slot_fn(L * T, 100)
    metrics = torch.zeros((L, T, 4), dtype=torch.float32, device=x.device)
    sigma_dev = torch.empty((1,), dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x)
    key = np.asarray(rng, np.uint32)
    for si in range(L):
        s = np.float32(sig[si])
        alpha = np.float32(epsilon) * (s / np.float32(sig[-1])) ** 2              # float32, like the jitted scan
        sigma_dev.copy_(torch.tensor([float(s)]))
        for i in range(T):
            key, step_key, infill_key = random.split(key, 3)
            eng.forward(x, sigma_dev, out=grad)                                 # model(state, sigma)
            slot = slot_of(si * T + i + 1)
            if slot >= collection.shape[0]:
                slot = -1       # repeated linspace indices (fewer than 100 steps in total): XLA's scatter drops the update
            eng.langevin_step(x, grad, float(alpha), float(np.sqrt(np.float32(2) * alpha)), step_key=step_key,
                              infill_x=ix, infill_mask=im, infill_sigma=float(s), infill_key=infill_key, x_next=x,
                              collection_slot=collection[slot] if slot >= 0 else None, metrics4=metrics[si, i])
    if denoise:                                                                   # utils/ebm_utils.py:193-196
        sigma_dev.copy_(torch.tensor([float(sig[-1])]))
        eng.forward(x, sigma_dev, out=grad)
        eng.langevin_step(x, grad, float(np.float32(sig[-1]) ** 2), 0.0, x_next=x, collection_slot=collection[-1])
    return x, collection, metrics.permute(2, 0, 1).contiguous()


def consistent_langevin_dynamics(rng, model, sigmas, init, epsilon, T, denoise=True, infill=False, infill_samples=None,
                                 infill_masks=None):
    """utils/ebm_utils.py:201-271 (Jolicoeur-Martineau et al.): one step per noise level,
    x += alpha * score + beta * sigma_{i+1} * z.  Returns (state, None, metrics (4, L, 1)) -- upstream returns two values
    although its caller unpacks three; the missing collection is returned as None."""
    from . import jrandom as random
    from .nn import _as_device_f32
    del T
    if infill:
        raise NotImplementedError
    x = _as_device_f32(init).clone()
    eng = model.engine(x.shape[0])
    sig = np.asarray(sigmas, np.float32)
    assert len(sig) >= 2
    L = len(sig)
    beta = np.sqrt(np.float32(1) - (np.float32(1) - np.float32(epsilon) / sig[-1] ** 2) ** 2).astype(np.float32)
    metrics = torch.zeros((L, 4), dtype=torch.float32, device=x.device)
    sigma_dev = torch.empty((1,), dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x)
    key = np.asarray(rng, np.uint32)
    for i in range(L):
        key, step_key = random.split(key)
        s = np.float32(sig[i])
        nxt = np.float32(sig[i + 1]) if i < L - 1 else np.float32(0)
        alpha = np.float32(epsilon) * (s / np.float32(sig[-1])) ** 2
        sigma_dev.copy_(torch.tensor([float(s)]))
        eng.forward(x, sigma_dev, out=grad)
        eng.langevin_step(x, grad, float(alpha), float(beta * nxt), step_key=step_key, x_next=x, metrics4=metrics[i])
    if denoise:
        sigma_dev.copy_(torch.tensor([float(sig[-1])]))
        eng.forward(x, sigma_dev, out=grad)
        eng.langevin_step(x, grad, float(np.float32(sig[-1]) ** 2), 0.0, x_next=x)
    return x, None, metrics.t().contiguous().unsqueeze(2)

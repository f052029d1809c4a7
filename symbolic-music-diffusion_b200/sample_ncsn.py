"""Sampling entry point with the reference's flag surface (sample_ncsn.py:51-66 on top of train_ncsn.py:48-128):

  python -m smd_b200.sample_ncsn --flagfile=configs/ddpm-mel-32seq-512.cfg --sample_size=1000 [--synthetic]

Restores the newest checkpoint of --model_dir into a freshly created model (sample_ncsn.py:331-342; like upstream
it samples from optimizer.target, never from the EMA parameters), runs the full reverse chain on the GPU and
writes ncsn/{real,generated,collection}.pkl under --sampling_dir (sample_ncsn.py:453-471).  With torchrun the
sample_size is sharded across ranks (replicated parameters, no per-step exchange) and gathered at the end.
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np
import torch
from absl import app, flags, logging

from smd_b200 import checkpoints, ebm_utils, input_pipeline, jrandom as random, parallel, train_utils
from smd_b200 import train_ncsn  # noqa: F401  (defines the shared training flags)

FLAGS = flags.FLAGS
flags.DEFINE_integer("sample_seed", 1, "PRNG seed of the sampler.")
flags.DEFINE_string("sampling_dir", "samples", "Output directory (relative to --model_dir).")
flags.DEFINE_integer("sample_size", 1000, "Number of samples to draw.")
flags.DEFINE_bool("compute_metrics", False, "Compute evaluation metrics (out of scope: needs note_seq).")
flags.DEFINE_bool("compute_final_only", False, "Metrics on the final samples only.")
flags.DEFINE_bool("flush", True, "Write pickles.")
flags.DEFINE_bool("animate", False, "Write sampling animations (out of scope).")
flags.DEFINE_bool("infill", False, "Infill the middle 16 latents of real examples (sample_ncsn.py:408-427).")
flags.DEFINE_bool("interpolate", False, "Interpolate between real examples in the diffusion latent space.")


def _restore(model_dir, shape, batch_size=1):
    """sample_ncsn.py:331-342: dummy-initialised (optimizer, ema, early_stop) filled from the newest checkpoint."""
    rng = random.PRNGKey(FLAGS.sample_seed)
    rng, model_rng = random.split(rng)
    model = train_ncsn.create_model(model_rng, shape, train_ncsn.model_kwargs(), batch_size=batch_size)
    optimizer = train_ncsn.create_optimizer(model, FLAGS.learning_rate)
    ema = train_utils.EMAHelper(mu=0, params=model.arena.clone())
    early_stop = train_utils.EarlyStopping()
    optimizer, ema, early_stop = checkpoints.restore_checkpoint(model_dir, (optimizer, ema, early_stop))
    return rng, optimizer


def generate_samples(sample_shape, num_samples, rng_seed=1):
    """sample_ncsn.py:313-365."""
    del rng_seed
    rng, optimizer = _restore(FLAGS.model_dir, sample_shape)
    sigmas = ebm_utils.create_noise_schedule(FLAGS.sigma_begin, FLAGS.sigma_end, FLAGS.num_sigmas, FLAGS.schedule_type)
    world, rank = parallel.world_size(), parallel.rank()
    rng, sample_rng = random.split(rng)          # sample_ncsn.py:350
    t0 = time.time()
    # every rank holds the same key; rank r generates rows [r*N/W, (r+1)*N/W) of the N-sample run (its slice of the
    # initial draw and of each step's noise), so the gathered result is the single-process result
    generated, collection, ld_metrics = train_ncsn.sample(optimizer.target, sigmas, sample_rng, sample_shape,
                                                          num_samples=num_samples, sampling=FLAGS.sampling,
                                                          epsilon=FLAGS.ld_epsilon, steps=FLAGS.ld_steps,
                                                          denoise=FLAGS.denoise, shard=(rank, world))
    torch.cuda.synchronize()
    logging.info("Generated samples in %f seconds", time.time() - t0)
    generated = parallel.gather_rows(generated)
    collection = parallel.gather_rows(collection.transpose(0, 1).contiguous()).transpose(0, 1)
    return generated.cpu().numpy(), collection.cpu().numpy(), ld_metrics


def infill_samples(samples, masks, rng_seed=1):
    """sample_ncsn.py:189-242: keep the masked (=1) entries of `samples`, regenerate the rest."""
    del rng_seed
    rng, optimizer = _restore(FLAGS.model_dir, samples.shape[1:])
    sigmas = ebm_utils.create_noise_schedule(FLAGS.sigma_begin, FLAGS.sigma_end, FLAGS.num_sigmas, FLAGS.schedule_type)
    init_rng, ld_rng = random.split(rng)
    init = random.uniform(init_rng, samples.shape)      # upstream initialises the infill chain with uniform noise
    generated, collection, ld_metrics = ebm_utils.diffusion_dynamics(ld_rng, optimizer.target, sigmas, init,
                                                                     FLAGS.ld_epsilon, FLAGS.ld_steps, FLAGS.denoise,
                                                                     True, samples, masks)
    return generated.cpu().numpy(), collection.cpu().numpy(), ebm_utils.collate_sampling_metrics(ld_metrics)


def diffusion_stochastic_encoder(samples, rng_seed=1):
    """sample_ncsn.py:245-265: z ~ q(x_T | x_0).  Upstream indexes ``alphas_prod[T]`` (one past the end), which jax
    clamps to the last element; the first output of ``split(rng)`` is the noise key."""
    assert FLAGS.sampling == "ddpm"
    rng = random.PRNGKey(rng_seed)
    betas = ebm_utils.create_noise_schedule(FLAGS.sigma_begin, FLAGS.sigma_end, FLAGS.num_sigmas, FLAGS.schedule_type)
    alphas_prod = np.cumprod((np.float32(1.0) - betas).astype(np.float32), dtype=np.float32)
    rng, _ = random.split(rng)
    x = torch.as_tensor(np.ascontiguousarray(samples, np.float32), device="cuda")
    noise = random.normal(rng, x.shape)
    last = alphas_prod[-1]                        # alphas_prod[T] clamped
    return float(np.sqrt(last)) * x + float(np.sqrt(np.float32(1.0) - last)) * noise


def diffusion_decoder(z_list, rng_seed=1):
    """sample_ncsn.py:268-310: one reverse chain per latent, all with the same sampling key."""
    assert FLAGS.sampling == "ddpm"
    rng = random.PRNGKey(rng_seed)
    rng, ld_rng, model_rng = random.split(rng, 3)
    del model_rng
    betas = ebm_utils.create_noise_schedule(FLAGS.sigma_begin, FLAGS.sigma_end, FLAGS.num_sigmas, FLAGS.schedule_type)
    _, optimizer = _restore(FLAGS.model_dir, tuple(z_list[0].shape[1:]))
    gen, collects, metrics = [], [], []
    for i, z in enumerate(z_list):
        generated, collection, ld_metrics = ebm_utils.diffusion_dynamics(ld_rng, optimizer.target, betas, z,
                                                                         FLAGS.ld_epsilon, FLAGS.ld_steps, FLAGS.denoise,
                                                                         False)
        gen.append(generated.cpu().numpy())
        collects.append(collection.cpu().numpy())
        metrics.append(ebm_utils.collate_sampling_metrics(ld_metrics))
        logging.info("Generated samples %i out of %i", i, len(z_list))
    return gen, collects, metrics


def _save(obj, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(obj, f, protocol=4)
    logging.info("Saved to %s", path)


def main(argv):
    del argv
    parallel.init_from_env()
    if FLAGS.compute_metrics or FLAGS.animate:
        raise ValueError("--compute_metrics / --animate depend on note_seq / matplotlib code that is out of scope")
    _, eval_ds = input_pipeline.get_dataset(
        dataset=FLAGS.dataset, data_shape=FLAGS.data_shape, problem=FLAGS.problem, batch_size=FLAGS.batch_size,
        normalize=FLAGS.normalize, pca_ckpt=FLAGS.pca_ckpt, slice_ckpt=FLAGS.slice_ckpt,
        dim_weights_ckpt=FLAGS.dim_weights_ckpt, include_cardinality=False, synthetic=FLAGS.synthetic,
        synthetic_examples=max(FLAGS.synthetic_examples, 8 * FLAGS.sample_size), seed=FLAGS.seed)
    real = []
    for batch in eval_ds:                                  # sample_ncsn.py:397-402
        real.append(batch)
        if sum(len(b) for b in real) >= FLAGS.sample_size:
            break
    real = np.concatenate(real)[:FLAGS.sample_size]
    shape = real.shape[1:]
    if FLAGS.infill:                                       # sample_ncsn.py:408-427 (sequence branch)
        if real.ndim != 3 or real.shape[1] != 32:
            raise ValueError("--infill needs (N, 32, C) latent sequences")
        samples = np.copy(real)
        samples[:, 8:-8, :] = 0                            # the middle 16 latents are regenerated
        masks = np.zeros(samples.shape, np.float32)
        masks[:, :8, :] = 1
        masks[:, -8:, :] = 1                               # first and last 8 are held fixed
        generated, collection, _ = infill_samples(samples, masks, FLAGS.sample_seed)
    elif FLAGS.interpolate:                                # sample_ncsn.py:429-438
        starts = real
        goals = np.roll(starts, shift=1, axis=0)
        starts_z = diffusion_stochastic_encoder(starts, FLAGS.sample_seed)
        goals_z = diffusion_stochastic_encoder(goals, FLAGS.sample_seed)
        interp_zs = [float(1 - alpha) * starts_z + float(alpha) * goals_z for alpha in np.linspace(0.0, 1.0, 9)]
        gen, coll, _ = diffusion_decoder(interp_zs, FLAGS.sample_seed)
        generated, collection = np.stack(gen), np.stack(coll)
    else:
        generated, collection, _ = generate_samples(shape, len(real), FLAGS.sample_seed)
    if FLAGS.flush and parallel.rank() == 0:
        slice_idx = input_pipeline.load(os.path.expanduser(FLAGS.slice_ckpt)) if FLAGS.slice_ckpt else None
        dim_w = input_pipeline.load(os.path.expanduser(FLAGS.dim_weights_ckpt)) if FLAGS.dim_weights_ckpt else None
        inv = lambda a: input_pipeline.inverse_data_transform(  # noqa: E731
            a, normalize=FLAGS.normalize, data_min=eval_ds.min, data_max=eval_ds.max, slice_idx=slice_idx,
            dim_weights=dim_w, out_channels=int(FLAGS.data_shape[-1]))
        out = os.path.join(FLAGS.model_dir, FLAGS.sampling_dir, "ncsn")
        _save(inv(collection), os.path.join(out, "collection.pkl"))
        _save(inv(real), os.path.join(out, "real.pkl"))
        _save(inv(generated), os.path.join(out, "generated.pkl"))
    parallel.shutdown()


if __name__ == "__main__":
    app.run(main)

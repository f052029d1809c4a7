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
flags.DEFINE_bool("infill", False, "Infill the second half of real examples.")
flags.DEFINE_bool("interpolate", False, "Interpolation mode (next scope row, SURVEY 8(f3)).")


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
    local = parallel.shard_size(num_samples)
    key = random.split(rng, world)[rank] if world > 1 else rng
    t0 = time.time()
    generated, collection, ld_metrics = train_ncsn.sample(optimizer.target, sigmas, key, sample_shape,
                                                          num_samples=local, sampling=FLAGS.sampling,
                                                          epsilon=FLAGS.ld_epsilon, steps=FLAGS.ld_steps,
                                                          denoise=FLAGS.denoise)
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
    init = random.normal(init_rng, samples.shape)
    generated, collection, ld_metrics = ebm_utils.diffusion_dynamics(ld_rng, optimizer.target, sigmas, init,
                                                                     FLAGS.ld_epsilon, FLAGS.ld_steps, FLAGS.denoise,
                                                                     True, samples, masks)
    return generated.cpu().numpy(), collection.cpu().numpy(), ebm_utils.collate_sampling_metrics(ld_metrics)


def _save(obj, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(obj, f, protocol=4)
    logging.info("Saved to %s", path)


def main(argv):
    del argv
    parallel.init_from_env()
    if FLAGS.interpolate:
        raise ValueError("--interpolate is the next scope row (SURVEY 8(f3)); not built yet")
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
    if FLAGS.infill:
        masks = np.zeros_like(real)
        masks[:, : real.shape[1] // 2] = 1.0
        generated, collection, _ = infill_samples(real, masks, FLAGS.sample_seed)
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

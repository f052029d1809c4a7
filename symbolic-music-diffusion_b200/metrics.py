"""Latent-space evaluation metrics with the reference's names and semantics (utils/metrics.py:24-77): Frechet
distance between Gaussian fits and kernel (MMD) distances between two sample sets.  Post-hoc evaluation on host arrays
(the reference runs them through scipy / scikit-learn on the CPU as well); the MIDI framewise statistics of
utils/metrics.py:80-244 need note_seq and are out of scope."""
from __future__ import annotations

import numpy as np
import scipy.linalg


def frechet_distance(real, fake, eps: float = 1e-6) -> float:
    """|mu1 - mu2|^2 + tr(S1) + tr(S2) - 2 tr(sqrtm(S1 S2))   (utils/metrics.py:24-55).  Lower is better.

    (Upstream's singular-product fallback references an undefined `eps`; 1e-6 -- the value of the FID code it was
    taken from -- is used here.)"""
    real, fake = np.asarray(real, np.float64), np.asarray(fake, np.float64)
    mu1, sigma1 = real.mean(axis=0), np.cov(real, rowvar=False)
    mu2, sigma2 = fake.mean(axis=0), np.cov(fake, rowvar=False)
    diff = mu1 - mu2
    covmean = scipy.linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = scipy.linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):       # numerical error may leave a slight imaginary component
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2.0 * np.trace(covmean))


def _sq_dists(x, y):
    xx = (x * x).sum(axis=1)[:, None]
    yy = (y * y).sum(axis=1)[None, :]
    return np.maximum(xx + yy - 2.0 * x.dot(y.T), 0.0)


def _mmd(real, fake, kernel) -> float:
    real, fake = np.asarray(real, np.float64), np.asarray(fake, np.float64)
    return float(kernel(real, real).mean() + kernel(fake, fake).mean() - 2.0 * kernel(real, fake).mean())


def mmd_rbf(real, fake, gamma: float = 1.0) -> float:
    """Biased MMD^2 estimate with k(x, y) = exp(-gamma |x - y|^2)   (utils/metrics.py:58-66)."""
    return _mmd(real, fake, lambda a, b: np.exp(-gamma * _sq_dists(a, b)))


def mmd_polynomial(real, fake, degree: int = 2, gamma: float = 1, coef0: float = 0) -> float:
    """Biased MMD^2 estimate with k(x, y) = (gamma <x, y> + coef0)^degree   (utils/metrics.py:69-77)."""
    return _mmd(real, fake, lambda a, b: (gamma * a.dot(b.T) + coef0) ** degree)

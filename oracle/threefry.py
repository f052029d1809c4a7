"""NumPy restatement of jax.random 0.2.8 (threefry2x32) -- TEST INFRASTRUCTURE ONLY.

The algorithm lives in an un-vendored third-party dependency of the reference:
``jax==0.2.8`` / ``jaxlib==0.1.57`` (requirements.txt:94-95).  Call sites on the hot
path: utils/losses.py:271-294, utils/ebm_utils.py:329,342,360, train_ncsn.py:318-319,
358,536-540.  Pinned by the published Random123 / JAX known-answer values in
tests/test_oracle_threefry.py.
"""
from __future__ import annotations

import numpy as np

_R = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x, r):
    return ((x << _U32(r)) | (x >> _U32(32 - r))).astype(_U32)


def threefry2x32_block(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  All args uint32 arrays (broadcastable)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, _U32)
        k1 = np.asarray(k1, _U32)
        x0 = np.asarray(x0, _U32).copy()
        x1 = np.asarray(x1, _U32).copy()
        ks = (k0, k1, (k0 ^ k1 ^ _U32(0x1BD11BDA)).astype(_U32))
        x0 = (x0 + ks[0]).astype(_U32)
        x1 = (x1 + ks[1]).astype(_U32)
        for i in range(5):
            for r in _R[i % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(_U32)
            x0 = (x0 + ks[(i + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(i + 2) % 3] + _U32(i + 1)).astype(_U32)
    return x0, x1


def prng_key(seed: int) -> np.ndarray:
    """jax.random.PRNGKey: [seed >> 32, seed & 0xFFFFFFFF]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def threefry_2x32(key, count: np.ndarray) -> np.ndarray:
    """jax.random.threefry_2x32: split the flat counter array in halves, run, concat."""
    count = np.asarray(count, _U32).ravel()
    n = count.size
    odd = n % 2
    if odd:
        count = np.concatenate([count, np.zeros((1,), _U32)])
    half = count.size // 2
    o0, o1 = threefry2x32_block(key[0], key[1], count[:half], count[half:])
    out = np.concatenate([o0, o1])
    return out[:-1] if odd else out


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split: threefry_2x32(key, iota(2*num)).reshape(num, 2)."""
    return threefry_2x32(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits(key, shape) -> np.ndarray:
    n = int(np.prod(shape)) if len(shape) else 1
    return threefry_2x32(key, np.arange(n, dtype=_U32)).reshape(shape)


def uniform01(key, shape) -> np.ndarray:
    """float32 in [0,1): bitcast((bits >> 9) | 0x3F800000) - 1."""
    bits = random_bits(key, shape)
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return f.astype(np.float32)


def uniform(key, shape, minval=0.0, maxval=1.0) -> np.ndarray:
    """jax 0.2.8 random.uniform: max(minval, u * (maxval - minval) + minval)."""
    u = uniform01(key, shape)
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    return np.maximum(minval, u * (maxval - minval) + minval).astype(np.float32)


def erfinv_f32(x: np.ndarray) -> np.ndarray:
    """XLA's float32 ErfInv (Giles' polynomial), evaluated in float32."""
    f32 = np.float32
    x = np.asarray(x, f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log(((f32(1) - x) * (f32(1) + x)).astype(f32)).astype(f32)
        lt = w < f32(5)
        w1 = (w - f32(2.5)).astype(f32)
        w2 = (np.sqrt(np.maximum(w, f32(0))) - f32(3)).astype(f32)
        c1 = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
              -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
        c2 = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
              -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]
        p1 = np.full_like(w1, f32(c1[0]))
        for c in c1[1:]:
            p1 = (f32(c) + p1 * w1).astype(f32)
        p2 = np.full_like(w2, f32(c2[0]))
        for c in c2[1:]:
            p2 = (f32(c) + p2 * w2).astype(f32)
        p = np.where(lt, p1, p2).astype(f32)
        r = (p * x).astype(f32)
    return np.where(np.abs(x) == f32(1), np.sign(x) * f32(np.inf), r).astype(f32)


def normal(key, shape) -> np.ndarray:
    """jax 0.2.8 random.normal float32: sqrt(2) * erf_inv(uniform(nextafter(-1,0), 1))."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(key, shape, lo, np.float32(1.0))
    return (np.float32(np.sqrt(2)) * erfinv_f32(u)).astype(np.float32)


def randint(key, shape, minval: int, maxval: int) -> np.ndarray:
    """jax 0.2.8 random.randint (int32)."""
    k1, k2 = split(key, 2)
    hi = random_bits(k1, shape).astype(np.uint64)
    lo = random_bits(k2, shape).astype(np.uint64)
    span = np.uint64((maxval - minval) & 0xFFFFFFFF)
    mult = np.uint64(2 ** 16) % span
    mult = (mult * mult) % span
    off = ((hi % span) * mult + (lo % span)) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)

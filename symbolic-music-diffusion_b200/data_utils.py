"""Host-side data helpers with the names of the reference's ``utils/data_utils.py`` (TF- and JAX-free, NumPy only).

The dataset statistics work on any re-iterable of NumPy batches (``smd_b200.input_pipeline.Dataset`` or a plain list);
caching follows upstream's file naming (``cache/{split}_{config}_{min,max,mean,stddev,cardinality}.pkl``) but writes
atomically, so concurrent data-parallel ranks never read a half-written pickle.
"""
from __future__ import annotations

import os
from typing import Iterable, Iterator, Optional, Sequence, Tuple

import numpy as np

from .input_pipeline import _save_atomic, load, save  # noqa: F401  (save / load: utils/data_utils.py:30-41)


def _cache_path(cache_dir: Optional[str], ds_split: str, config: str, what: str) -> Optional[str]:
    return None if cache_dir is None else os.path.join(cache_dir, f"cache/{ds_split}_{config}_{what}.pkl")


def compute_dataset_cardinality(ds: Iterable, ds_split: str = "train", cache: bool = False,
                                cache_dir: Optional[str] = None, config: str = "") -> int:
    """Number of elements of the (batched or unbatched) dataset (utils/data_utils.py:63-90)."""
    path = _cache_path(cache_dir, ds_split, config, "cardinality")
    if path and os.path.exists(path):
        n = int(load(path))
    else:
        n = sum(1 for _ in ds)
    if cache:
        assert path is not None, "cache=True needs cache_dir"
        _save_atomic(n, path)
    return n


def compute_dataset_statistics(ds: Iterable, ds_split: str = "train", cache: bool = False,
                               cache_dir: Optional[str] = None, config: str = ""):
    """Element-wise mean and standard deviation over the dataset's elements, sqrt(E[x^2] - E[x]^2) like upstream
    (utils/data_utils.py:93-125), accumulated in float64 in one pass."""
    pm, ps = _cache_path(cache_dir, ds_split, config, "mean"), _cache_path(cache_dir, ds_split, config, "stddev")
    if pm and ps and os.path.exists(pm) and os.path.exists(ps):
        mean, std = load(pm), load(ps)
    else:
        s1 = s2 = None
        n = 0
        for x in ds:
            x = np.asarray(x, np.float64)
            s1 = x.copy() if s1 is None else s1 + x
            s2 = x * x if s2 is None else s2 + x * x
            n += 1
        if n == 0:
            raise ValueError("empty dataset")
        mean = s1 / n
        std = np.sqrt(np.maximum(s2 / n - mean * mean, 0.0))
        mean, std = mean.astype(np.float32), std.astype(np.float32)
    if cache:
        assert pm is not None, "cache=True needs cache_dir"
        _save_atomic(mean, pm)
        _save_atomic(std, ps)
    return mean, std


def compute_dataset_min_max(ds: Iterable, ds_split: str = "train", cache: bool = False,
                            cache_dir: Optional[str] = None, config: str = "") -> Tuple[float, float]:
    """Global scalar minimum and maximum over every element (utils/data_utils.py:128-156), one pass."""
    pmin, pmax = _cache_path(cache_dir, ds_split, config, "min"), _cache_path(cache_dir, ds_split, config, "max")
    if pmin and pmax and os.path.exists(pmin) and os.path.exists(pmax):
        lo, hi = load(pmin), load(pmax)
    else:
        lo, hi = np.float32(np.finfo(np.float32).max), np.float32(np.finfo(np.float32).min)
        for x in ds:
            x = np.asarray(x, np.float32)
            lo, hi = min(lo, x.min()), max(hi, x.max())
    if cache:
        assert pmin is not None, "cache=True needs cache_dir"
        _save_atomic(lo, pmin)
        _save_atomic(hi, pmax)
    return lo, hi


def _truncate_embeddings(embeddings: np.ndarray, length: int) -> np.ndarray:
    """First `length` rows, zero-padded when there are fewer (utils/data_utils.py:194-205)."""
    embeddings = np.asarray(embeddings)
    out = np.zeros((length, embeddings.shape[-1]), dtype=np.result_type(embeddings.dtype, np.float64))
    k = min(length, len(embeddings))
    out[:k] = embeddings[:k]
    return out


def self_similarity(embeddings: np.ndarray, normalized: bool = True, max_len: int = 80) -> np.ndarray:
    """(max_len, max_len) Gram matrix of the truncated / padded embedding sequence; cosine similarities when
    `normalized`, with zero rows giving 0 instead of NaN (utils/data_utils.py:208-218)."""
    e = _truncate_embeddings(embeddings, max_len)
    if normalized:
        nrm = np.linalg.norm(e, axis=1, keepdims=True)
        e = np.divide(e, nrm, out=np.zeros_like(e), where=nrm > 0)
    return e @ e.T


def unroll_upper_triangular(matrix: np.ndarray) -> list:
    """Strict upper triangle of a square matrix, row by row, as a list (utils/data_utils.py:221-231)."""
    matrix = np.asarray(matrix)
    if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1]:
        raise AssertionError("Not a square matrix.")
    return list(matrix[np.triu_indices(matrix.shape[0], 1)])


def roll_upper_triangular(vector: Sequence, size: int) -> np.ndarray:
    """Inverse of unroll_upper_triangular: symmetric (size, size) matrix with a unit diagonal
    (utils/data_utils.py:234-245)."""
    vector = np.asarray(vector, np.float64)
    if len(vector) != size * (size - 1) // 2:
        raise AssertionError("vector length does not match size")
    out = np.ones((size, size))
    iu = np.triu_indices(size, 1)
    out[iu] = vector
    out[(iu[1], iu[0])] = vector
    return out


def erase_bars(embeddings: np.ndarray, indices) -> np.ndarray:
    """Copy of `embeddings` with the rows `indices` zeroed (utils/data_utils.py:248-258; functional like
    jax.ops.index_update: the input is not modified)."""
    out = np.array(embeddings, copy=True)
    out[np.asarray(indices, dtype=np.intp)] = 0
    return out


def infill_bars(embeddings: np.ndarray, chunk_params, erased_chunk_indices) -> np.ndarray:
    """Copy of `embeddings` with rows `erased_chunk_indices` replaced by `chunk_params`
    (utils/data_utils.py:261-275)."""
    assert len(chunk_params) == len(erased_chunk_indices)
    out = np.array(embeddings, copy=True)
    out[np.asarray(erased_chunk_indices, dtype=np.intp)] = np.asarray(chunk_params, dtype=out.dtype)
    return out


def batches(data: np.ndarray, labels: Optional[np.ndarray] = None, batch_size: int = 32) -> Iterator:
    """Consecutive full batches (the remainder is dropped), with their labels if given
    (utils/data_utils.py:278-298)."""
    if labels is not None:
        assert len(data) == len(labels)
    for j in range(0, (len(data) // batch_size) * batch_size, batch_size):
        yield (data[j:j + batch_size], labels[j:j + batch_size]) if labels is not None else data[j:j + batch_size]


def shuffle(data: np.ndarray, labels: Optional[np.ndarray] = None):
    """One permutation (NumPy's global RNG, like upstream) applied to data and labels (utils/data_utils.py:301-320)."""
    idx = np.random.permutation(len(data))
    if labels is None:
        return data[idx]
    assert len(data) == len(labels)
    return data[idx], labels[idx]

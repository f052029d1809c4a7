"""smd_b200/data_utils.py: the helper surface of the reference's utils/data_utils.py, checked against hand-computed
values (CPU only)."""
import os

import numpy as np
import pytest

from smd_b200 import data_utils as D


def test_statistics_min_max_cardinality_and_cache(tmp_path):
    ds = [np.array([[1.0, -2.0], [3.0, 4.0]], np.float32), np.array([[5.0, 0.0], [-7.0, 8.0]], np.float32)]
    assert D.compute_dataset_cardinality(ds) == 2
    lo, hi = D.compute_dataset_min_max(ds)
    assert (float(lo), float(hi)) == (-7.0, 8.0)
    mean, std = D.compute_dataset_statistics(ds)
    np.testing.assert_allclose(mean, [[3.0, -1.0], [-2.0, 6.0]])
    np.testing.assert_allclose(std, [[2.0, 1.0], [5.0, 2.0]])               # sqrt(E[x^2] - E[x]^2) over the 2 elements
    # cache files use upstream's names and take precedence on the next call
    D.compute_dataset_min_max(ds, "train", True, str(tmp_path), "cfg")
    assert os.path.exists(tmp_path / "cache" / "train_cfg_min.pkl") and os.path.exists(tmp_path / "cache" / "train_cfg_max.pkl")
    lo2, hi2 = D.compute_dataset_min_max([np.zeros((2, 2), np.float32)], "train", False, str(tmp_path), "cfg")
    assert (float(lo2), float(hi2)) == (-7.0, 8.0)
    D.compute_dataset_cardinality(ds, "eval", True, str(tmp_path), "cfg")
    assert D.compute_dataset_cardinality([], "eval", False, str(tmp_path), "cfg") == 2


def test_truncate_and_self_similarity():
    e = np.array([[3.0, 4.0], [0.0, 0.0], [1.0, 0.0]])
    t = D._truncate_embeddings(e, 5)
    assert t.shape == (5, 2) and np.all(t[3:] == 0) and np.all(t[:3] == e)
    assert np.all(D._truncate_embeddings(e, 2) == e[:2])
    s = D.self_similarity(e, normalized=True, max_len=4)
    expect = np.zeros((4, 4))
    expect[0, 0] = expect[2, 2] = 1.0
    expect[0, 2] = expect[2, 0] = 0.6                                        # cos((3,4), (1,0))
    np.testing.assert_allclose(s, expect, atol=1e-12)                         # zero rows give 0, not NaN
    raw = D.self_similarity(e, normalized=False, max_len=3)
    np.testing.assert_allclose(raw, e @ e.T)


def test_upper_triangular_round_trip():
    m = np.array([[1.0, 2.0, 3.0, 4.0], [2.0, 1.0, 5.0, 6.0], [3.0, 5.0, 1.0, 7.0], [4.0, 6.0, 7.0, 1.0]])
    v = D.unroll_upper_triangular(m)
    assert v == [2.0, 3.0, 4.0, 5.0, 6.0, 7.0]
    np.testing.assert_array_equal(D.roll_upper_triangular(v, 4), m)
    with pytest.raises(AssertionError):
        D.unroll_upper_triangular(np.zeros((2, 3)))
    with pytest.raises(AssertionError):
        D.roll_upper_triangular([1.0, 2.0], 4)


def test_erase_and_infill_are_functional():
    e = np.arange(12, dtype=np.float32).reshape(4, 3)
    erased = D.erase_bars(e, [1, 3])
    assert np.all(erased[[1, 3]] == 0) and np.all(erased[[0, 2]] == e[[0, 2]]) and e[1, 0] == 3.0   # input untouched
    filled = D.infill_bars(erased, np.array([[9, 9, 9], [8, 8, 8]], np.float32), [1, 3])
    assert np.all(filled[1] == 9) and np.all(filled[3] == 8) and np.all(erased[1] == 0)
    with pytest.raises(AssertionError):
        D.infill_bars(erased, np.zeros((1, 3)), [1, 3])


def test_batches_and_shuffle():
    x = np.arange(10)[:, None] * np.ones((1, 2))
    y = np.arange(10)
    got = list(D.batches(x, y, batch_size=4))
    assert len(got) == 2 and np.all(got[1][1] == [4, 5, 6, 7]) and got[0][0].shape == (4, 2)      # remainder dropped
    assert [b.shape for b in D.batches(x, batch_size=3)] == [(3, 2)] * 3
    np.random.seed(0)
    xs, ys = D.shuffle(x, y)
    assert sorted(ys.tolist()) == list(range(10)) and np.all(xs[:, 0] == ys)                      # same permutation
    assert sorted(D.shuffle(y).tolist()) == list(range(10))

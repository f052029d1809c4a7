"""TF-free input pipeline (the data format either side of the hot path): TFRecord framing, tf.train.Example
parsing, slice / normalise transforms, caches -- behaviour of input_pipeline.py:36-48,113-235 upstream."""
import os
import pickle
import struct

import numpy as np
import pytest

from smd_b200 import input_pipeline as ip


def test_crc32c_known_answers():
    assert ip.crc32c(b"123456789") == 0xE3069283          # CRC-32C (Castagnoli) check value
    assert ip.crc32c(b"") == 0
    # TFRecord mask: rotate right by 15, add 0xa282ead8
    c = ip.crc32c(b"abc")
    assert ip.masked_crc(b"abc") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_roundtrip_and_wire_format(tmp_path):
    x = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4) * 0.5 - 3
    payload = ip.serialize_example(x)
    ex = ip.parse_example(payload)
    np.testing.assert_array_equal(ex["inputs"], x.reshape(-1))
    np.testing.assert_array_equal(ex["input_shape"], [2, 3, 4])
    # hand-built message with NON-packed floats / ints (both encodings are legal protobuf)
    def ld(n, p):
        return bytes([(n << 3) | 2, len(p)]) + p
    floats = b"".join(bytes([(1 << 3) | 5]) + struct.pack("<f", v) for v in (1.5, -2.0))
    ints = bytes([(1 << 3) | 0, 7, (1 << 3) | 0, 9])
    feat_a = ld(1, ld(1, b"inputs") + ld(2, ld(2, floats)))
    feat_b = ld(1, ld(1, b"input_shape") + ld(2, ld(3, ints)))
    ex2 = ip.parse_example(ld(1, feat_a + feat_b))
    np.testing.assert_array_equal(ex2["inputs"], [1.5, -2.0])
    np.testing.assert_array_equal(ex2["input_shape"], [7, 9])
    path = str(tmp_path / "a.tfrecord")
    ip.write_tfrecord(path, [payload, payload])
    recs = list(ip.read_tfrecord(path, verify=True))
    assert recs == [payload, payload]
    raw = bytearray(open(path, "rb").read())
    raw[20] ^= 0xFF
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        list(ip.read_tfrecord(path, verify=True))


def _make_dataset(root, n_train=70, n_eval=40, shape=(32, 512), seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    data = {}
    for split, n in (("train", n_train), ("eval", n_eval)):
        arr = rng.standard_normal((n, *shape)).astype(np.float32)
        data[split] = arr
        half = n // 2
        ip.write_tfrecord(os.path.join(root, f"{split}-0000.tfrecord"), [ip.serialize_example(a) for a in arr[:half]])
        ip.write_tfrecord(os.path.join(root, f"{split}-0001.tfrecord"), [ip.serialize_example(a) for a in arr[half:]])
    return data


def test_get_dataset_slice_normalise_cache(tmp_path):
    root = str(tmp_path / "ds")
    data = _make_dataset(root)
    idx = np.sort(np.random.default_rng(1).choice(512, 42, replace=False)).astype(np.int64)
    slice_path = str(tmp_path / "slice-mel-512.pkl")
    pickle.dump(idx, open(slice_path, "wb"))
    train, ev = ip.get_dataset(dataset=root, data_shape=["32", "512"], problem="vae", batch_size=16,
                               slice_ckpt=slice_path)
    assert train.examples == 70 // 16 and ev.examples == 40 // 16       # batches per epoch, drop_remainder
    # every complete batch is used for min/max: recompute from the batches the stream actually yields
    batches = list(train)
    assert len(batches) == 4 and all(b.shape == (16, 32, 42) and b.dtype == np.float32 for b in batches)
    allb = np.concatenate(batches)
    # min / max come from ONE pass over complete batches (utils/data_utils.py:93-126): an example that fell into that
    # pass's dropped remainder may land marginally outside [-1, 1] in a later epoch, exactly as upstream
    assert allb.min() >= -1.05 and allb.max() <= 1.05
    # one global scalar min/max per split (input_pipeline.py:185-207)
    sl = data["train"][..., idx]
    assert train.min >= sl.min() - 1e-6 and train.max <= sl.max() + 1e-6
    for name in ("train_slice-mel-512_min.pkl", "train_slice-mel-512_max.pkl", "eval_slice-mel-512_min.pkl",
                 "train_16_cardinality.pkl"):
        assert os.path.exists(os.path.join(root, "cache", name)), name
    # examples come back un-permuted in content: each normalised row maps to exactly one source example
    inv = ip.inverse_data_transform(batches[0][:1], data_min=train.min, data_max=train.max, slice_idx=idx)
    src = data["train"][..., idx]
    d = np.abs(src - inv[0][..., idx][None]).reshape(len(src), -1).max(1)
    assert d.min() < 1e-4
    assert inv.shape == (1, 32, 512) and inv.dtype == np.float64     # randn fill, as upstream
    # second call hits the caches
    train2, _ = ip.get_dataset(dataset=root, data_shape=["32", "512"], problem="vae", batch_size=16,
                               slice_ckpt=slice_path)
    assert train2.min == train.min and train2.examples == train.examples


def test_get_dataset_errors_and_synthetic(tmp_path):
    with pytest.raises(ValueError):
        ip.get_dataset(problem="tokens2")
    with pytest.raises(FileNotFoundError):
        list(ip.get_dataset(dataset=str(tmp_path / "missing"), data_shape=[4], batch_size=2)[0])
    tr, ev = ip.get_dataset(data_shape=[32, 512], batch_size=8, synthetic=True, synthetic_examples=64)
    b = next(iter(tr))
    assert b.shape == (8, 32, 512) and tr.examples == 8 and abs(b).max() <= 1.0 + 1e-6


def test_reads_records_serialized_by_the_protobuf_runtime():
    """tests/golden/tfrecord_protobuf.tfrecord was written by Google's protobuf runtime from the tf.train.Example
    schema (scripts/make_tfrecord_fixture.py) with an independent bit-wise crc32c: the hand-written TFRecord /
    protobuf reader must decode it exactly (layout of scripts/transform_encoded_data.py:71-92 upstream)."""
    here = os.path.join(os.path.dirname(__file__), "golden")
    want = np.load(os.path.join(here, "tfrecord_protobuf_expected.npz"))
    recs = list(ip.read_tfrecord(os.path.join(here, "tfrecord_protobuf.tfrecord"), verify=True))
    assert len(recs) == 3
    for i, payload in enumerate(recs):
        ex = ip.parse_example(payload)
        a = want[f"ex{i}"]
        assert ex["inputs"].dtype == np.float32 and ex["input_shape"].dtype == np.int64
        assert tuple(ex["input_shape"]) == a.shape
        np.testing.assert_array_equal(ex["inputs"].reshape(a.shape).view(np.uint32), a.view(np.uint32))   # bit-exact
    ex1 = ip.parse_example(recs[1])
    np.testing.assert_array_equal(ex1["targets"].reshape(tuple(ex1["target_shape"])), want["ex1_targets"])

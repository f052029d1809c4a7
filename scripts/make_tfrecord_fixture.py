"""Generates tests/golden/tfrecord_protobuf.tfrecord (+ .npz with the arrays it must decode to).

The records are serialized by GOOGLE'S protobuf runtime from a schema declared here with the field numbers of
tensorflow/core/example/{feature,example}.proto (TensorFlow itself is not installable in this image), in the shape
the reference's writer produces (scripts/transform_encoded_data.py:71-92: `inputs` FloatList + `input_shape` Int64List,
plus `targets` / `target_shape` in 'sequences' mode) -- i.e. by an encoder that is independent of the hand-written
parser in smd_b200/input_pipeline.py.  Framing per the TFRecord format: u64 length, masked crc32c(length), payload,
masked crc32c(payload).

  python scripts/make_tfrecord_fixture.py
"""
import os
import struct
import sys

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def example_class():
    f = descriptor_pb2.FileDescriptorProto(name="smd_fixture_example.proto", package="tensorflow", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = f.message_type.add()
        m.name = name
        return m

    def field(m, name, number, typ, label=T.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
        fd = m.field.add()
        fd.name, fd.number, fd.type, fd.label = name, number, typ, label
        if type_name:
            fd.type_name = type_name
        if packed is not None:
            fd.options.packed = packed
        if oneof is not None:
            fd.oneof_index = oneof
        return fd

    field(msg("BytesList"), "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    field(msg("FloatList"), "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, packed=True)
    field(msg("Int64List"), "value", 1, T.TYPE_INT64, T.LABEL_REPEATED, packed=True)
    feat = msg("Feature")
    feat.oneof_decl.add().name = "kind"
    field(feat, "bytes_list", 1, T.TYPE_MESSAGE, type_name=".tensorflow.BytesList", oneof=0)
    field(feat, "float_list", 2, T.TYPE_MESSAGE, type_name=".tensorflow.FloatList", oneof=0)
    field(feat, "int64_list", 3, T.TYPE_MESSAGE, type_name=".tensorflow.Int64List", oneof=0)
    feats = msg("Features")
    entry = feats.nested_type.add()
    entry.name = "FeatureEntry"
    entry.options.map_entry = True
    field(entry, "key", 1, T.TYPE_STRING)
    field(entry, "value", 2, T.TYPE_MESSAGE, type_name=".tensorflow.Feature")
    field(feats, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name=".tensorflow.Features.FeatureEntry")
    field(msg("Example"), "features", 1, T.TYPE_MESSAGE, type_name=".tensorflow.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow.Example"))


def crc32c_bitwise(data: bytes) -> int:
    """Castagnoli CRC, bit by bit (deliberately not the table-driven routine of the product)."""
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 & -(c & 1))
    return c ^ 0xFFFFFFFF


def masked(data: bytes) -> int:
    c = crc32c_bitwise(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def main():
    assert crc32c_bitwise(b"123456789") == 0xE3069283      # the standard CRC-32C check value
    Example = example_class()
    rng = np.random.default_rng(20260924)
    arrays = {"ex0": rng.standard_normal((32, 512)).astype(np.float32),      # one MusicVAE latent sequence
              "ex1": rng.standard_normal((4, 8)).astype(np.float32),
              "ex2": np.array([[np.float32(-0.0), np.float32(1e-38), np.float32(3.4e38), np.float32(np.pi)]], np.float32)}
    targets = {"ex1": rng.standard_normal((2, 8)).astype(np.float32)}          # 'sequences' mode adds a target tensor
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "tfrecord_protobuf.tfrecord"), "wb") as f:
        for name, a in arrays.items():
            ex = Example()
            ex.features.feature["inputs"].float_list.value.extend(a.reshape(-1).tolist())
            ex.features.feature["input_shape"].int64_list.value.extend(list(a.shape))
            if name in targets:
                t = targets[name]
                ex.features.feature["targets"].float_list.value.extend(t.reshape(-1).tolist())
                ex.features.feature["target_shape"].int64_list.value.extend(list(t.shape))
            payload = ex.SerializeToString(deterministic=True)
            head = struct.pack("<Q", len(payload))
            f.write(head + struct.pack("<I", masked(head)) + payload + struct.pack("<I", masked(payload)))
    np.savez(os.path.join(out, "tfrecord_protobuf_expected.npz"), **arrays, ex1_targets=targets["ex1"])
    print("wrote", os.path.join(out, "tfrecord_protobuf.tfrecord"))


if __name__ == "__main__":
    main()

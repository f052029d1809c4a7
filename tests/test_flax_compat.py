"""flax-0.3.0 checkpoint wire format (smd_b200/flax_compat.py).  flax itself is not installable here, so
these tests pin the encoder's byte layout, the parameter-tree naming rule and the round trip -- not agreement with flax."""
import msgpack
import numpy as np
import pytest
import torch

from smd_b200 import checkpoints, flax_compat, ncsn, optim, train_utils
from smd_b200 import jrandom  # noqa: F401  (host key helpers only; no GPU call below)


def torch_equal(a, b):
    return bool(torch.equal(a.cpu(), b.cpu()))


def _target(arch="TransformerDDPM", shape=(32, 42), **kw):
    module = getattr(ncsn, arch).partial(**kw)
    from smd_b200 import nn
    _, params = module.init_by_shape(np.array([0, 3], np.uint32), [((2,) + shape, np.float32), ((2, 1, 1), np.float32)])
    model = nn.Model(module, params)
    opt = optim.Adam(learning_rate=1e-3).create(model)
    g = torch.Generator().manual_seed(0)
    opt.grad_ema.copy_(torch.randn(opt.grad_ema.shape, generator=g))
    opt.grad_sq_ema.copy_(torch.rand(opt.grad_sq_ema.shape, generator=g))
    opt.step = 17
    ema = train_utils.EMAHelper(mu=0.999, params=model.arena.clone())
    ema.params.flat.mul_(0.5)
    es = train_utils.EarlyStopping(patience=3, best_metric=0.25, patience_count=1)
    return opt, ema, es


def test_ndarray_ext_encoding_bytes():
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    blob = flax_compat.msgpack_serialize({"w": a, "s": np.int32(7)})
    raw = msgpack.unpackb(blob, raw=False)
    assert isinstance(raw["w"], msgpack.ExtType) and raw["w"].code == 1 and raw["s"].code == 3
    shape, dtype, buf = msgpack.unpackb(raw["w"].data, raw=True)
    assert shape == [2, 3] and dtype == b"float32" and buf == a.tobytes("C")
    back = flax_compat.msgpack_restore(blob)
    assert np.array_equal(back["w"], a) and back["s"] == 7 and back["s"].dtype == np.int32


def test_pre_linen_tree_naming_transformer():
    opt, ema, es = _target(num_layers=2, num_heads=8, num_mlp_layers=2, mlp_dims=2048)
    st = flax_compat.to_flax_state((opt, ema, es))
    p = st["0"]["target"]["params"]
    # counter per parent over ALL submodules: 0 is the parameter-less positional encoding, layers take 5 each
    expect = {"Dense_1", "LayerNorm_2", "MultiHeadDotProductAttention_3", "LayerNorm_4", "Dense_5", "Dense_6",
              "LayerNorm_7", "MultiHeadDotProductAttention_8", "LayerNorm_9", "Dense_10", "Dense_11",
              "LayerNorm_12", "Dense_13", "DenseFiLM_14", "DenseResBlock_15", "DenseFiLM_16", "DenseResBlock_17",
              "LayerNorm_18", "Dense_19"}
    assert set(p) == expect
    assert p["MultiHeadDotProductAttention_3"]["query"]["kernel"].shape == (128, 8, 16) and p["MultiHeadDotProductAttention_3"]["key"]["bias"].shape == (8, 16)
    assert p["MultiHeadDotProductAttention_3"]["out"]["kernel"].shape == (8, 16, 128)
    assert set(p["DenseFiLM_14"]) == {"Dense_1", "Dense_2", "Dense_3", "Dense_4"} and p["DenseFiLM_14"]["Dense_4"]["kernel"].shape == (512, 2048)
    assert set(p["DenseResBlock_15"]) == {"LayerNorm_0", "Dense_2", "LayerNorm_3", "Dense_5"}
    assert p["Dense_1"]["kernel"].shape == (42, 128) and p["Dense_19"]["kernel"].shape == (2048, 42)
    ps = st["0"]["state"]["param_states"]["Dense_13"]["kernel"]
    assert set(ps) == {"grad_ema", "grad_sq_ema"} and ps["grad_ema"].shape == (128, 2048)
    assert st["0"]["state"]["step"].dtype == np.int32 and int(st["0"]["state"]["step"]) == 17
    assert set(st["1"]) == {"mu", "params"} and set(st["2"]) == {"min_delta", "patience", "best_metric", "patience_count", "should_stop"}


def test_tree_naming_dense_ddpm():
    opt, ema, es = _target("DenseDDPM", shape=(512,), num_layers=3, mlp_dims=2048)
    p = flax_compat.to_flax_state((opt, ema, es))["0"]["target"]["params"]
    assert set(p) == {"Dense_0", "DenseFiLM_1", "DenseResBlock_2", "DenseFiLM_3", "DenseResBlock_4", "DenseFiLM_5",
                      "DenseResBlock_6", "LayerNorm_7", "Dense_8"}


@pytest.mark.parametrize("fmt", ["flax", "native"])
def test_checkpoint_round_trip_both_formats(tmp_path, fmt):
    opt, ema, es = _target(num_layers=1, num_heads=8, num_mlp_layers=1, mlp_dims=2048)
    checkpoints.save_checkpoint(str(tmp_path), (opt, ema, es), 5, keep=2, fmt=fmt)
    opt2, ema2, es2 = _target(num_layers=1, num_heads=8, num_mlp_layers=1, mlp_dims=2048)
    opt2.target.arena.flat.zero_(); opt2.grad_ema.zero_(); opt2.grad_sq_ema.zero_(); ema2.params.flat.zero_()
    opt2.step = 0
    opt3, ema3, es3 = checkpoints.restore_checkpoint(str(tmp_path), (opt2, ema2, train_utils.EarlyStopping()))
    def same(a, b):      # tensor by tensor: the alignment padding between arena tensors is not part of a flax tree
        return all(torch.equal(a[o:o + int(np.prod(s))], b[o:o + int(np.prod(s))]) for _, o, s in opt.target.arena.layout)
    assert same(opt3.target.arena.flat, opt.target.arena.flat)
    assert same(opt3.grad_ema, opt.grad_ema) and same(opt3.grad_sq_ema, opt.grad_sq_ema)
    assert same(ema3.params.flat, ema.params.flat) and ema3.mu == pytest.approx(0.999)
    assert opt3.step == 17 and es3.patience == 3 and es3.best_metric == pytest.approx(0.25) and es3.patience_count == 1


def test_restore_rejects_a_checkpoint_of_another_architecture(tmp_path):
    checkpoints.save_checkpoint(str(tmp_path), _target(num_layers=1, num_heads=8, num_mlp_layers=1), 1, fmt="flax")
    with pytest.raises((KeyError, ValueError)):
        checkpoints.restore_checkpoint(str(tmp_path), _target(num_layers=2, num_heads=8, num_mlp_layers=1))


def test_restore_accepts_the_round1_attention_block_spelling():
    """Checkpoints written before the naming fix called the attention block SelfAttention_<i>."""
    opt, ema, es = _target(num_layers=1, num_heads=8, num_mlp_layers=1, mlp_dims=2048)
    st = flax_compat.to_flax_state((opt, ema, es))

    def rename(tree):
        if not isinstance(tree, dict):
            return tree
        return {k.replace("MultiHeadDotProductAttention_", "SelfAttention_"): rename(v) for k, v in tree.items()}

    old = rename(st)
    assert "SelfAttention_3" in old["0"]["target"]["params"]
    opt2, ema2, es2 = _target(num_layers=1, num_heads=8, num_mlp_layers=1, mlp_dims=2048)
    before = opt.target.arena.flat.clone()
    opt2, ema2, es2 = flax_compat.load_flax_state(old, (opt2, ema2, es2))
    assert torch_equal(opt2.target.arena.flat, before)

"""Pins the jax.random 0.2.8 restatement (oracle/threefry.py) against published known answers
(Random123 threefry2x32 KATs and the values JAX documents for PRNGKey(0) / PRNGKey(42))."""
import numpy as np

from oracle import threefry as tf


def test_block_function_kats():
    assert [int(v) for v in np.ravel(tf.threefry2x32_block(0, 0, 0, 0))] == [0x6B200159, 0x99BA4EFE]
    assert [int(v) for v in np.ravel(tf.threefry2x32_block(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF))] == \
        [0x1CB996FC, 0xBB002BE7]
    assert [int(v) for v in np.ravel(tf.threefry2x32_block(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3))] == \
        [0xC4923A9C, 0x483DF7A0]


def test_split_prngkey0():
    np.testing.assert_array_equal(tf.split(tf.prng_key(0)),
                                  np.array([[4146024105, 967050713], [2718843009, 1272950319]], np.uint32))


def test_uniform_normal_known_values():
    assert tf.uniform(tf.prng_key(0), (1,))[0] == np.float32(0.41845703)
    assert tf.normal(tf.prng_key(0), (1,))[0] == np.float32(-0.20584226)
    np.testing.assert_array_equal(tf.normal(tf.prng_key(42), (3,)),
                                  np.array([0.18693547, -1.2806505, -1.5593132], np.float32))


def test_odd_sizes_and_shapes():
    a = tf.random_bits(tf.prng_key(7), (5,))
    b = tf.random_bits(tf.prng_key(7), (6,))
    assert a.shape == (5,) and b.shape == (6,)
    # odd length pads the counter array with one zero before splitting in halves
    o0, o1 = tf.threefry2x32_block(0, 7, np.array([0, 1, 2], np.uint32), np.array([3, 4, 0], np.uint32))
    np.testing.assert_array_equal(a, np.concatenate([o0, o1])[:-1])
    assert tf.normal(tf.prng_key(3), (2, 3, 4)).shape == (2, 3, 4)


def test_randint_range_and_uniform_degenerate():
    r = tf.randint(tf.prng_key(5), (1000,), 1, 1001)
    assert r.min() >= 1 and r.max() <= 1000 and r.dtype == np.int32
    # minval > maxval collapses to minval
    lo, hi = np.float32(0.9), np.float32(0.8)
    u = tf.uniform(tf.prng_key(1), (64,), lo, hi)
    assert np.all(u == lo)

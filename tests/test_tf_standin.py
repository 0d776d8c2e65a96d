"""Conformance of the numpy TensorFlow stand-in (oracle/ref_exec/stubs/tensorflow) with the DOCUMENTED semantics of the TF 1.x ops the
reference uses: the worked examples of the TF API docs (scatter_nd, pad, tile, gather, concat / stack, strided slices, SAME padding
rule) and independent formulations (torch max-pool / conv with explicit pads, numpy einsum).  The vectors in
tests/golden/ref_exec_v1.npz are only as good as this layer, so it gets its own tests.  CPU only; runs in a sub-interpreter-free way by
importing the stand-in under a private module name (the product never sees a module called `tensorflow`)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

STUBS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'ref_exec', 'stubs')


@pytest.fixture(scope='module')
def tf():
    """Import the stand-in as `tensorflow` for the duration of this module only, then remove every trace of it from sys.modules."""
    before = set(sys.modules)
    sys.path.insert(0, STUBS)
    try:
        mod = importlib.import_module('tensorflow')
        assert 'numpy-standin' in mod.__version__
        yield mod
    finally:
        sys.path.remove(STUBS)
        for name in set(sys.modules) - before:
            if name == 'tensorflow' or name.startswith('tensorflow.'):
                del sys.modules[name]


def run(tf, t, feed=None):
    return tf.Session().run(t, feed)


def test_documented_examples_of_core_ops(tf):
    tf.reset_default_graph()
    # tf.scatter_nd, API-doc example: indices [[4],[3],[1],[7]], updates [9,10,11,12], shape [8]
    out = run(tf, tf.scatter_nd(tf.constant([[4], [3], [1], [7]]), tf.constant([9., 10., 11., 12.]), [8]))
    assert out.tolist() == [0, 11, 0, 10, 9, 0, 0, 12]
    out = run(tf, tf.scatter_nd(tf.constant([[1], [1], [3]]), tf.constant([1., 2., 5.]), [4]))           # duplicates accumulate
    assert out.tolist() == [0, 3, 0, 5]
    # tf.pad, API-doc example
    t = tf.constant([[1, 2, 3], [4, 5, 6]])
    assert run(tf, tf.pad(t, [[1, 1], [2, 2]])).tolist() == [[0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 2, 3, 0, 0], [0, 0, 4, 5, 6, 0, 0],
                                                             [0, 0, 0, 0, 0, 0, 0]]
    # tf.tile: [a b c d] x [2] -> [a b c d a b c d]; 2-D multiples repeat whole blocks
    assert run(tf, tf.tile(tf.constant([1, 2, 3, 4]), [2])).tolist() == [1, 2, 3, 4, 1, 2, 3, 4]
    assert run(tf, tf.tile(tf.constant([[1, 2], [3, 4]]), [2, 1])).tolist() == [[1, 2], [3, 4], [1, 2], [3, 4]]
    # tf.concat / tf.stack, API-doc examples
    t1, t2 = tf.constant([[1, 2, 3], [4, 5, 6]]), tf.constant([[7, 8, 9], [10, 11, 12]])
    assert run(tf, tf.concat([t1, t2], 0)).shape == (4, 3) and run(tf, tf.concat([t1, t2], 1)).shape == (2, 6)
    x, y, z = tf.constant([1, 4]), tf.constant([2, 5]), tf.constant([3, 6])
    assert run(tf, tf.stack([x, y, z])).tolist() == [[1, 4], [2, 5], [3, 6]]
    assert run(tf, tf.stack([x, y, z], axis=1)).tolist() == [[1, 2, 3], [4, 5, 6]]
    # tf.gather along an axis; tf.where(cond, x, y); tf.expand_dims / squeeze; tf.eye; tf.range
    p = tf.constant(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    assert np.array_equal(run(tf, tf.gather(params=p, indices=tf.constant([2, 0]), axis=1)), np.arange(24.).reshape(2, 3, 4)[:, [2, 0]])
    assert run(tf, tf.where(tf.constant([True, False, True]), tf.constant([1., 2., 3.]), tf.constant([9., 9., 9.]))).tolist() == [1, 9, 3]
    assert run(tf, tf.expand_dims(tf.constant([1., 2.]), -1)).shape == (2, 1)
    assert run(tf, tf.squeeze(tf.constant(np.zeros((2, 1, 1, 5), np.float32)), axis=[1, 2])).shape == (2, 5)
    assert run(tf, tf.range(0, 4) * 9).tolist() == [0, 9, 18, 27] and run(tf, tf.range(0, 4) * 9).dtype == np.int32
    # strided slices incl. negative starts (models.py:351 `[:, -10:]`, :355 `[:, 3:3 + 72]`) and an int index dropping an axis
    a = np.arange(2 * 85, dtype=np.float32).reshape(2, 85)
    ta = tf.constant(a)
    assert np.array_equal(run(tf, ta[:, -10:]), a[:, -10:]) and np.array_equal(run(tf, ta[:, 3:3 + 72]), a[:, 3:75])
    assert np.array_equal(run(tf, ta[0, :3]), a[0, :3])


def test_dtype_and_shape_rules(tf):
    tf.reset_default_graph()
    x = tf.constant(np.ones((3, 2), np.float32))
    assert run(tf, x + 1e-8).dtype == np.float32 and run(tf, 1 - x).dtype == np.float32      # python scalars take the tensor's dtype
    assert x.shape[0].value == 3 and x.shape.as_list() == [3, 2] and x.shape[1:] == (2,) and len(x.shape) == 2
    assert (x.shape[0] * 4).value == 12 and int(4 * x.shape[0]) == 12
    assert tf.constant((), shape=(2, 0, 25, 3)).shape.as_list() == [2, 0, 25, 3]              # omega.py:31-34
    assert run(tf, tf.concat((tf.constant((), shape=(2, 0, 3)), tf.constant(np.ones((2, 4, 3), np.float32))), axis=1)).shape == (2, 4, 3)
    assert run(tf, tf.ones([x.shape[0], 1])).shape == (3, 1)                                  # Dimension objects inside a shape list
    s = tf.shape(x)
    assert run(tf, tf.reshape(x, [s[0], -1])).shape == (3, 2)                                 # projection.py:27-29: tensor-valued shape
    with pytest.raises(TypeError):
        bool(x)                                                                               # graph tensors have no truth value
    assert isinstance(x, tf.Tensor)
    # tf.div on integers floors (TF1), on floats divides
    assert run(tf, tf.div(tf.constant([7, -7]), tf.constant([2, 2]))).tolist() == [3, -4]
    assert np.allclose(run(tf, tf.div(tf.constant([7.]), tf.constant([2.]))), 3.5)


def test_matmul_norm_reductions_against_numpy(tf):
    tf.reset_default_graph()
    rng = np.random.RandomState(0)
    a, b = rng.normal(size=(4, 5, 3, 6)).astype(np.float32), rng.normal(size=(4, 5, 7, 6)).astype(np.float32)
    got = run(tf, tf.matmul(a=tf.constant(a), b=tf.constant(b), transpose_b=True))
    assert np.allclose(got, np.einsum('abik,abjk->abij', a, b), atol=1e-5)
    v = rng.normal(size=(6, 3)).astype(np.float32)
    assert np.allclose(run(tf, tf.norm(tf.constant(v) + 1e-8, axis=1)), np.sqrt(((v + np.float32(1e-8)) ** 2).sum(1)), rtol=1e-6)
    m = rng.normal(size=(5, 3, 3)).astype(np.float32)
    assert np.allclose(run(tf, tf.trace(tf.constant(m))), np.trace(m, axis1=1, axis2=2), atol=1e-6)
    assert run(tf, tf.reduce_mean(tf.constant(a), [1, 2], keep_dims=True)).shape == (4, 1, 1, 6)
    assert np.allclose(run(tf, tf.clip_by_value(tf.constant([-2., 0.5, 3.]), -1, 1)), [-1, 0.5, 1])


def test_same_padding_rule_and_pooling_against_torch(tf):
    """TF 'SAME': out = ceil(in / stride), extra padding at the END -- pool1 112 -> 56 pads bottom / right only."""
    from tensorflow import _kernels
    assert _kernels.same_pads(112, 3, 2) == (0, 1) and _kernels.same_pads(20, 3, 1) == (1, 1) and _kernels.same_pads(7, 3, 2) == (1, 1)
    assert _kernels.same_pads(224, 7, 2) == (2, 3)              # (why slim's conv2d_same pads 3+3 explicitly instead: A.2)
    rng = np.random.RandomState(1)
    x = rng.normal(size=(2, 12, 12, 5)).astype(np.float32)
    tf.reset_default_graph()
    got = run(tf, tf.nn.max_pool(tf.constant(x), [1, 3, 3, 1], [1, 2, 2, 1], 'SAME'))
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    ref = torch.nn.functional.max_pool2d(torch.nn.functional.pad(xt, (0, 1, 0, 1), value=float('-inf')), 3, 2).permute(0, 2, 3, 1).numpy()
    assert got.shape == (2, 6, 6, 5) and np.array_equal(got, ref)
    w = rng.normal(size=(3, 1, 5, 4)).astype(np.float32)          # the temporal conv's kernel shape [3,1] (models.py:176)
    xc = rng.normal(size=(2, 9, 1, 5)).astype(np.float32)
    got = run(tf, tf.nn.conv2d(tf.constant(xc), tf.constant(w), [1, 1, 1, 1], 'SAME'))
    ref = np.zeros((2, 9, 1, 4), np.float32)
    xp = np.pad(xc, ((0, 0), (1, 1), (0, 0), (0, 0)))
    for t in range(9):
        ref[:, t, 0] = np.einsum('nkc,kco->no', xp[:, t:t + 3, 0], w[:, 0])
    assert np.allclose(got, ref, atol=1e-5)


def test_variable_scopes_reuse_and_saver(tf, tmp_path):
    tf.reset_default_graph()
    with tf.variable_scope('a'):
        with tf.variable_scope('b'):
            v = tf.get_variable('w', [2, 3])
        assert v.name == 'a/b/w:0' and v.op_name == 'a/b/w'
        with pytest.raises(ValueError):                           # same name again without reuse
            with tf.variable_scope('b'):
                tf.get_variable('w', [2, 3])
        with tf.variable_scope('b', reuse=tf.AUTO_REUSE):
            assert tf.get_variable('w', [2, 3]) is v              # AUTO_REUSE returns the existing variable (models.py:405)
            w2 = tf.get_variable('fresh', [1])                    # ... and creates missing ones
        with tf.variable_scope('b', reuse=True):
            with pytest.raises(ValueError):
                tf.get_variable('missing', [1])
    plain = tf.Variable(np.zeros((1, 85)), name='mean_param', dtype=tf.float32, trainable=True)
    assert plain.name == 'mean_param:0' and plain.probe.dtype == np.float32
    assert [x.op_name for x in tf.contrib.framework.get_variables('a/b')] == ['a/b/w', 'a/b/fresh']
    with pytest.raises(RuntimeError):                             # uninitialised variable, like FailedPreconditionError
        run(tf, v + 1.0)
    np.savez(str(tmp_path / 'ck.npz'), **{'a/b/w': np.arange(6, dtype=np.float32).reshape(2, 3), 'mean_param': np.ones((1, 85), np.float32)})
    with pytest.raises(KeyError):                                 # a variable missing from the checkpoint is an error, like NotFoundError
        tf.train.Saver([v, w2, plain]).restore(None, str(tmp_path / 'ck.npz'))
    tf.train.Saver([v, plain]).restore(None, str(tmp_path / 'ck.npz'))
    assert run(tf, v).tolist() == [[0, 1, 2], [3, 4, 5]] and float(run(tf, plain)[0, 0]) == 1.0
    with pytest.raises(ValueError):                               # shape mismatch on restore
        np.savez(str(tmp_path / 'bad.npz'), **{'a/b/w': np.zeros((3, 2), np.float32)})
        tf.train.Saver([v]).restore(None, str(tmp_path / 'bad.npz'))


def test_session_feeds_and_nested_fetches(tf):
    tf.reset_default_graph()
    pl = tf.placeholder(tf.float32, shape=(2, 3))
    y = tf.reduce_sum(pl * 2.0, axis=1)
    assert pl.shape.as_list() == [2, 3] and y.shape.as_list() == [2]                       # static shapes exist at graph-build time
    r = tf.Session().run({'a': y, 'b': [pl, {'c': y + 1.0}]}, {pl: np.ones((2, 3))})
    assert r['a'].tolist() == [6, 6] and r['b'][0].shape == (2, 3) and r['b'][1]['c'].tolist() == [7, 7]
    with pytest.raises(ValueError):
        tf.Session().run(y, {pl: np.ones((1, 3))})                                        # static placeholder shape (tester.py:64-66)
    with pytest.raises(ValueError):
        tf.Session().run(y)                                                                # unfed placeholder


def test_contrib_layers_against_independent_formulas(tf):
    """[TF-ext] layers of the stand-in vs the formulas SURVEY App. A states (A.5 BN, A.7 group_norm, A.8 fully_connected)."""
    tf.reset_default_graph()
    rng = np.random.RandomState(2)
    x = rng.normal(size=(2, 6, 1, 64)).astype(np.float32)
    y = tf.contrib.layers.group_norm(tf.constant(x), channels_axis=-1, reduction_axes=(-3, -2), scope='gn')
    names = {v.op_name: v for v in tf.global_variables()}
    assert set(names) == {'gn/beta', 'gn/gamma'}
    g, b = rng.uniform(0.5, 1.5, 64).astype(np.float32), rng.normal(size=64).astype(np.float32)
    names['gn/gamma'].load(g)
    names['gn/beta'].load(b)
    xr = x.astype(np.float64).reshape(2, 6, 1, 32, 2)
    mean, var = xr.mean(axis=(1, 2, 4), keepdims=True), xr.var(axis=(1, 2, 4), keepdims=True)
    ref = ((xr - mean) / np.sqrt(var + 1e-6)).reshape(2, 6, 1, 64) * g + b
    assert np.allclose(run(tf, y), ref, atol=2e-5)
    z = tf.contrib.slim.fully_connected(tf.constant(x[:, :, 0]), 8, scope='fc')             # rank 3 input: contraction over the last axis
    w, bias = rng.normal(size=(64, 8)).astype(np.float32), rng.normal(size=8).astype(np.float32)
    tf.contrib.framework.get_variables('fc')[0].load(w)
    tf.contrib.framework.get_variables('fc')[1].load(bias)
    assert [v.op_name for v in tf.contrib.framework.get_variables('fc')] == ['fc/weights', 'fc/biases']
    assert np.allclose(run(tf, z), np.maximum(x[:, :, 0] @ w + bias, 0), atol=1e-5)        # default activation is ReLU
    bn = tf.contrib.layers.batch_norm(tf.constant(x), is_training=False, scale=True, epsilon=1e-5, scope='bn')
    vs = {v.op_name.split('/')[-1]: v for v in tf.contrib.framework.get_variables('bn')}
    assert set(vs) == {'beta', 'gamma', 'moving_mean', 'moving_variance'}
    vals = {'beta': b, 'gamma': g, 'moving_mean': rng.normal(size=64).astype(np.float32), 'moving_variance': rng.uniform(0.5, 1.5, 64).astype(np.float32)}
    for k, v in vs.items():
        v.load(vals[k])
    ref = g * (x - vals['moving_mean']) / np.sqrt(vals['moving_variance'] + 1e-5) + b
    assert np.allclose(run(tf, bn), ref, atol=1e-5)

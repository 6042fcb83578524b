"""Known-answer tests of oracle/tf1_shim against the TensorFlow 1.13 API documentation.

The golden vectors of tests/golden/ are produced by running the reference's own files over this shim, so the one
thing between "the reference's logic" and "the reference's numbers" is the shim's restatement of each tf.* op.  Every
case below is either the worked example of the op's TF 1.13 documentation page or a property that page states
(tie order of top_k, floor semantics of mod, row selection of a rank-1 where, the band of matrix_band_part, the
truncation of the truncated normal ...), for every op the reference's path calls (func.py, models/transformer*.py,
modules/rpr.py, modules/rela.py, search.py, utils/util.py).  CPU only.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "tf1_shim")


@pytest.fixture(scope="module")
def tf():
    sys.path.insert(0, SHIM)
    try:
        import tensorflow as shim
        assert shim.__version__.endswith("-shim")
        yield shim
    finally:
        sys.path.remove(SHIM)
        for name in [m for m in sys.modules if m == "tensorflow" or m.startswith("tensorflow.")]:
            del sys.modules[name]


def _np(t):
    return np.asarray(torch.Tensor.detach(t).numpy() if isinstance(t, torch.Tensor) else t)


def test_matrix_band_part_documentation_example(tf):
    x = tf.constant([[0, 1, 2, 3], [-1, 0, 1, 2], [-2, -1, 0, 1], [-3, -2, -1, 0]])
    np.testing.assert_array_equal(_np(tf.matrix_band_part(x, 1, -1)),
                                  [[0, 1, 2, 3], [-1, 0, 1, 2], [0, -1, 0, 1], [0, 0, -1, 0]])
    np.testing.assert_array_equal(_np(tf.matrix_band_part(x, 2, 1)),
                                  [[0, 1, 0, 0], [-1, 0, 1, 0], [-2, -1, 0, 1], [0, -2, -1, 0]])
    # func.attention_bias "causal": band_part(ones, -1, 0) is the lower triangle, diagonal included
    np.testing.assert_array_equal(_np(tf.matrix_band_part(tf.ones([3, 3]), -1, 0)), np.tril(np.ones((3, 3))))
    # batched: the band applies to the two innermost axes
    y = tf.ones([2, 3, 4])
    np.testing.assert_array_equal(_np(tf.matrix_band_part(y, 0, -1))[1], np.triu(np.ones((3, 4))))


def test_pad_tile_documentation_examples(tf):
    t = tf.constant([[1, 2, 3], [4, 5, 6]])
    np.testing.assert_array_equal(_np(tf.pad(t, [[1, 1], [2, 2]])),
                                  [[0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 2, 3, 0, 0], [0, 0, 4, 5, 6, 0, 0], [0, 0, 0, 0, 0, 0, 0]])
    # the shifted decoder input of models/transformer.py:119: one zero row in front of axis 1, last row dropped later
    x = tf.constant(np.arange(12, dtype=np.float32).reshape(1, 3, 4))
    p = _np(tf.pad(x, [[0, 0], [1, 0], [0, 0]]))
    assert p.shape == (1, 4, 4) and (p[0, 0] == 0).all() and (p[0, 1:] == _np(x)[0]).all()
    np.testing.assert_array_equal(_np(tf.tile(tf.constant([[1, 2], [3, 4]]), [2, 3])),
                                  np.tile(np.array([[1, 2], [3, 4]]), (2, 3)))


def test_gather_and_gather_nd_documentation_examples(tf):
    params = tf.constant([[10, 11], [20, 21]])
    np.testing.assert_array_equal(_np(tf.gather_nd(params, tf.constant([[0, 0], [1, 1]]))), [10, 21])
    np.testing.assert_array_equal(_np(tf.gather_nd(params, tf.constant([[1], [0]]))), [[20, 21], [10, 11]])
    # batched coordinates, the way search.py gathers (batch, beam) pairs: indices [B, K, 2] -> [B, K, ...]
    p3 = tf.constant(np.arange(24).reshape(2, 3, 4))
    idx = tf.constant([[[0, 2], [0, 0]], [[1, 1], [1, 2]]])
    np.testing.assert_array_equal(_np(tf.gather_nd(p3, idx)), [[np.arange(8, 12), np.arange(0, 4)],
                                                              [np.arange(16, 20), np.arange(20, 24)]])
    emb = tf.constant(np.arange(15, dtype=np.float32).reshape(5, 3))
    ids = tf.constant([[4, 0], [2, 2]])
    np.testing.assert_array_equal(_np(tf.gather(emb, ids)), _np(emb)[np.array([[4, 0], [2, 2]])])


def test_top_k_sorted_descending_and_ties_keep_the_lower_index_first(tf):
    v, i = tf.nn.top_k(tf.constant([[1.0, 3.0, 3.0, 2.0, 3.0], [5.0, 5.0, 5.0, 5.0, 5.0]]), k=3)
    np.testing.assert_array_equal(_np(v), [[3, 3, 3], [5, 5, 5]])
    np.testing.assert_array_equal(_np(i), [[1, 2, 4], [0, 1, 2]])
    # -inf everywhere (a finished sentence's continuations in search.py): still the first k columns, in order
    v, i = tf.nn.top_k(tf.constant([[-np.inf] * 6]), k=4)
    np.testing.assert_array_equal(_np(i), [[0, 1, 2, 3]])
    rng = np.random.default_rng(0)
    x = rng.integers(0, 4, (7, 50)).astype(np.float32)            # many ties
    v, i = tf.nn.top_k(tf.constant(x), k=10)
    order = np.argsort(-x, axis=-1, kind="stable")[:, :10]
    np.testing.assert_array_equal(_np(i), order)
    np.testing.assert_array_equal(_np(v), np.take_along_axis(x, order, -1))


def test_where_selects_rows_with_a_rank_one_condition(tf):
    cond = tf.constant([True, False, True])
    x = tf.constant(np.ones((3, 2, 2), np.float32))
    y = tf.constant(np.zeros((3, 2, 2), np.float32))
    out = _np(tf.where(cond, x, y))
    assert (out[0] == 1).all() and (out[1] == 0).all() and (out[2] == 1).all()
    # same-shape condition: element-wise
    np.testing.assert_array_equal(_np(tf.where(tf.constant([[True, False]]), tf.constant([[1, 2]]), tf.constant([[3, 4]]))), [[1, 4]])


def test_boolean_mask_documentation_examples(tf):
    np.testing.assert_array_equal(_np(tf.boolean_mask(tf.constant([0, 1, 2, 3]), tf.constant([True, False, True, False]))), [0, 2])
    np.testing.assert_array_equal(_np(tf.boolean_mask(tf.constant([[1, 2], [3, 4], [5, 6]]), tf.constant([True, False, True]))),
                                  [[1, 2], [5, 6]])


def test_one_hot_documentation_examples(tf):
    np.testing.assert_array_equal(_np(tf.one_hot(tf.constant([0, 1, 2]), 3)), np.eye(3))
    oh = tf.one_hot(tf.constant([0, 2]), 3, on_value=5.0, off_value=-1.0)
    np.testing.assert_array_equal(_np(oh), [[5, -1, -1], [-1, -1, 5]])
    assert _np(oh).dtype == np.float32
    # utils/util.label_smooth: one_hot(labels, V, on_value=1-eps, off_value=eps/(V-1)) sums to one
    V, eps = 7, 0.1
    s = _np(tf.one_hot(tf.constant([[3, 0]]), V, on_value=1.0 - eps, off_value=eps / (V - 1)))
    np.testing.assert_allclose(s.sum(-1), 1.0, rtol=1e-6)
    assert s.shape == (1, 2, V) and abs(s[0, 0, 3] - 0.9) < 1e-7


def test_split_concat_stack_squeeze_expand_transpose(tf):
    x = tf.constant(np.arange(30).reshape(5, 6))
    a, b, c = tf.split(x, 3, axis=1)
    assert a.shape == (5, 2) and (_np(c) == _np(x)[:, 4:]).all()
    a, b, c = tf.split(x, [1, 3, 2], axis=1)
    assert (a.shape, b.shape, c.shape) == ((5, 1), (5, 3), (5, 2)) and (_np(b) == _np(x)[:, 1:4]).all()
    np.testing.assert_array_equal(_np(tf.concat([a, b, c], 1)), _np(x))
    assert tuple(tf.stack([x, x], axis=1).shape) == (5, 2, 6)
    assert tuple(tf.expand_dims(x, -1).shape) == (5, 6, 1) and tuple(tf.expand_dims(x, 0).shape) == (1, 5, 6)
    assert tuple(tf.squeeze(tf.ones([1, 3, 1])).shape) == (3,) and tuple(tf.squeeze(tf.ones([1, 3, 1]), 0).shape) == (3, 1)
    y = tf.constant(np.arange(24).reshape(2, 3, 4))
    np.testing.assert_array_equal(_np(tf.transpose(y)), np.transpose(_np(y)))                 # default: reversed axes
    np.testing.assert_array_equal(_np(tf.transpose(y, [0, 2, 1])), np.transpose(_np(y), (0, 2, 1)))
    np.testing.assert_array_equal(_np(tf.shape(y)), [2, 3, 4])
    assert y.get_shape().as_list() == [2, 3, 4] and y.shape.ndims == 3


def test_mod_is_floored_and_integer_cast_truncates(tf):
    # tf.mod == tf.floormod: the result takes the sign of the divisor
    np.testing.assert_array_equal(_np(tf.mod(tf.constant([-7, 7, -1, 5]), 3)), [2, 1, 2, 2])
    assert tf.mod(7, 3) == 1
    # tf.cast float -> int32 truncates toward zero; tf.to_float widens exactly
    np.testing.assert_array_equal(_np(tf.cast(tf.constant([1.8, -1.8, 2.0, -0.2]), tf.int32)), [1, -1, 2, 0])
    assert _np(tf.to_float(tf.constant([3, -4]))).dtype == np.float32
    # search.py: beam index = flat index // vocabulary, word id = flat index % vocabulary (non-negative operands)
    flat = tf.constant([0, 31999, 32000, 95999])
    np.testing.assert_array_equal(_np(flat // 32000), [0, 0, 1, 2])
    np.testing.assert_array_equal(_np(flat % 32000), [0, 31999, 0, 31999])


def test_reductions_documentation_examples(tf):
    x = tf.constant([[1.0, 1.0, 1.0], [1.0, 1.0, 1.0]])
    assert float(tf.reduce_sum(x)) == 6.0
    np.testing.assert_array_equal(_np(tf.reduce_sum(x, 0)), [2, 2, 2])
    np.testing.assert_array_equal(_np(tf.reduce_sum(x, 1, keepdims=True)), [[3], [3]])
    np.testing.assert_array_equal(_np(tf.reduce_sum(x, [0, 1])), 6)
    np.testing.assert_array_equal(_np(tf.reduce_sum(x, axis=1, keep_dims=True)), [[3], [3]])   # the TF1 spelling
    m = tf.constant([[1.0, 1.0], [2.0, 2.0]])
    assert float(tf.reduce_mean(m)) == 1.5
    np.testing.assert_array_equal(_np(tf.reduce_mean(m, 0)), [1.5, 1.5])
    np.testing.assert_array_equal(_np(tf.reduce_mean(m, 1)), [1.0, 2.0])
    z = tf.constant([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    np.testing.assert_allclose(float(tf.reduce_logsumexp(z)), math.log(6.0), rtol=1e-6)
    np.testing.assert_allclose(_np(tf.reduce_logsumexp(z, 0)), [math.log(2.0)] * 3, rtol=1e-6)
    np.testing.assert_allclose(_np(tf.reduce_logsumexp(z, 1, keepdims=True)), [[math.log(3.0)]] * 2, rtol=1e-6)
    b = tf.constant([[True, True], [False, False]])
    assert not bool(tf.reduce_all(b)) and bool(tf.reduce_any(b))
    np.testing.assert_array_equal(_np(tf.reduce_all(b, 1)), [True, False])
    np.testing.assert_array_equal(_np(tf.reduce_any(b, 0)), [True, True])
    np.testing.assert_array_equal(_np(tf.reduce_max(tf.constant([[1.0, 5.0], [7.0, 2.0]]), 1)), [5, 7])
    np.testing.assert_array_equal(_np(tf.reduce_min(tf.constant([[1.0, 5.0], [7.0, 2.0]]), 0)), [1, 2])


def test_cumsum_range_fill_clip(tf):
    np.testing.assert_array_equal(_np(tf.cumsum(tf.constant([1.0, 2.0, 3.0]))), [1, 3, 6])
    np.testing.assert_array_equal(_np(tf.cumsum(tf.constant([[1, 2], [3, 4]]), axis=1)), [[1, 3], [3, 7]])
    np.testing.assert_array_equal(_np(tf.range(4)), [0, 1, 2, 3])
    np.testing.assert_array_equal(_np(tf.range(3, 18, 3)), [3, 6, 9, 12, 15])
    np.testing.assert_array_equal(_np(tf.fill([2, 3], 9)), np.full((2, 3), 9))
    np.testing.assert_array_equal(_np(tf.clip_by_value(tf.constant([-3, 0, 5]), -1, 2)), [-1, 0, 2])
    np.testing.assert_array_equal(_np(tf.eye(2)), np.eye(2))
    np.testing.assert_array_equal(_np(tf.zeros_like(tf.constant([[1, 2]]))), [[0, 0]])
    assert float(tf.add_n([tf.constant(1.0), tf.constant(2.0), tf.constant(4.0)])) == 7.0


def test_matmul_transposes_and_batched_broadcast(tf):
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((2, 3, 4)).astype(np.float32), rng.standard_normal((2, 5, 4)).astype(np.float32)
    np.testing.assert_allclose(_np(tf.matmul(tf.constant(a), tf.constant(b), transpose_b=True)),
                               np.einsum("bik,bjk->bij", a, b), rtol=1e-5, atol=1e-6)
    c = rng.standard_normal((2, 3, 5)).astype(np.float32)
    np.testing.assert_allclose(_np(tf.matmul(tf.constant(a), tf.constant(c), transpose_a=True)),
                               np.einsum("bki,bkj->bij", a, c), rtol=1e-5, atol=1e-6)


def test_softmax_and_cross_entropy_known_answers(tf):
    x = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]], np.float32)
    e = np.exp(x - x.max(-1, keepdims=True))
    np.testing.assert_allclose(_np(tf.nn.softmax(tf.constant(x))), e / e.sum(-1, keepdims=True), rtol=1e-6)
    np.testing.assert_allclose(_np(tf.nn.softmax(tf.constant(x), dim=0)).sum(0), [1, 1, 1], rtol=1e-6)
    # softmax_cross_entropy_with_logits_v2(labels, logits) = -sum(labels * log_softmax(logits)), soft labels allowed
    labels = np.array([[0.0, 0.0, 1.0], [0.2, 0.3, 0.5]], np.float32)
    want = -(labels * np.log(e / e.sum(-1, keepdims=True))).sum(-1)
    got = _np(tf.nn.softmax_cross_entropy_with_logits_v2(labels=tf.constant(labels), logits=tf.constant(x)))
    np.testing.assert_allclose(got, want, rtol=1e-6)
    np.testing.assert_allclose(got[1], math.log(3.0), rtol=1e-6)
    np.testing.assert_allclose(got[0], math.log(1 + math.exp(-1) + math.exp(-2)), rtol=1e-6)


def test_dropout_keeps_with_keep_prob_and_rescales(tf):
    x = tf.ones([200, 200])
    assert tf.nn.dropout(x, 1.0) is x or bool((tf.nn.dropout(x, 1.0) == 1).all())
    tf.set_random_seed(5)
    y = _np(tf.nn.dropout(x, 0.8))
    assert set(np.unique(y).round(6)) == {0.0, 1.25}                     # kept elements are scaled by 1 / keep_prob
    assert abs((y != 0).mean() - 0.8) < 0.01 and abs(y.mean() - 1.0) < 0.02


def test_control_flow_runs_the_python_callables(tf):
    out = tf.while_loop(lambda i, acc: i < 5, lambda i, acc: (i + 1, acc + i), [tf.constant(0), tf.constant(0)])
    assert [int(v) for v in out] == [5, 10]
    assert tf.cond(tf.constant(True), lambda: 1, lambda: 2) == 1 and tf.cond(tf.constant(False), lambda: 1, lambda: 2) == 2
    # back_prop=False loops (search.py) build no autograd graph
    w = tf.constant(2.0)
    w.requires_grad_(True)
    (acc,) = tf.while_loop(lambda a: a < 10.0, lambda a: (a * w,), [tf.constant(1.0)], back_prop=False)
    assert float(acc) == 16.0 and not acc.requires_grad


def test_variable_scopes_prefix_names_reuse_and_inherit(tf):
    tf.reset_default_graph(seed=3)
    calls = []

    def getter(true_getter, name, *a, **kw):
        calls.append(name)
        return true_getter(name, *a, **kw)

    with tf.variable_scope("model", initializer=tf.zeros_initializer(), custom_getter=getter):
        with tf.variable_scope("encoder"):
            with tf.variable_scope("layer_0", initializer=tf.ones_initializer()):
                w = tf.get_variable("W", [2, 3])
                assert tf.get_variable_scope().name == "model/encoder/layer_0"
            b = tf.get_variable("b", [3])
        with tf.variable_scope("encoder", reuse=tf.AUTO_REUSE):
            with tf.variable_scope("layer_0"):
                w2 = tf.get_variable("W", [2, 3])
    assert w2 is w and calls == ["model/encoder/layer_0/W", "model/encoder/b", "model/encoder/layer_0/W"]
    assert list(tf.all_variables().keys()) == ["model/encoder/layer_0/W", "model/encoder/b"]
    assert bool((w == 1).all()) and bool((b == 0).all())          # nearest enclosing initializer wins
    assert w.requires_grad and [v is w or v is b for v in tf.trainable_variables()] == [True, True]
    tf.reset_default_graph()
    assert len(tf.all_variables()) == 0 and tf.get_variable_scope().name == ""


@pytest.mark.parametrize("mode,distribution", [("fan_in", "uniform"), ("fan_avg", "uniform"), ("fan_out", "normal"),
                                               ("fan_in", "truncated_normal"), ("fan_avg", "untruncated_normal")])
def test_variance_scaling_initializer_moments_and_support(tf, mode, distribution):
    tf.reset_default_graph(seed=11)
    shape, scale = [300, 500], 1.7
    fan = {"fan_in": 300.0, "fan_out": 500.0, "fan_avg": 400.0}[mode]
    x = _np(tf.variance_scaling_initializer(scale, mode, distribution)(shape)).astype(np.float64)
    want_std = math.sqrt(scale / fan)
    assert abs(x.mean()) < 4 * want_std / math.sqrt(x.size)
    assert abs(x.std() / want_std - 1.0) < 0.01, (x.std(), want_std)      # every distribution has variance scale / n
    if distribution == "uniform":
        limit = math.sqrt(3.0 * scale / fan)
        assert np.abs(x).max() <= limit and np.abs(x).max() > 0.999 * limit
    elif distribution in ("normal", "truncated_normal"):
        # truncated at two standard deviations of the UNDERLYING normal (stddev / 0.87962566103423978)
        cut = 2.0 * want_std / 0.87962566103423978
        assert np.abs(x).max() <= cut and np.abs(x).max() > 0.98 * cut
    else:
        assert np.abs(x).max() > 3.5 * want_std


def test_glorot_uniform_is_the_default_initializer_of_get_variable(tf):
    tf.reset_default_graph(seed=2)
    w = _np(tf.get_variable("W", [400, 200]))
    limit = math.sqrt(6.0 / 600.0)
    assert np.abs(w).max() <= limit and abs(w.std() - limit / math.sqrt(3.0)) < 0.01 * limit
    # convolution-style shapes: the receptive field multiplies both fans
    k = _np(tf.glorot_uniform_initializer()([3, 3, 16, 32]))
    assert np.abs(k).max() <= math.sqrt(6.0 / (9 * 16 + 9 * 32))
    tf.reset_default_graph()


def test_comparisons_and_logical_ops(tf):
    a, b = tf.constant([1, 2, 3]), tf.constant([3, 2, 1])
    np.testing.assert_array_equal(_np(tf.equal(a, b)), [False, True, False])
    np.testing.assert_array_equal(_np(tf.less(a, b)), [True, False, False])
    np.testing.assert_array_equal(_np(tf.less_equal(a, b)), [True, True, False])
    np.testing.assert_array_equal(_np(tf.greater(a, b)), [False, False, True])
    np.testing.assert_array_equal(_np(tf.greater_equal(a, b)), [False, True, True])
    t, f = tf.constant([True, True, False]), tf.constant([True, False, False])
    np.testing.assert_array_equal(_np(tf.logical_and(t, f)), [True, False, False])
    np.testing.assert_array_equal(_np(tf.logical_or(t, f)), [True, True, False])
    np.testing.assert_array_equal(_np(tf.logical_not(f)), [False, True, True])


def test_dtype_objects_follow_the_reference_usage(tf):
    # utils/dtype.py: tf.as_dtype(name), .min / .max for the additive masks (func.py:106-118 uses dtype.inf())
    assert tf.as_dtype("float32") == tf.float32 and tf.as_dtype(tf.float16) == tf.float16
    assert tf.float32.max == np.finfo(np.float32).max and tf.float16.min == np.finfo(np.float16).min
    assert tf.int32.max == np.iinfo(np.int32).max
    assert _np(tf.constant(1.5)).dtype == np.float32 and _np(tf.constant([1, 2])).dtype == np.int64
    assert _np(tf.constant(1.5, dtype=tf.float64)).dtype == np.float64
    # int32 ids / indices are carried as int64 (same values; torch indexing wants int64), the limits are int32's
    assert _np(tf.ones([2], dtype=tf.int32)).dtype == np.int64 and tf.int32.min == -2 ** 31


def test_bias_add_pow_unary_ops_and_random_initializers(tf):
    x = tf.constant(np.arange(6, dtype=np.float32).reshape(1, 2, 3))
    np.testing.assert_array_equal(_np(tf.nn.bias_add(x, tf.constant([10.0, 20.0, 30.0])))[0, 1], [13, 24, 35])
    np.testing.assert_allclose(_np(tf.pow(tf.constant([2.0, 3.0]), 2.0)), [4, 9])
    np.testing.assert_allclose(_np(tf.pow(10000.0, tf.constant([0.0, 0.5]))), [1, 100], rtol=1e-6)   # timing signal scales
    v = np.array([0.25, 1.0, 4.0], np.float32)
    for name, fn in [("exp", np.exp), ("log", np.log), ("sin", np.sin), ("cos", np.cos), ("tanh", np.tanh),
                     ("rsqrt", lambda a: 1 / np.sqrt(a)), ("sigmoid", lambda a: 1 / (1 + np.exp(-a)))]:
        np.testing.assert_allclose(_np(getattr(tf, name)(tf.constant(v))), fn(v), rtol=1e-6, err_msg=name)
    np.testing.assert_array_equal(_np(tf.nn.relu(tf.constant([-1.0, 0.0, 2.0]))), [0, 0, 2])
    np.testing.assert_allclose(_np(tf.nn.sigmoid(tf.constant([0.0]))), [0.5])
    tf.reset_default_graph(seed=9)
    n = _np(tf.random_normal_initializer(0.5, 0.1)([400, 400])).astype(np.float64)
    assert abs(n.mean() - 0.5) < 1e-3 and abs(n.std() - 0.1) < 1e-3
    u = _np(tf.random_uniform_initializer(-0.3, 0.3)([400, 400])).astype(np.float64)
    assert u.min() >= -0.3 and u.max() < 0.3 and abs(u.std() - 0.6 / math.sqrt(12.0)) < 1e-3
    r = _np(tf.random_uniform([1000], minval=2, maxval=5))
    assert r.min() >= 2 and r.max() < 5
    assert bool((tf.zeros_initializer()([3]) == 0).all()) and bool((tf.ones_initializer()([3]) == 1).all())


def test_the_shim_answers_every_tf_name_the_reference_path_uses(tf):
    """`grep -oh "tf\\.[A-Za-z_.0-9]*"` over func.py, models/transformer{,_aan,_rpr,_rela,_fuse}.py, modules/rpr.py,
    modules/rela.py, modules/initializer.py, search.py, utils/util.py, utils/dtype.py of the reference checkout."""
    names = """variable_scope cast get_variable as_dtype shape reshape AUTO_REUSE float32 expand_dims matmul concat
        reduce_sum equal bool variance_scaling_initializer cond random_normal_initializer nn.bias_add gather zeros
        reduce_mean constant TensorShape zeros_like pad reduce_all range ones_like int32 gather_nd
        nn.softmax_cross_entropy_with_logits_v2 transpose stack reduce_any nn.softmax name_scope log where squeeze split
        rsqrt ones_initializer one_hot nn.top_k logging.warn gfile.Exists fill cumsum zeros_initializer tile sigmoid
        random_uniform pow nn.relu logging.info boolean_mask add_n while_loop trainable_variables to_float tanh sin
        reduce_min reduce_logsumexp random_uniform_initializer ones nn.sigmoid nn.dropout mod matrix_band_part
        logical_or logical_not logical_and less_equal less greater_equal greater glorot_uniform_initializer float32.min
        eye exp cos convert_to_tensor clip_by_value Session ConfigProto""".split()
    for name in names:
        obj = tf
        for part in name.split("."):
            assert hasattr(obj, part), "tf.%s is missing from the shim" % name
            obj = getattr(obj, part)
    ref = "/root/reference"
    if os.path.isdir(ref):          # build container only: the list above is complete
        import re
        used = set()
        for rel in ["func.py", "search.py", "utils/util.py", "utils/dtype.py", "modules/rpr.py", "modules/rela.py",
                    "modules/initializer.py"] + ["models/transformer%s.py" % s for s in ("", "_aan", "_rpr", "_rela", "_fuse")]:
            used |= set(re.findall(r"tf\.([A-Za-z_.0-9]*)", open(os.path.join(ref, rel)).read()))
        assert used <= set(names), sorted(used - set(names))

"""Known-answer tests of oracle/tf1_shim (the torch-backed stand-in for the TF-1.x calls the reference's model files make;
test infrastructure, see its header).  Each case is a semantic the golden vectors of oracle/gen_refgraph_golden.py rest on,
checked against the value TensorFlow documents or defines for it.  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def tf():
    shim = os.path.join(ROOT, 'oracle', 'tf1_shim')
    sys.path.insert(0, shim)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'tensorflow' or k.startswith('tensorflow.')}
    try:
        import tensorflow as t
        t.reset_default_graph()
        yield t
    finally:
        sys.path.remove(shim)
        for k in [k for k in sys.modules if k == 'tensorflow' or k.startswith('tensorflow.')]:
            del sys.modules[k]
        sys.modules.update(saved)


def _var(tf, name, value):
    v = tf.get_variable(name=name, shape=list(np.shape(value)), initializer=None)
    tf.INIT_OVERRIDE[name] = np.asarray(value, np.float32)
    return v


def test_clip_by_norm_rows(tf):
    # tf.clip_by_norm docs: t * clip_norm / l2norm(t) when l2norm > clip_norm, unchanged otherwise; axes=[1] -> per row
    x = _var(tf, 'x', [[3.0, 4.0], [0.3, 0.4], [0.0, 0.0]])
    with tf.Session() as s:
        s.run(tf.global_variables_initializer())
        out = s.run(tf.clip_by_norm(x, 2.5, axes=[1]))
    np.testing.assert_allclose(out, [[1.5, 2.0], [0.3, 0.4], [0.0, 0.0]], rtol=1e-6)


def test_l2_loss_relu_squared_difference_and_reductions(tf):
    a = tf.placeholder(tf.float32, shape=[None, None])
    with tf.Session() as s:
        x = np.array([[1.0, -2.0, 3.0], [0.0, 0.5, -0.5]], np.float32)
        assert s.run(tf.nn.l2_loss(a), {a: x}) == pytest.approx(0.5 * float((x ** 2).sum()))          # sum(t ** 2) / 2
        np.testing.assert_array_equal(s.run(tf.nn.relu(a), {a: x}), np.maximum(x, 0))
        np.testing.assert_allclose(s.run(tf.reduce_sum(tf.squared_difference(a, 1.0), reduction_indices=1), {a: x}),
                                   ((x - 1) ** 2).sum(1))
        np.testing.assert_allclose(s.run(tf.reduce_min(a, 1), {a: x}), x.min(1))
        np.testing.assert_allclose(s.run(tf.reduce_mean(tf.cast(a > 0, tf.float32), 1), {a: x}), (x > 0).mean(1))
        np.testing.assert_allclose(s.run(tf.reduce_sum(tf.expand_dims(a, 1) * tf.expand_dims(a, 0), reduction_indices=[1, 2]), {a: x}),
                                   (x[:, None, :] * x[None, :, :]).sum((1, 2)))


def test_top_k_is_sorted_and_breaks_ties_towards_the_lower_index(tf):
    a = tf.placeholder(tf.float32, shape=[None, None])
    with tf.Session() as s:
        vals, idx = s.run(tf.nn.top_k(a, 3), {a: np.array([[1.0, 5.0, 5.0, 0.0, 5.0], [2.0, 2.0, 1.0, 3.0, 2.0]], np.float32)})
    np.testing.assert_array_equal(idx, [[1, 2, 4], [3, 0, 1]])
    np.testing.assert_array_equal(vals, [[5, 5, 5], [3, 2, 2]])


def test_adagrad_known_answer_and_duplicate_rows_are_summed_before_the_apply(tf):
    # TF1 AdagradOptimizer: accumulator starts at 0.1; accum += g^2; var -= lr * g / sqrt(accum).  Duplicate indices of an
    # embedding_lookup gradient are summed first (optimizer.py::_apply_sparse_duplicate_indices); untouched rows stay put.
    v = _var(tf, 'v', [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    ids = tf.placeholder(tf.int32, shape=[None])
    loss = tf.reduce_sum(tf.nn.embedding_lookup(v, ids) * 2.0)         # d loss / d row = 2 per occurrence
    opt = tf.train.AdagradOptimizer(0.5)
    op = opt.minimize(loss, var_list=[v])
    with tf.Session() as s:
        s.run(tf.global_variables_initializer())
        s.run(op, {ids: [0, 0, 2]})                                    # row 0 twice: g = 4; row 2 once: g = 2; row 1: none
        out = s.run(v)
    want = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    want[0] -= 0.5 * 4 / np.sqrt(0.1 + 16)
    want[2] -= 0.5 * 2 / np.sqrt(0.1 + 4)
    np.testing.assert_allclose(out, want, rtol=1e-6)
    np.testing.assert_allclose(opt.accum[v].numpy(), [[16.1, 16.1], [0.1, 0.1], [4.1, 4.1]], rtol=1e-6)
    # a second optimizer instance (the reference builds one per evaluation of its __optimize__ property) has its own slots
    assert tf.train.AdagradOptimizer(0.5).accum == {}


def test_reduce_min_gradient_is_shared_by_tied_minima(tf):
    # math_grad.py::_MinOrMaxGrad: indicators / num_selected * grad
    v = _var(tf, 'v', [[2.0, 1.0, 1.0, 3.0]])
    op = tf.train.AdagradOptimizer(1.0, initial_accumulator_value=1.0).minimize(tf.reduce_sum(tf.reduce_min(v, 1)), var_list=[v])
    with tf.Session() as s:
        s.run(tf.global_variables_initializer())
        s.run(op)
        out = s.run(v)
    step = 0.5 / np.sqrt(1.0 + 0.25)
    np.testing.assert_allclose(out, [[2.0, 1.0 - step, 1.0 - step, 3.0]], rtol=1e-6)


def test_loss_fetched_beside_the_train_op_is_the_pre_update_loss_and_control_dependencies_order_the_clip(tf):
    # the reference's train_op = (self.__optimize__, self.__loss) and CML's "step, then clip" (cml.py:119-129)
    v = _var(tf, 'v', [[3.0, 4.0]])
    loss = lambda: tf.reduce_sum(v * v)
    gds = [tf.train.AdagradOptimizer(1.0, initial_accumulator_value=1.0).minimize(loss(), var_list=[v])]
    with tf.control_dependencies(gds):
        clip = [tf.assign(v, tf.clip_by_norm(v, 1.0, axes=[1]))]
    with tf.Session() as s:
        s.run(tf.global_variables_initializer())
        _, l = s.run((gds + [clip], loss()))
        out = s.run(v)
    assert l == pytest.approx(25.0)                                     # evaluated at the initial value
    stepped = np.array([3.0 - 6 / np.sqrt(37.0), 4.0 - 8 / np.sqrt(65.0)])
    np.testing.assert_allclose(out, [stepped / np.linalg.norm(stepped)], rtol=1e-6)    # the clip saw the UPDATED row


def test_slicing_operators_and_python_scalars(tf):
    p = tf.placeholder(tf.int32, shape=[None, 3])
    f = tf.placeholder(tf.float32, shape=[None])
    with tf.Session() as s:
        x = np.array([[1, 2, 3], [4, 5, 6]])
        np.testing.assert_array_equal(s.run(p[:, 0], {p: x}), [1, 4])
        np.testing.assert_array_equal(s.run(p[:, 1:], {p: x}), [[2, 3], [5, 6]])
        y = np.array([0.5, -1.0], np.float32)
        np.testing.assert_allclose(s.run(-tf.log(tf.sigmoid(f)) * 2 + 1.0 - f / 2, {f: y}),
                                   -np.log(1 / (1 + np.exp(-y))) * 2 + 1.0 - y / 2, rtol=1e-6)
        np.testing.assert_allclose(s.run(tf.add(tf.divide(f, tf.cast(4, tf.float32)), 0.25 * f), {f: y}), y / 2, rtol=1e-6)
        assert s.run(tf.matmul(tf.expand_dims(f, 0), tf.transpose(tf.expand_dims(f, 0))), {f: y}).item() == pytest.approx(1.25)

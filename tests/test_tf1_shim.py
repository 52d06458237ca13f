"""Self-tests of the TensorFlow-1.4 API stand-in (oracle/tf1_shim) that runs the reference's model code for the golden
fixtures.  Each check states a TF r1.4 behaviour the reference relies on and verifies the stand-in against an independent,
written-out computation (scalar loops / hand arithmetic) — not against the oracle, which the fixtures are meant to pin."""
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "tf1_shim")


@pytest.fixture()
def tf():
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    import tensorflow as tf_
    tf_.set_float_precision(True)                # float64: the checks below are exact to rounding
    tf_.reset_default_graph()
    yield tf_
    tf_.set_float_precision(False)
    tf_.reset_default_graph()
    # leave no `tensorflow` importable behind: libraries that probe for it (find_spec) must keep seeing none
    if SHIM in sys.path and not any(m in sys.modules for m in ("models.tacotron",)):
        sys.path.remove(SHIM)
        for name in [m for m in sys.modules if m == "tensorflow" or m.startswith("tensorflow.")]:
            del sys.modules[name]


def _t(tf, a):
    return tf.Tensor(torch.as_tensor(np.asarray(a, dtype=np.float64)))


def test_variable_scope_naming_rules(tf):
    """Unnamed tf.layers calls take dense, dense_1, ... per enclosing scope; named ones keep their name; a cell object owns
    one scope for all its calls (the while_loop body is traced once in TF)."""
    x = _t(tf, np.ones((2, 3)))
    with tf.variable_scope("inference"):
        tf.layers.dense(x, 4)
        tf.layers.dense(x, 4)
        with tf.variable_scope("prenet"):
            tf.layers.dense(x, 4, name="dense_1")
            tf.layers.dense(x, 4, name="dense_1")          # same variables (get-or-create), not dense_1_1
        with tf.variable_scope("cbhg"):
            tf.layers.dense(x, 4)                           # counting restarts per scope
        cell = tf.contrib.rnn.GRUCell(5)
        h = cell.zero_state(2, tf.float32)
        with tf.variable_scope("decoder"):
            for _ in range(3):
                _, h = cell(x, h)
        tf.layers.dense(x, 4)
    names = list(tf.shim_state().vars)
    assert names == ["inference/dense/kernel", "inference/dense/bias", "inference/dense_1/kernel", "inference/dense_1/bias",
                     "inference/prenet/dense_1/kernel", "inference/prenet/dense_1/bias", "inference/cbhg/dense/kernel",
                     "inference/cbhg/dense/bias", "inference/decoder/gru_cell/gates/kernel", "inference/decoder/gru_cell/gates/bias",
                     "inference/decoder/gru_cell/candidate/kernel", "inference/decoder/gru_cell/candidate/bias",
                     "inference/dense_2/kernel", "inference/dense_2/bias"]
    assert float(tf.shim_state().vars["inference/decoder/gru_cell/gates/bias"].t.detach()[0]) == 1.0     # GRUCell gate bias starts at 1


def test_conv1d_same_padding_and_max_pool(tf):
    rng = np.random.RandomState(0)
    x = rng.randn(2, 7, 3)
    for k in (1, 2, 3, 4, 5):
        tf.reset_default_graph()
        y = tf.layers.conv1d(_t(tf, x), filters=2, kernel_size=k, padding="same").numpy()
        W = tf.shim_state().vars["conv1d/kernel"].t.detach().numpy()          # [k, in, out]
        left = (k - 1) // 2                                                    # SAME: the extra pad goes to the end
        want = np.zeros((2, 7, 2))
        for t in range(7):
            for j in range(k):
                s = t + j - left
                if 0 <= s < 7:
                    want[:, t] += x[:, s] @ W[j]
        assert np.abs(y - want).max() < 1e-12, k
    # max_pooling1d(pool 2, stride 1, 'same'): m[t] = max(x[t], x[t+1]), last frame alone; ties -> first element owns the gradient
    v = torch.tensor([[[1.0], [3.0], [3.0], [2.0]]], dtype=torch.float64, requires_grad=True)
    m = tf.layers.max_pooling1d(tf.Tensor(v), pool_size=2, strides=1, padding="same")
    assert m.numpy().reshape(-1).tolist() == [3.0, 3.0, 3.0, 2.0]
    m.t.sum().backward()
    assert v.grad.reshape(-1).tolist() == [0.0, 2.0, 1.0, 1.0]


def test_batch_normalization_training_and_update_ops(tf):
    rng = np.random.RandomState(1)
    x = rng.randn(3, 5, 4) * 2 + 1
    y = tf.layers.batch_normalization(_t(tf, x), training=True).numpy()
    mean, var = x.reshape(-1, 4).mean(0), x.reshape(-1, 4).var(0)              # biased variance over all but the channel axis
    assert np.abs(y - (x - mean) / np.sqrt(var + 1e-3)).max() < 1e-12
    st = tf.shim_state().vars
    assert float(st["batch_normalization/moving_mean"].t.abs().max()) == 0.0  # nothing moves until UPDATE_OPS runs
    with tf.control_dependencies(tf.get_collection(tf.GraphKeys.UPDATE_OPS)):
        pass
    assert np.abs(st["batch_normalization/moving_mean"].t.numpy() - 0.01 * mean).max() < 1e-12
    assert np.abs(st["batch_normalization/moving_variance"].t.numpy() - (0.99 + 0.01 * var)).max() < 1e-12
    assert not st["batch_normalization/moving_mean"].trainable and st["batch_normalization/gamma"].trainable
    y2 = tf.layers.batch_normalization(_t(tf, x), training=False, name="batch_normalization").numpy()   # inference: moving statistics
    mm, mv = st["batch_normalization/moving_mean"].t.numpy(), st["batch_normalization/moving_variance"].t.numpy()
    assert np.abs(y2 - (x - mm) / np.sqrt(mv + 1e-3)).max() < 1e-12


def test_gru_cell_and_length_aware_bidirectional_rnn(tf):
    rng = np.random.RandomState(2)
    N, T, C, H = 3, 6, 4, 5
    x = rng.randn(N, T, C)
    L = np.array([6, 2, 4])
    fw, bw = tf.contrib.rnn.GRUCell(H), tf.contrib.rnn.GRUCell(H)
    (of, ob), (sf, sb) = tf.nn.bidirectional_dynamic_rnn(fw, bw, _t(tf, x), sequence_length=tf.Tensor(torch.tensor(L, dtype=torch.int32)),
                                                         dtype=tf.float32)
    st = {k: v.t.detach().numpy() for k, v in tf.shim_state().vars.items()}
    sig = lambda z: 1 / (1 + np.exp(-z))  # noqa: E731

    def run(prefix, seq):                   # scalar-loop GRUCell: reset gate multiplies h BEFORE the candidate matmul
        Wg, bg, Wc, bc = (st[prefix + s] for s in ("gates/kernel", "gates/bias", "candidate/kernel", "candidate/bias"))
        h = np.zeros(H); outs = []
        for xt in seq:
            g = sig(np.concatenate([xt, h]) @ Wg + bg)
            r, u = g[:H], g[H:]
            c = np.tanh(np.concatenate([xt, r * h]) @ Wc + bc)
            h = u * h + (1 - u) * c
            outs.append(h)
        return np.array(outs), h
    for n in range(N):
        o, h = run("bidirectional_rnn/fw/gru_cell/", x[n, :L[n]])
        assert np.abs(of.numpy()[n, :L[n]] - o).max() < 1e-12 and np.abs(of.numpy()[n, L[n]:]).sum() == 0      # zeros past the length
        assert np.abs(sf.numpy()[n] - h).max() < 1e-12                                                         # state carried through
        o, h = run("bidirectional_rnn/bw/gru_cell/", x[n, :L[n]][::-1])
        assert np.abs(ob.numpy()[n, :L[n]] - o[::-1]).max() < 1e-12 and np.abs(ob.numpy()[n, L[n]:]).sum() == 0


def test_monotonic_attention_parallel_equals_recursion_and_initial_alignments(tf):
    rng = np.random.RandomState(3)
    p = 1 / (1 + np.exp(-rng.randn(2, 9)))
    prev = rng.rand(2, 9); prev /= prev.sum(1, keepdims=True)
    par = tf.contrib.seq2seq.monotonic_attention(_t(tf, p), _t(tf, prev), "parallel").numpy()
    rec = tf.contrib.seq2seq.monotonic_attention(_t(tf, p), _t(tf, prev), "recursive").numpy()
    q = np.zeros(2); want = np.zeros_like(p)                                   # Raffel et al.: q_j = (1-p_{j-1}) q_{j-1} + prev_j
    for j in range(9):
        q = (q * (1 - p[:, j - 1]) if j else q) + prev[:, j]
        want[:, j] = p[:, j] * q
    assert np.abs(par - want).max() < 1e-12 and np.abs(rec - want).max() < 1e-12
    mech = tf.contrib.seq2seq.BahdanauMonotonicAttention(4, _t(tf, rng.randn(2, 9, 6)))
    assert mech.initial_alignments(2, tf.float32).numpy().tolist() == [[1.0] + [0.0] * 8] * 2        # one-hot on position 0
    soft = tf.contrib.seq2seq.BahdanauAttention(4, _t(tf, rng.randn(2, 9, 6)))
    assert soft.initial_alignments(2, tf.float32).numpy().sum() == 0.0
    assert "memory_layer/kernel" in tf.shim_state().vars                     # keys are computed once, at construction


def test_dynamic_decode_runs_until_maximum_iterations(tf):
    class Cell(tf.contrib.rnn.RNNCell):
        state_size = 1
        output_size = 2

        def call(self, inputs, state):
            return tf.concat([inputs, state], -1), state + 1.0

    class Helper(tf.contrib.seq2seq.Helper):
        batch_size = 2

        def initialize(self, name=None):
            return tf.tile([False], [2]), tf.zeros([2, 1])

        def sample(self, time, outputs, state, name=None):
            return tf.tile([0], [2])

        def next_inputs(self, time, outputs, state, sample_ids, name=None):
            return tf.reduce_all(tf.equal(outputs, 123.0), axis=1), outputs[:, -1:], state     # never finishes by itself

    dec = tf.contrib.seq2seq.BasicDecoder(Cell(), Helper(), tf.zeros([2, 1]))
    (out, _), final_state, lengths = tf.contrib.seq2seq.dynamic_decode(dec, maximum_iterations=5)
    assert out.numpy().shape == (2, 5, 2) and lengths.numpy().tolist() == [5, 5] and final_state.numpy().tolist() == [[5.0], [5.0]]
    assert out.numpy()[0, :, 1].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]


def test_clip_by_global_norm_and_adam_first_step(tf):
    w = tf.get_variable("w", [3], initializer=tf.constant_initializer(1.0))
    v = tf.get_variable("v", [2], initializer=tf.constant_initializer(-2.0))
    loss = tf.reduce_sum(w * _t(tf, [3.0, 0.0, 4.0])) + tf.reduce_sum(v * _t(tf, [0.0, 12.0]))
    opt = tf.train.AdamOptimizer(0.5, 0.9, 0.999)
    grads, variables = zip(*opt.compute_gradients(loss))
    assert [g.numpy().tolist() for g in grads] == [[3.0, 0.0, 4.0], [0.0, 12.0]]
    clipped, norm = tf.clip_by_global_norm(grads, 1.0)
    assert abs(float(norm) - 13.0) < 1e-12 and np.abs(clipped[0].numpy() - np.array([3, 0, 4]) / 13).max() < 1e-12
    small, _ = tf.clip_by_global_norm([_t(tf, [0.3, 0.4])], 1.0)
    assert np.abs(small[0].numpy() - [0.3, 0.4]).max() < 1e-15                # below the threshold: untouched
    step = tf.Variable(0, name="global_step", trainable=False)
    opt.apply_gradients(zip(clipped, variables), global_step=step)
    g = 3.0 / 13
    lr_t = 0.5 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = 1.0 - lr_t * (0.1 * g) / (math.sqrt(0.001 * g * g) + 1e-8)         # epsilon outside the bias-corrected sqrt
    assert abs(float(w.t[0]) - want) < 1e-12 and float(w.t[1]) == 1.0 and int(step) == 1
    assert abs(float(tf.train.exponential_decay(1.0, _t(tf, 6000.0), 3000, 0.95)) - 0.95 ** 2) < 1e-12

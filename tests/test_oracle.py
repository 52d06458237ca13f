"""Self-tests that pin the CPU oracle (the reference ships no tests or golden vectors; SURVEY.md §8c items i-ix)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tacotron_oracle as O
from oracle import griffin_lim_oracle as G


def test_monotonic_parallel_matches_recursion_fp64():
    g = torch.Generator().manual_seed(0)
    # keep cumprod(1-p) above the 1e-10 clip of the closed form (beyond it TF's 'parallel' mode itself deviates from
    # the recursion; the parity target is the clipped closed form)
    p = torch.sigmoid(torch.randn(4, 40, generator=g, dtype=torch.float64) * 0.5 - 2.0)
    prev = torch.softmax(torch.randn(4, 40, generator=g, dtype=torch.float64), -1)
    a = O.monotonic_attention_parallel(p, prev)
    b = O.monotonic_attention_recursive(p, prev)
    assert (a - b).abs().max() < 1e-9
    p_hard = torch.sigmoid(torch.randn(4, 40, generator=g, dtype=torch.float64) * 2.0)
    a2, b2 = O.monotonic_attention_parallel(p_hard, prev), O.monotonic_attention_recursive(p_hard, prev)
    assert (a2[:, :8] - b2[:, :8]).abs().max() < 1e-9          # agreement away from the clip region
    assert (a >= 0).all() and (a.sum(-1) <= 1 + 1e-9).all()


def test_monotonic_limits():
    one_hot = F.one_hot(torch.zeros(2, dtype=torch.long), 10).double()
    stay = O.monotonic_attention_parallel(torch.ones(2, 10, dtype=torch.float64), one_hot)
    assert torch.allclose(stay, one_hot)                       # p == 1: attention stays on position 0
    gone = O.monotonic_attention_parallel(torch.zeros(2, 10, dtype=torch.float64), one_hot)
    assert gone.abs().max() == 0                               # p == 0: the mass vanishes


def _gru_scalar(x, h, Wg, bg, Wc, bc):
    H = h.shape[0]
    xh = np.concatenate([x, h])
    r = np.array([1 / (1 + math.exp(-(xh @ Wg[:, j] + bg[j]))) for j in range(H)])
    u = np.array([1 / (1 + math.exp(-(xh @ Wg[:, H + j] + bg[H + j]))) for j in range(H)])
    xrh = np.concatenate([x, r * h])
    c = np.array([math.tanh(xrh @ Wc[:, j] + bc[j]) for j in range(H)])
    return u * h + (1 - u) * c


def test_gru_cell_against_scalar_loops_and_differs_from_torch_gru():
    g = torch.Generator().manual_seed(1)
    I, H = 5, 7
    P = {"g/gates_kernel": torch.randn(I + H, 2 * H, generator=g, dtype=torch.float64), "g/gates_bias": torch.ones(2 * H, dtype=torch.float64),
         "g/cand_kernel": torch.randn(I + H, H, generator=g, dtype=torch.float64), "g/cand_bias": torch.zeros(H, dtype=torch.float64)}
    x, h = torch.randn(1, I, generator=g, dtype=torch.float64), torch.randn(1, H, generator=g, dtype=torch.float64)
    got = O.gru_cell(x, h, P, "g")[0].numpy()
    ref = _gru_scalar(x[0].numpy(), h[0].numpy(), P["g/gates_kernel"].numpy(), P["g/gates_bias"].numpy(),
                      P["g/cand_kernel"].numpy(), P["g/cand_bias"].numpy())
    assert np.abs(got - ref).max() < 1e-12
    # cuDNN/torch GRU applies the reset gate AFTER the candidate matmul: same weights give a different state
    Wc = P["g/cand_kernel"]; Wg = P["g/gates_kernel"]
    r = torch.sigmoid(torch.cat([x, h], -1) @ Wg[:, :H] + 1)
    u = torch.sigmoid(torch.cat([x, h], -1) @ Wg[:, H:] + 1)
    c_torch_style = torch.tanh(x @ Wc[:I] + r * (h @ Wc[I:]))
    other = (u * h + (1 - u) * c_torch_style)[0].numpy()
    assert np.abs(other - ref).max() > 1e-3


def test_conv_maxpool_dense_bn_match_torch_functional():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 11, 6, generator=g)
    for k in (1, 2, 3, 4, 7):
        w, b = torch.randn(k, 6, 5, generator=g), torch.randn(5, generator=g)
        got = O.conv1d_same(x, w, b)
        ref = F.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), b, padding="same").transpose(1, 2)
        assert (got - ref).abs().max() < 1e-5
    mp = O.maxpool_same_2(x)
    ref = F.max_pool1d(F.pad(x.transpose(1, 2), (0, 1), value=-float("inf")), 2, 1).transpose(1, 2)
    assert torch.equal(mp, ref)
    P = {"s/gamma": torch.rand(6) + 0.5, "s/beta": torch.randn(6), "s/moving_mean": torch.randn(6), "s/moving_var": torch.rand(6) + 0.5}
    ev = O.batch_norm(x, P, "s", False, None)
    ref = F.batch_norm(x.reshape(-1, 6), P["s/moving_mean"], P["s/moving_var"], P["s/gamma"], P["s/beta"], False, 0.0, 1e-3).reshape(x.shape)
    assert (ev - ref).abs().max() < 1e-5
    st = {}
    tr = O.batch_norm(x, P, "s", True, st)
    flat = ((tr - P["s/beta"]) / P["s/gamma"]).reshape(-1, 6)
    assert flat.mean(0).abs().max() < 1e-5 and (flat.var(0, unbiased=False) - 1).abs().max() < 2e-2
    var = x.reshape(-1, 6).var(0, unbiased=False)
    assert torch.allclose(st["s/moving_var"], P["s/moving_var"] * 0.99 + var * 0.01, atol=1e-6)


def test_maxpool_tie_gradient_goes_to_first_element():
    x = torch.zeros(1, 3, 1, requires_grad=True)
    O.maxpool_same_2(x).sum().backward()
    assert x.grad.flatten().tolist() == [1.0, 1.0, 1.0]        # windows (0,1),(1,2),(2,-inf): first max wins each


def test_bidirectional_length_handling(hp5, tb):
    g = torch.Generator().manual_seed(3)
    H, C = 4, 4
    P = {}
    for d in ("fw", "bw"):
        P["c/gru_%s/gates_kernel" % d] = torch.randn(C + H, 2 * H, generator=g)
        P["c/gru_%s/gates_bias" % d] = torch.ones(2 * H)
        P["c/gru_%s/cand_kernel" % d] = torch.randn(C + H, H, generator=g)
        P["c/gru_%s/cand_bias" % d] = torch.zeros(H)
    x = torch.randn(2, 6, C, generator=g)
    L = torch.tensor([6, 3])
    out = O.bidirectional_gru(x, L, P, "c")
    assert out[1, 3:].abs().max() == 0                          # outputs past the length are zero
    # backward direction of the short row == forward pass of the bw cell over the reversed prefix
    ref = O.dynamic_rnn(x[1:2, :3].flip(1), None, torch.zeros(1, H), P, "c/gru_bw").flip(1)
    assert (out[1:2, :3, H:] - ref).abs().max() < 1e-6


def test_teacher_forcing_indexing_and_shapes(hp5, tb):
    P = tb.params.init_params(hp5, 1, seed=5)
    g = torch.Generator().manual_seed(4)
    N, Ti, To = 2, 9, 15
    inp = torch.randint(2, 80, (N, Ti), generator=g)
    L = torch.tensor([9, 6])
    mel, lin = torch.rand(N, To, 80, generator=g), torch.rand(N, To, 1025, generator=g)
    out = O.forward(P, hp5, inp, L, 1, None, mel, lin, speaker_mode="none", want_taps=True)
    assert out["mel_outputs"].shape == (N, To, 80) and out["linear_outputs"].shape == (N, To, 1025)
    assert out["alignments"].shape == (N, Ti, To // 5)
    assert out["taps"]["memory"].shape == (N, Ti, 256) and out["taps"]["post_outputs"].shape == (N, To, 512)
    # changing a target frame that is never fed back (index not = r-1 mod r) must not change the decoder outputs
    mel2 = mel.clone(); mel2[:, 7] += 1.0                      # 7 % 5 != 4
    out2 = O.forward(P, hp5, inp, L, 1, None, mel2, lin, speaker_mode="none")
    assert torch.equal(out["mel_outputs"], out2["mel_outputs"])
    mel3 = mel.clone(); mel3[:, 4] += 1.0                      # frame r-1 feeds step 1
    out3 = O.forward(P, hp5, inp, L, 1, None, mel3, lin, speaker_mode="none")
    assert torch.equal(out["mel_outputs"][:, :5], out3["mel_outputs"][:, :5])
    assert not torch.equal(out["mel_outputs"][:, 5:10], out3["mel_outputs"][:, 5:10])


def test_parameter_inventory_matches_survey(hp5, tb):
    specs = tb.params.param_specs(hp5, 1)
    assert sum(s.numel for s in specs) == 9336610
    assert sum(s.numel for s in specs if not s.trainable) == 9376


def test_oracle_gradients_finite_difference_fp64(tb):
    hp = tb.hparams.override(reduction_factor=5, enc_bank_size=3, post_bank_size=2, enc_highway_depth=1, post_highway_depth=1)
    P = {k: v.double() for k, v in tb.params.init_params(hp, 1, seed=11).items()}
    g = torch.Generator().manual_seed(5)
    N, Ti, To = 2, 5, 10
    inp = torch.randint(2, 80, (N, Ti), generator=g); L = torch.tensor([5, 4])
    mel, lin = torch.rand(N, To, 80, generator=g, dtype=torch.float64), torch.rand(N, To, 1025, generator=g, dtype=torch.float64)
    names = ["attention/v", "attention_gru/cand_bias", "mel_proj/bias", "enc_cbhg/gru_bw/gates_bias", "attention/score_bias"]

    def loss_of(Pd):
        out = O.forward(Pd, hp, inp, L, 1, None, mel, lin, speaker_mode="none")
        return O.losses(out, mel, lin, torch.ones(N, dtype=torch.float64), hp)["loss"]

    leaf = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in P.items()}
    grads = torch.autograd.grad(loss_of(leaf), [leaf[k] for k in names])
    for k, gk in zip(names, grads):
        idx = tuple(0 for _ in P[k].shape)
        eps = 1e-6
        Pp = dict(P); Pm = dict(P)
        tp = P[k].clone(); tp[idx] += eps; Pp[k] = tp
        tm = P[k].clone(); tm[idx] -= eps; Pm[k] = tm
        fd = (float(loss_of(Pp)) - float(loss_of(Pm))) / (2 * eps)
        assert abs(fd - float(gk[idx])) < 1e-6 + 1e-4 * abs(fd), (k, fd, float(gk[idx]))


def test_adam_step_and_lr_schedule(hp5):
    P = {"w": torch.tensor([1.0, -2.0])}
    g = {"w": torch.tensor([0.5, -0.25])}
    m = {"w": torch.zeros(2)}; v = {"w": torch.zeros(2)}
    P2, m2, v2 = O.adam_step(dict(P), g, m, v, 1, 0.1)
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    exp = P["w"] - lr_t * (0.1 * g["w"]) / ((0.001 * g["w"] ** 2).sqrt() + 1e-8)
    assert torch.allclose(P2["w"], exp, atol=1e-7)
    assert abs(O.learning_rate(hp5, 0, True) - 0.002 * 4000 ** 0.5 * 4000 ** -1.5) < 1e-12
    assert abs(O.learning_rate(hp5, 3999, True) - 0.002) < 1e-9
    assert abs(O.learning_rate(hp5, 39999, False) - 0.002) < 1e-9
    clipped, gn = O.clip_by_global_norm({"a": torch.tensor([3.0, 4.0])}, 1.0)
    assert abs(gn - 5.0) < 1e-9 and torch.allclose(clipped["a"], torch.tensor([0.6, 0.8]))


def test_griffin_lim_oracle_roundtrip_and_convergence():
    n_fft, hop, win = G.stft_parameters()
    assert (n_fft, hop, win) == (2048, 300, 1200)
    rng = np.random.RandomState(0)
    y = rng.randn(hop * 20).astype(np.float32) * 0.1
    S = G.stft(y, n_fft, hop, win)
    assert S.shape == (1025, 21)
    back = G.istft(S, hop, win)
    assert np.abs(back - y).max() < 1e-4
    x = rng.randn(500)
    ref = np.zeros_like(x); acc = 0.0
    for i in range(500):
        acc = x[i] + 0.97 * acc; ref[i] = acc
    assert np.abs(G.lfilter_inv_preemphasis(x.astype(np.float32)) - ref).max() < 1e-3
    # spectral convergence improves with iterations
    mag = np.abs(S).T                                            # [T, F] magnitudes of a real signal
    spec = np.clip((20 * np.log10(np.maximum(1e-5, mag)) - 20 + 100) / 100, 0, 1).astype(np.float32)
    errs = []
    for iters in (0, 5, 20):
        w = G.inv_spectrogram(spec, rng.rand(*spec.shape).astype(np.float32), n_iters=iters, power=1.0, preemphasis=0.0)
        est = np.abs(G.stft(w, n_fft, hop, win)).T
        tgt = np.power(10.0, (spec * 100 - 100 + 20) * 0.05)
        errs.append(np.linalg.norm(est - tgt) / np.linalg.norm(tgt))
    assert errs[2] < errs[1] < errs[0]

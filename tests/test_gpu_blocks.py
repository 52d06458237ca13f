"""Block-level parity (SURVEY.md 8b, rows E1-E7 / P1 / D0-D9 on their own): the C ABI's operator and block entry points and their
Python mirrors of the reference's models/modules.py against the CPU oracle's block functions and its `taps`."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu

from oracle import tacotron_oracle as O  # noqa: E402

LIM = {"fp32": 2e-5, "tf32": 6e-3, "bf16": 2e-2}


def _modules():
    from importlib import import_module
    return import_module("multi-speaker-tacotron-tensorflow_b200.models.modules")


def _dev(P):
    return {k: v.cuda() for k, v in P.items()}


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_prenet_highwaynet_conv1d_mirrors_vs_oracle_blocks(tb, hp5, prec):
    """modules.py:18-25 prenet, :105-120 highwaynet, :123-131 conv1d (activation BEFORE batch norm, 'same' padding with the extra
    frame on the right for even widths, training / inference statistics) through taco_gemm / taco_highway_combine / taco_batch_norm."""
    M = _modules()
    P = tb.params.init_params(hp5, 1, seed=5, randomize_bn_state=True)
    Pd = _dev(P)
    g = torch.Generator().manual_seed(1)
    lim = LIM[prec]
    x = torch.randn(3, 37, 256, generator=g)
    got = M.prenet(x.cuda(), True, Pd, "enc_prenet", precision=prec)
    assert (got.cpu() - O.prenet(x, P, "enc_prenet", 2)).abs().max().item() <= lim * 4
    xh = torch.randn(3, 37, 128, generator=g)
    got = M.highwaynet(xh.cuda(), Pd, "enc_cbhg/highway_2", precision=prec)
    assert (got.cpu() - O.highwaynet(xh, P, "enc_cbhg/highway_2")).abs().max().item() <= lim * 4
    for scope, cin, act, oact in (("enc_cbhg/bank_4", 128, "relu", torch.relu), ("enc_cbhg/bank_7", 128, "relu", torch.relu),
                                  ("post_cbhg/bank_2", 80, "relu", torch.relu), ("post_cbhg/proj_2", 256, None, None)):
        xc = torch.randn(3, 37, cin, generator=g)
        for training in (True, False):
            new_state = {}
            ref = O.conv1d_bn(xc, P, scope, oact, training, new_state)
            out, mean, var = M.conv1d(xc.cuda(), act, training, Pd, scope, precision=prec)
            assert (out.cpu() - ref).abs().max().item() <= lim * 20, (scope, training)      # BN divides by std ~ 0.1-1
            if training:       # the batch moments are the ones the moving-statistics update uses (momentum .99)
                mm = P[scope + "/moving_mean"] * 0.99 + mean.cpu() * 0.01
                assert (mm - new_state[scope + "/moving_mean"]).abs().max().item() <= lim


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("which", [0, 1])
def test_cbhg_block_forward_backward_vs_oracle(tb, hp5, prec, which):
    """taco_cbhg_forward / taco_cbhg_backward (modules.py:27-96) on their own: outputs, the oracle's taps (highway input, RNN input),
    input gradient, before_highway / initial-state gradients and every parameter gradient of the block."""
    M = _modules()
    scope = "post_cbhg" if which else "enc_cbhg"
    P = tb.params.init_params(hp5, 1, seed=9, randomize_bn_state=True)
    g = torch.Generator().manual_seed(2)
    N, Ti, To = 3, 19, 25
    T, cin = (To, 80) if which else (Ti, 128)
    x = torch.randn(N, T, cin, generator=g)
    lengths = None if which else torch.tensor([19, 11, 4], dtype=torch.int32)
    before = None if which else torch.randn(N, 128, generator=g) * 0.3
    h0 = None if which else torch.randn(N, 256, generator=g) * 0.3
    names = [k for k in P if k.startswith(scope) and not k.endswith(("moving_mean", "moving_var"))]
    leaf = {k: (P[k].clone().requires_grad_(True) if k in names else P[k]) for k in P}
    xr = x.clone().requires_grad_(True)
    br = before.clone().requires_grad_(True) if before is not None else None
    hr = h0.clone().requires_grad_(True) if h0 is not None else None
    taps = {}
    hp = hp5
    ref = O.cbhg(xr, lengths, True, leaf, scope, hp.post_bank_size if which else hp.enc_bank_size, 2,
                 hp.post_highway_depth if which else hp.enc_highway_depth, before_highway=br, rnn_init_state=hr, new_state={}, taps=taps)
    dy = torch.randn(ref.shape, generator=g)
    wanted = [xr] + ([br, hr] if not which else []) + [leaf[k] for k in names]
    grads = torch.autograd.grad((ref * dy).sum(), wanted, allow_unused=True)
    lim = LIM[prec]
    eng = tb.Engine(hp5, 1, precision=prec, named_params=P)
    eng.plan(N, Ti, To, training=True)
    out = M.cbhg(eng, x, lengths, True, scope, before_highway=before, encoder_rnn_init_state=h0)
    assert out.shape == ref.shape and (out.cpu() - ref.detach()).abs().max().item() <= lim * 10
    # the oracle's taps against the workspace regions of the block (padded time layout: valid rows only)
    geo = (To + hp.post_bank_size - 1, (hp.post_bank_size - 1) // 2) if which else (Ti + hp.enc_bank_size - 1, (hp.enc_bank_size - 1) // 2)
    valid = lambda name, C: eng.region(scope + "/" + name).view(N, geo[0], C)[:, geo[1]:geo[1] + T].cpu()
    Hh = 256 if which else 128
    assert (valid("hw_0" if which else "hw0", Hh) - taps[scope + "/highway_input"].detach()).abs().max().item() <= lim * 20
    assert (valid("hw_4", Hh) - taps[scope + "/rnn_input"].detach()).abs().max().item() <= lim * 20
    if lengths is not None:
        for n, L in enumerate(lengths.tolist()):
            assert out[n, L:].abs().max().item() == 0 if L < T else True          # padded steps emit zeros (modules.py:92)
    eng.grads.zero_()
    dx, db, dh = eng.cbhg_backward(which, dy, lengths, want_before=not which, want_init_state=not which)
    rel = lambda a, b: float((a.cpu().double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))
    glim = {"fp32": 2e-3, "tf32": 3e-2, "bf16": 8e-2}[prec]
    assert rel(dx, grads[0]) <= glim
    k0 = 1
    if not which:
        assert rel(db, grads[1]) <= glim and rel(dh, grads[2]) <= glim
        k0 = 3
    got = eng.named_gradients()
    a = torch.cat([got[k].cpu().reshape(-1) for k in names]).double()
    b = torch.cat([(grads[k0 + i] if grads[k0 + i] is not None else torch.zeros_like(P[k])).reshape(-1) for i, k in enumerate(names)]).double()
    cos = float(a @ b / (a.norm() * b.norm()))
    assert cos >= {"fp32": 0.99999, "tf32": 0.9994, "bf16": 0.996}[prec], cos
    others = [k for k in got if k not in names]
    assert all(got[k].abs().max().item() == 0 for k in others)                     # nothing outside the block is touched
    eng.close()


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_decoder_block_forward_backward_vs_oracle(tb, hp5, prec):
    """taco_decoder_forward / taco_decoder_backward (tacotron.py:127-214, rnn_wrappers.py:218-415, helpers.py) on a given encoder
    memory: mel outputs and alignments against the oracle's decoder fed with the same memory; gradient wrt the memory."""
    P = tb.params.init_params(hp5, 1, seed=13, randomize_bn_state=True)
    g = torch.Generator().manual_seed(3)
    N, Ti, To = 3, 14, 30
    tok = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
    L = torch.tensor([14, 9, 5], dtype=torch.int32)
    mel_t = torch.rand(N, To, 80, generator=g); lin_t = torch.rand(N, To, 1025, generator=g)
    names = [k for k in P if not k.endswith(("moving_mean", "moving_var"))]
    leaf = {k: (P[k].clone().requires_grad_(True) if k in names else P[k]) for k in P}
    ref = O.forward(leaf, hp5, tok, L, 1, None, mel_t, lin_t, speaker_mode="none", want_taps=True)
    mem = ref["taps"]["memory"]
    mem.retain_grad()
    dmel = torch.randn(N, To, 80, generator=g)
    (ref["mel_outputs"] * dmel).sum().backward()
    lim = LIM[prec]
    eng = tb.Engine(hp5, 1, precision=prec, named_params=P)
    out = eng.decoder_forward(mem.detach(), tok, L, None, mel_t, lin_t)
    assert (out["mel_outputs"].cpu() - ref["mel_outputs"].detach()).abs().max().item() <= lim * 10
    assert (out["alignments"].cpu() - ref["alignments"].detach()).abs().max().item() <= lim * 10
    eng.grads.zero_()
    d_mem = eng.decoder_backward(dmel)
    rel = float((d_mem.cpu().double() - mem.grad.double()).norm() / mem.grad.double().norm())
    assert rel <= {"fp32": 2e-3, "tf32": 5e-2}[prec], rel
    got = eng.named_gradients()
    dec = [k for k in names if k.split("/")[0] in ("attention", "attention_gru", "dec_prenet", "concat_proj", "dec_gru_1", "dec_gru_2", "mel_proj")]
    a = torch.cat([got[k].cpu().reshape(-1) for k in dec]).double()
    b = torch.cat([leaf[k].grad.reshape(-1) for k in dec]).double()
    cos = float(a @ b / (a.norm() * b.norm()))
    assert cos >= (0.99999 if prec == "fp32" else 0.9994), cos
    # free-running on the same memory (inference plan)
    with torch.no_grad():
        inf = O.forward(P, hp5, tok, L, 1, None, max_iters=5, speaker_mode="none", want_taps=True)
    out = eng.decoder_forward(inf["taps"]["memory"], tok, L, decoder_steps=5)
    assert (out["mel_outputs"].cpu() - inf["mel_outputs"]).abs().max().item() <= lim * 20
    eng.close()

"""Staged GPU-vs-oracle diagnostic (development aid; the judged checks live in tests/).  Prints max-abs errors per
stage instead of stopping at the first failure.  Usage: python tools/gpu_diag.py [small|c1|c2] [fp32|bf16]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb  # noqa: E402
from oracle import tacotron_oracle as O  # noqa: E402


def make_batch(N, Ti, To, lengths, seed=1234, F=1025, M=80):
    g = torch.Generator().manual_seed(seed)
    inp = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
    L = torch.tensor(lengths, dtype=torch.int32)
    for n in range(N):
        inp[n, L[n] - 1] = 1
        inp[n, L[n]:] = 0
    mel = torch.rand(N, To, M, generator=g)
    lin = torch.rand(N, To, F, generator=g)
    coeff = torch.rand(N, generator=g) + 0.5
    return dict(inputs=inp, input_lengths=L, mel_targets=mel, linear_targets=lin, loss_coeff=coeff)


def unpad(t, N, T, Tp, PL):
    return t.view(N, Tp, -1)[:, PL:PL + T]


def report(name, got, ref, tol=1e-4):
    got = got.detach().float().cpu()
    ref = ref.detach().float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    bad = not (err <= tol * max(1.0, scale))
    print("%-34s max|err| %.3e  (ref max %.3e) %s%s" % (name, err, scale, "FAIL" if bad else "ok",
                                                      "  NaN!" if torch.isnan(got).any() else ""), flush=True)
    return not bad


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    hp = tb.hparams.override(reduction_factor=5)
    if which == "small":
        N, Ti, To, lengths = 3, 13, 20, [13, 9, 5]
    elif which == "c1":
        N, Ti, To, lengths = 2, 50, 200, [50, 50]
    else:
        N, Ti, To = 32, 128, 800
        g = torch.Generator().manual_seed(7)
        lengths = torch.randint(96, 129, (N,), generator=g).tolist(); lengths[0] = 128
    batch = make_batch(N, Ti, To, lengths)
    named = tb.params.init_params(hp, 1, seed=4321, randomize_bn_state=True)
    # non-trivial gamma/beta/biases so every term is exercised
    g = torch.Generator().manual_seed(99)
    for k, v in named.items():
        if k.endswith("/gamma"):
            v.add_(torch.randn(v.shape, generator=g) * 0.2)
        elif k.endswith(("/beta", "/bias", "_bias", "score_bias")):
            v.add_(torch.randn(v.shape, generator=g) * 0.1)
    tol = 1e-4 if prec == "fp32" else 3e-2

    t0 = time.time()
    P0 = {k: v.clone() for k, v in named.items()}
    names = [k for k in P0 if not (k.endswith("moving_mean") or k.endswith("moving_var"))]
    leaf = {k: (P0[k].clone().requires_grad_(True) if k in names else P0[k]) for k in P0}
    out = O.forward(leaf, hp, batch["inputs"], batch["input_lengths"], 1, None, batch["mel_targets"], batch["linear_targets"],
                    speaker_mode="none", want_taps=True)
    ls = O.losses(out, batch["mel_targets"], batch["linear_targets"], batch["loss_coeff"], hp)
    grads = torch.autograd.grad(ls["loss"], [leaf[k] for k in names], allow_unused=True)
    ref_g = {k: (g_ if g_ is not None else torch.zeros_like(P0[k])) for k, g_ in zip(names, grads)}
    print("oracle fwd+bwd: %.1fs  loss=%.6f" % (time.time() - t0, float(ls["loss"])), flush=True)

    from importlib import import_module
    eng_mod = import_module("multi-speaker-tacotron-tensorflow_b200.engine")
    eng = eng_mod.Engine(hp, 1, precision=prec, named_params=named)
    dev = eng.dev
    b = {k: v.to(dev) for k, v in batch.items()}
    res = eng.forward(b["inputs"], b["input_lengths"], None, b["mel_targets"], b["linear_targets"], b["loss_coeff"])
    torch.cuda.synchronize()
    ok = True
    ge, gp = None, None
    E = {"Tp": Ti + hp.enc_bank_size - 1, "PL": (hp.enc_bank_size - 1) // 2}
    Pp = {"Tp": To + hp.post_bank_size - 1, "PL": (hp.post_bank_size - 1) // 2}
    taps = out["taps"]
    ok &= report("enc highway_input", unpad(eng.region("enc_cbhg/hw0"), N, Ti, E["Tp"], E["PL"]), taps["enc_cbhg/highway_input"], tol)
    ok &= report("enc rnn_input", unpad(eng.region("enc_cbhg/hw_4"), N, Ti, E["Tp"], E["PL"]), taps["enc_cbhg/rnn_input"], tol)
    ok &= report("memory", eng.region("enc_cbhg/rnn_out").view(N, Ti, -1), taps["memory"], tol)
    ok &= report("keys", eng.region("dec/keys").view(N, Ti, -1), taps["keys"], tol)
    ok &= report("alignments", res["alignments"], out["alignments"], tol)
    ok &= report("mel_outputs", res["mel_outputs"], out["mel_outputs"], tol)
    ok &= report("post highway_input", unpad(eng.region("post_cbhg/hw_0"), N, To, Pp["Tp"], Pp["PL"]), taps["post_cbhg/highway_input"], tol)
    ok &= report("post rnn_input", unpad(eng.region("post_cbhg/hw_4"), N, To, Pp["Tp"], Pp["PL"]), taps["post_cbhg/rnn_input"], tol)
    ok &= report("post_outputs", eng.region("post_cbhg/rnn_out").view(N, To, -1), taps["post_outputs"], tol)
    ok &= report("linear_outputs", res["linear_outputs"], out["linear_outputs"], tol)

    eng.backward()
    torch.cuda.synchronize()
    sc = eng.scalars()
    print("scalars", sc, "oracle", {k: float(v) for k, v in ls.items()}, flush=True)
    got_g = eng.named_gradients()
    worst = []
    for k in names:
        gg = got_g[k].detach().float().cpu()
        rg = ref_g[k]
        denom = rg.norm().item() + 1e-12
        rel = (gg - rg).norm().item() / denom
        worst.append((rel, k, denom))
    worst.sort(reverse=True)
    gtol = 1e-3 if prec == "fp32" else 5e-2
    nbad = sum(1 for w in worst if not (w[0] <= gtol) and w[2] > 1e-9)
    print("gradients: %d tensors, %d above rel-L2 %.0e" % (len(worst), nbad, gtol))
    for rel, k, dn in worst[:25]:
        print("   %-48s rel %.3e  |ref| %.3e" % (k, rel, dn))
    ok &= nbad == 0

    # one optimizer step vs the oracle
    m0 = {k: torch.zeros_like(P0[k]) for k in names}
    v0 = {k: torch.zeros_like(P0[k]) for k in names}
    clipped, gn = O.clip_by_global_norm(ref_g, 1.0)
    lr = O.learning_rate(hp, 0, True)
    sub, _, _ = O.adam_step({k: P0[k].clone() for k in names}, clipped, m0, v0, 1, lr, hp.adam_beta1, hp.adam_beta2)
    eng.optimizer_step(True)
    torch.cuda.synchronize()
    sc = eng.scalars()
    print("grad_norm got %.6f ref %.6f   lr got %.6e ref %.6e" % (sc["grad_norm"], gn, sc["learning_rate"], lr))
    newp = eng.named_parameters()
    werr = max(((newp[k].float().cpu() - sub[k]).abs().max().item(), k) for k in names)
    print("param after Adam: worst abs err %.3e (%s)" % werr)
    berr = max(((newp[k].float().cpu() - out["new_bn_state"][k]).abs().max().item(), k) for k in out["new_bn_state"])
    print("BN moving stats: worst abs err %.3e (%s)" % berr)
    print("launches so far:", eng.launch_count())

    # timing of a full train step
    for _ in range(2):
        eng.train_step(b)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        eng.train_step(b)
    ev1.record(); torch.cuda.synchronize()
    print("train step: %.3f ms" % (ev0.elapsed_time(ev1) / 3))
    print("DIAG", "PASS" if ok else "FAIL")


if __name__ == "__main__":
    main()

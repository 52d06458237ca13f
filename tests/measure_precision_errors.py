"""Not a test: prints the errors of the reduced-precision modes (tf32, bf16) against the CPU oracle on the cases the GPU
parity tests use, so that the bounds written in tests/test_gpu_parity.py can be set to <= 2x what is measured.
Run on the GPU box:  python tests/measure_precision_errors.py [small|full|all]  > gpurun_out/precision_errors.txt"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import tacotron_b200 as tb  # noqa: E402
from oracle import tacotron_oracle as O  # noqa: E402


def batch(N, Ti, To, lengths, seed=1234):
    g = torch.Generator().manual_seed(seed)
    inp = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
    L = torch.tensor(lengths, dtype=torch.int32)
    for n in range(N):
        inp[n, L[n] - 1] = 1
        inp[n, L[n]:] = 0
    return dict(inputs=inp, input_lengths=L, mel_targets=torch.rand(N, To, 80, generator=g),
                linear_targets=torch.rand(N, To, 1025, generator=g), loss_coeff=torch.rand(N, generator=g) + 0.5)


def oracle_grads(named, hp, b, S, spk, mode):
    names = [k for k in named if not k.endswith(("moving_mean", "moving_var"))]
    leaf = {k: (named[k].clone().requires_grad_(True) if k in names else named[k]) for k in named}
    ref = O.forward(leaf, hp, b["inputs"], b["input_lengths"], S, spk, b["mel_targets"], b["linear_targets"], speaker_mode=mode)
    ls = O.losses(ref, b["mel_targets"], b["linear_targets"], b["loss_coeff"], hp)
    gl = torch.autograd.grad(ls["loss"], [leaf[k] for k in names], allow_unused=True)
    return ref, {k: float(v) for k, v in ls.items()}, {k: (gg if gg is not None else torch.zeros_like(named[k])) for k, gg in zip(names, gl)}


def metrics(out, ref):
    res = {}
    for k in ("mel_outputs", "linear_outputs", "alignments"):
        a, r = out[k].cpu().double(), ref[k].detach().double()
        res[k] = dict(maxabs=float((a - r).abs().max()), rel_l2=float((a - r).norm() / r.norm()))
    al, rl = out["alignments"].cpu(), ref["alignments"].detach()
    res["argmax_agree"] = float((al.argmax(1) == rl.argmax(1)).float().mean())     # attended input position per (row, decoder step)
    return res


def train_case(name, hp, S, mode, N, Ti, To, lengths, spk=None, seed=31, precs=("tf32", "bf16")):
    named = tb.params.init_params(hp, S, seed=seed, randomize_bn_state=True)
    b = batch(N, Ti, To, lengths)
    t0 = time.time()
    ref, ls, ref_g = oracle_grads(named, hp, b, S, spk, mode)
    print("== %s (oracle %.1fs)" % (name, time.time() - t0), flush=True)
    names = sorted(ref_g)
    rb = torch.cat([ref_g[k].reshape(-1) for k in names]).double()
    for prec in precs:
        eng = tb.Engine(hp, S, precision=prec, named_params=named)
        out = eng.forward(b["inputs"], b["input_lengths"], spk, b["mel_targets"], b["linear_targets"], b["loss_coeff"])
        mt = metrics(out, ref)
        eng.backward()
        sc = eng.scalars()
        got = eng.named_gradients()
        ga = torch.cat([got[k].cpu().reshape(-1) for k in names]).double()
        cos = float(ga @ rb / (ga.norm() * rb.norm()))
        worst = max(((got[k].cpu().double() - ref_g[k].double()).norm() / ref_g[k].double().norm()).item() for k in names if ref_g[k].norm() > 1e-7)
        print("  %-5s out maxabs mel %.2e lin %.2e align %.2e | rel-L2 mel %.2e lin %.2e align %.2e | argmax agree %.4f | loss rel %.2e | grad cos %.6f 1-cos %.2e norm rel %.2e worst tensor rel-L2 %.2e"
              % (prec, mt["mel_outputs"]["maxabs"], mt["linear_outputs"]["maxabs"], mt["alignments"]["maxabs"],
                 mt["mel_outputs"]["rel_l2"], mt["linear_outputs"]["rel_l2"], mt["alignments"]["rel_l2"], mt["argmax_agree"],
                 abs(sc["loss"] - ls["loss"]) / ls["loss"], cos, 1 - cos, abs(float(ga.norm() - rb.norm())) / float(rb.norm()), worst), flush=True)
        eng.close()


def infer_case(name, hp, S, mode, N, Ti, steps, lengths, spk=None, seed=43, precs=("fp32", "tf32", "bf16")):
    named = tb.params.init_params(hp, S, seed=seed, randomize_bn_state=True)
    b = batch(N, Ti, 5, lengths)
    t0 = time.time()
    with torch.no_grad():
        ref = O.forward(named, hp, b["inputs"], b["input_lengths"], S, spk, max_iters=steps, speaker_mode=mode)
    print("== %s (oracle %.1fs)" % (name, time.time() - t0), flush=True)
    for prec in precs:
        eng = tb.Engine(hp, S, precision=prec, named_params=named)
        out = eng.forward(b["inputs"], b["input_lengths"], spk, decoder_steps=steps)
        mt = metrics(out, ref)
        print("  %-5s out maxabs mel %.2e lin %.2e align %.2e | rel-L2 mel %.2e lin %.2e align %.2e | argmax agree %.4f"
              % (prec, mt["mel_outputs"]["maxabs"], mt["linear_outputs"]["maxabs"], mt["alignments"]["maxabs"],
                 mt["mel_outputs"]["rel_l2"], mt["linear_outputs"]["rel_l2"], mt["alignments"]["rel_l2"], mt["argmax_agree"]), flush=True)
        eng.close()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    hp5 = tb.hparams.override(reduction_factor=5)
    if what in ("small", "all"):
        train_case("ragged 3x13x20", hp5, 1, "none", 3, 13, 20, [13, 9, 5])
        train_case("C1 2x50x200", hp5, 1, "none", 2, 50, 200, [50, 50])
        hpd = tb.hparams.override(reduction_factor=5, model_type="deepvoice")
        train_case("deepvoice 5x12x15", hpd, 3, "deepvoice", 5, 12, 15, [12, 7, 12, 3, 9], spk=torch.tensor([0, 2, 1, 2, 0], dtype=torch.int32), seed=17)
        infer_case("infer 2x11 6 steps", hp5, 1, "none", 2, 11, 6, [11, 7])
    if what in ("full", "all"):
        g = torch.Generator().manual_seed(6)
        lengths = torch.randint(96, 129, (32,), generator=g).tolist(); lengths[0] = 128
        train_case("C2 32x128x800 single speaker", hp5, 1, "none", 32, 128, 800, lengths, seed=41)
        hpd = tb.hparams.override(reduction_factor=5, model_type="deepvoice", batch_size=32)
        spk = torch.randint(0, 3, (32,), generator=g, dtype=torch.int32)
        train_case("C3 32x128x800 deepvoice", hpd, 3, "deepvoice", 32, 128, 800, lengths, spk=spk, seed=41)
        infer_case("C4 1x128 200 steps", hp5, 1, "none", 1, 128, 200, [128])
        hpd4 = tb.hparams.override(reduction_factor=5, model_type="deepvoice")
        l5 = torch.randint(120, 201, (64,), generator=g).tolist(); l5[3] = 200
        spk5 = torch.randint(0, 4, (64,), generator=g, dtype=torch.int32)
        infer_case("C5 64x200 200 steps deepvoice", hpd4, 4, "deepvoice", 64, 200, 200, l5, spk=spk5, seed=47, precs=("fp32", "tf32", "bf16"))


if __name__ == "__main__":
    main()

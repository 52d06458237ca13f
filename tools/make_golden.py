"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot run here: TensorFlow 1.x / librosa are not
installable — see oracle/tacotron_oracle.py).  Deterministic: params from init_params(seed), inputs from a seeded generator.
Run:  python tools/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb  # noqa: E402
from oracle import tacotron_oracle as O  # noqa: E402
from oracle import griffin_lim_oracle as G  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_batch(N=2, Ti=11, To=15, seed=2024):
    g = torch.Generator().manual_seed(seed)
    inp = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
    L = torch.tensor([Ti, Ti - 4], dtype=torch.int32)[:N]
    for n in range(N):
        inp[n, L[n] - 1] = 1
        inp[n, L[n]:] = 0
    return dict(inputs=inp, input_lengths=L, mel_targets=torch.rand(N, To, 80, generator=g),
                linear_targets=torch.rand(N, To, 1025, generator=g), loss_coeff=torch.rand(N, generator=g) + 0.5)


def golden_params(hp, seed=99):
    named = tb.params.init_params(hp, 1, seed=seed, randomize_bn_state=True)
    g = torch.Generator().manual_seed(seed + 1)
    for k, v in named.items():
        if k.endswith("/gamma"):
            v.add_(torch.randn(v.shape, generator=g) * 0.2)
        elif k.endswith(("/beta", "/bias", "_bias", "score_bias")):
            v.add_(torch.randn(v.shape, generator=g) * 0.1)
    return named


def main():
    os.makedirs(OUT, exist_ok=True)
    hp = tb.hparams.override(reduction_factor=5)
    named = golden_params(hp)
    b = golden_batch()
    m = {k: torch.zeros_like(v) for k, v in named.items() if not k.endswith(("moving_mean", "moving_var"))}
    v = {k: torch.zeros_like(t) for k, t in m.items()}
    res = O.train_step({k: t.clone() for k, t in named.items()}, m, v, hp, b, 0, True, 1, "none")
    out = res["outputs"]
    names = sorted(res["grads"])
    np.savez_compressed(
        os.path.join(OUT, "tacotron_train_small.npz"),
        mel_outputs=out["mel_outputs"].detach().numpy(), linear_outputs=out["linear_outputs"].detach().numpy(),
        alignments=out["alignments"].detach().numpy(),
        scalars=np.array([res["loss"], res["mel_loss"], res["linear_loss"], res["loss_without_coeff"], res["grad_norm"], res["lr"]]),
        grad_names=np.array(names), grad_norms=np.array([float(res["grads"][k].norm()) for k in names]),
        grad_attention_v=res["grads"]["attention/v"].numpy(), grad_mel_proj_bias=res["grads"]["mel_proj/bias"].numpy(),
        grad_embedding=res["grads"]["embedding"].numpy(),
        param_after_attention_v=res["params"]["attention/v"].numpy(),
        bn_after_enc_p1_mean=res["params"]["enc_cbhg/proj_1/moving_mean"].numpy())
    # eval-mode (moving statistics) teacher-forced forward is not reachable through the reference API (targets => training);
    # the inference golden is the free-running forward
    inf = O.forward(named, hp, b["inputs"], b["input_lengths"], 1, None, speaker_mode="none", max_iters=6)
    np.savez_compressed(os.path.join(OUT, "tacotron_infer_small.npz"), mel_outputs=inf["mel_outputs"].numpy(),
                        linear_outputs=inf["linear_outputs"].numpy(), alignments=inf["alignments"].numpy())
    rng = np.random.RandomState(7)
    T = 12
    spec = rng.rand(T, 1025).astype(np.float32)
    phase = rng.rand(T, 1025).astype(np.float32)
    wav = G.inv_spectrogram(spec, phase, n_iters=4)
    np.savez_compressed(os.path.join(OUT, "griffin_lim_small.npz"), spec=spec, phase=phase, wav=wav, n_iters=4)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()

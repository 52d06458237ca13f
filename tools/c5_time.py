"""Batched free-running inference (C5 shape: deepvoice, 4 speakers, text_len 200, 200 decoder steps) at a given number of rows:
device time of Engine.forward, median of 5.  usage: python tools/c5_time.py <rows> [precision];  TACO_ATT_FREE=0 forces the
general free-running kernel for rows <= 32."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb
from importlib import import_module

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
hp = tb.hparams.override(reduction_factor=5, model_type="deepvoice")
eng = Engine(hp, 4, precision=prec, randomize_bn_state=True, seed=4321)
g = torch.Generator().manual_seed(3)
Ti = 200
L = torch.randint(120, Ti + 1, (N,), generator=g, dtype=torch.int32); L[0] = Ti
tok = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
for n in range(N):
    tok[n, L[n] - 1] = 1; tok[n, L[n]:] = 0
spk = torch.randint(0, 4, (N,), generator=g, dtype=torch.int32)
for _ in range(2):
    out = eng.forward(tok, L, spk, decoder_steps=200)
torch.cuda.synchronize()
ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.forward(tok, L, spk, decoder_steps=200); e1.record()
    torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
ms.sort()
print("C5-shape inference rows=%d %s TACO_ATT_FREE=%s: median %.3f ms per batch = %.0f utterances/s; finite %s"
      % (N, prec, os.environ.get("TACO_ATT_FREE", "1"), ms[2], N / ms[2] * 1e3, bool(torch.isfinite(out["linear_outputs"]).all())))

"""C4 forward timing (batch 1, 128 tokens, 200 free-running decoder steps -> 1000 frames): CUDA events around Engine.forward,
median of `reps`.  TACO_ATT_FREE=0 selects the general free-running kernel (attention.cu) instead of the resident-weight
one (att_free.cu); run once with each to compare.  Also prints max|diff| of the outputs between two identical calls."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb
from importlib import import_module
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
hp = tb.hparams.override(reduction_factor=5)
eng = Engine(hp, 1, precision=prec, randomize_bn_state=True, seed=4321)
tok, L, _ = bench.synth_inputs()
for _ in range(2):
    out = eng.forward(tok, L, decoder_steps=200)
torch.cuda.synchronize()
ref = {k: out[k].clone() for k in ("mel_outputs", "linear_outputs", "alignments")}
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.forward(tok, L, decoder_steps=200); e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms.sort()
rep = max(float((out[k] - ref[k]).abs().max()) for k in ref)
print("C4 forward %s TACO_ATT_FREE=%s: median %.3f ms (min %.3f, max %.3f) over %d calls; launches/call %d; repeatability max|diff| %.2e; finite %s"
      % (prec, os.environ.get("TACO_ATT_FREE", "1"), ms[len(ms) // 2], ms[0], ms[-1], reps, eng.launch_count() // (reps + 2),
         rep, bool(torch.isfinite(out["linear_outputs"]).all())))
torch.save({k: v.cpu() for k, v in ref.items()}, os.path.join(ROOT, "gpurun_out", "c4_out_%s_free%s.pt" % (prec, os.environ.get("TACO_ATT_FREE", "1"))))

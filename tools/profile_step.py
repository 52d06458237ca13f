"""One C2 training step between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb
from importlib import import_module
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
hp = tb.hparams.override(reduction_factor=5)
eng = Engine(hp, 1, precision=prec)
b = {k: v.to(eng.dev) for k, v in bench.synth_batch(0).items()}
for _ in range(2):
    eng.train_step(b)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.train_step(b)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", eng.scalars())

"""Warm per-call timing of every launch_gemm group in one C2 training step (CUDA events, no profiler).

usage (GPU box): python tools/gemm_spans.py [tf32]
"""
import ctypes as C, os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb
from importlib import import_module
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
hp = tb.hparams.override(reduction_factor=5)
eng = Engine(hp, 1, precision=prec)
b = {k: v.to(eng.dev) for k, v in bench.synth_batch(0).items()}
for _ in range(3):
    eng.train_step(b)
torch.cuda.synchronize()
lib = eng.lib
lib.taco_debug_profile_spans.restype = C.c_int
lib.taco_debug_profile_spans.argtypes = [C.c_char_p, C.c_int64]
lib.taco_profile(1, None, None)
eng.train_step(b)
ms = (C.c_double * 4)(); cnt = (C.c_int64 * 4)()
lib.taco_profile(0, ms, cnt)
buf = C.create_string_buffer(1 << 20)
lib.taco_debug_profile_spans(buf, len(buf))
rows = [l.split() for l in buf.value.decode().splitlines()]
print("class totals ms:", list(ms), list(cnt))
tot = 0.0
for i, r in enumerate(rows):
    cls, t = int(r[0]), float(r[1])
    if cls == 0:
        M, N, K, n = map(int, r[2:6])
        tot += t
        print(f"{i:4d} gemm {t*1e3:8.1f} us  first M={M:6d} N={N:5d} K={K:6d} problems={n:2d}  ({2.0*M*N*K*1e-9/max(t,1e-6):7.1f} TF/s if alone)")
    else:
        print(f"{i:4d} {'gru' if cls == 1 else 'att'}  {t*1e3:8.1f} us")
print("gemm total ms", tot)

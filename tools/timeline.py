"""Timeline of one warm C2 training step under the real two-stream backward schedule: stage markers, GEMM groups and
recurrence kernels with start offsets (CUDA events; run with TACO_PROF_OVERLAP=1 to keep the overlap on while profiling).

usage (GPU box): TACO_PROF_OVERLAP=1 python tools/timeline.py [tf32]
"""
import ctypes as C, os, sys
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record()
_h0 = time.perf_counter()
eng.train_step(b)
_host_ms = (time.perf_counter() - _h0) * 1e3
e1.record()
ms = (C.c_double * 4)(); cnt = (C.c_int64 * 4)()
lib.taco_profile(0, ms, cnt)
buf = C.create_string_buffer(1 << 21)
lib.taco_debug_profile_spans(buf, len(buf))
rows = []
for l in buf.value.decode().splitlines():
    f = l.split()
    rows.append(dict(cls=int(f[0]), ms=float(f[1]), tag=list(map(int, f[2:6])), t0=float(f[6]), side=int(f[7]), name=f[8]))
rows.sort(key=lambda r: r["t0"])
print("step %.3f ms (events around train_step, profiling events included); host issue time %.3f ms; class totals ms %s" % (e0.elapsed_time(e1), _host_ms, list(ms)))
names = {0: "gemm", 1: "gru", 2: "att", 9: "mark"}
for r in rows:
    lane = {0: "main", 1: "side", 2: "gru1", 3: "gru2"}[r["side"]]
    if r["cls"] == 9:
        print("%9.3f            %s  ---- %s" % (r["t0"], lane, r["name"]))
    elif r["cls"] == 0:
        M, N, K, n = r["tag"]
        print("%9.3f %9.3f  %s  gemm x%-2d first M=%d N=%d K=%d  (+%.3f)" % (r["t0"], r["t0"] + r["ms"], lane, n, M, N, K, r["ms"]))
    else:
        print("%9.3f %9.3f  %s  %s (+%.3f)" % (r["t0"], r["t0"] + r["ms"], lane, names[r["cls"]], r["ms"]))

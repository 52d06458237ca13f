import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import tacotron_b200 as tb, bench
from importlib import import_module
eng = tb.Engine(tb.hparams.override(reduction_factor=5), 1, precision="tf32")
b = {k: v.to(eng.dev) for k, v in bench.synth_batch(0).items()}
eng.forward(b["inputs"], b["input_lengths"], None, b["mel_targets"], b["linear_targets"], b["loss_coeff"])
torch.cuda.synchronize()
lib = eng.lib; lib.taco_debug_gru_prof.argtypes = [C.POINTER(C.c_longlong)]
buf = (C.c_longlong * 16)(); lib.taco_debug_gru_prof(buf)
v = list(buf); steps = max(v[10], 1)
names = ["top(expect,prefetch)", "gate mma+red store", "sync1", "act (reduce,sigmoid,stage)", "sync2", "st.async issue", "wait rh", "cand phase total (to wait h)"]
print("last GRU fwd launch (post bi-GRU, %d steps), cycles per step:" % steps)
for n, c in zip(names, v[:8]): print("  %-32s %8.1f" % (n, c / steps))
print("  total %.1f cycles/step" % (sum(v[:8]) / steps))

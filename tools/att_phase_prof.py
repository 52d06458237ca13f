import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import tacotron_b200 as tb, bench
eng = tb.Engine(tb.hparams.override(reduction_factor=5), 1, precision="tf32")
b = {k: v.to(eng.dev) for k, v in bench.synth_batch(0).items()}
eng.forward(b["inputs"], b["input_lengths"], None, b["mel_targets"], b["linear_targets"], b["loss_coeff"]); eng.backward()
torch.cuda.synchronize()
lib = eng.lib; lib.taco_debug_att_prof.argtypes = [C.POINTER(C.c_longlong)]
buf = (C.c_longlong * 40)(); lib.taco_debug_att_prof(buf)
v = list(buf); steps = 160 // max(1, int(os.environ.get("TACO_DEC_CHUNKS", "4")))
fn = ["P1 z1", "P2 z", "P3 gates", "P4 cand", "P5 q", "P6 scores", "P7 align+ctx"]
bn = ["Bp1 dctx", "Bp2 da", "Bp3+4 scan,gq", "Bp5 dcp", "Bp6 dg", "Bp7 dzp", "Bp8 dz1p"]
for title, names, off in (("forward", fn, 0), ("backward", bn, 20)):
    print(title, "cycles per step (compute until push | wait):")
    tot = 0
    for i, n in enumerate(names):
        c, w = v[off + 2 * i] / steps, v[off + 2 * i + 1] / steps; tot += c + w
        print("  %-16s %8.1f | %8.1f" % (n, c, w))
    print("  tail %8.1f   total %.1f" % (v[off + 14] / steps, tot + v[off + 14] / steps))
    print("  last launch: set-up loads %d cycles, first cluster barrier %d cycles, %.1f us inside the kernel (CTA 0)" % (v[off + 15], v[off + 16], v[off + 17] / 1e3))

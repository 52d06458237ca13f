"""Where does the end-to-end loop lose time against the device-resident loop?  Variants of bench.py's e2e loop, ms/step:
  a resident batch, no read-back   b resident + async scalars   c staged H2D + async scalars (= e2e)   d staged H2D, no read-back
usage (GPU box): python tools/e2e_diag.py [steps]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb
import bench
from importlib import import_module
Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
hp = tb.hparams.override(reduction_factor=5, batch_size=32)
eng = Engine(hp, 1, precision="tf32", seed=4321)
host = bench.synth_batch(0)
pinned = {k: v.pin_memory() for k, v in host.items()}
dev = {k: v.to(eng.dev) for k, v in host.items()}
bufs = [{k: torch.empty_like(v, device=eng.dev) for k, v in host.items()} for _ in range(2)]
copy = torch.cuda.Stream(device=eng.dev)
for _ in range(3):
    eng.train_step(dev)
torch.cuda.synchronize()

def run(stage_copies, readback, n, when="start"):
    free = [None, None]
    def stage(i):
        with torch.cuda.stream(copy):
            if free[i % 2] is not None:
                copy.wait_event(free[i % 2])
            for k in pinned:
                bufs[i % 2][k].copy_(pinned[k], non_blocking=True)
            e = torch.cuda.Event(); e.record(copy)
        return e
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ready = stage(0) if stage_copies else None
    pending = None
    for i in range(n):
        if stage_copies:
            torch.cuda.current_stream().wait_event(ready)
            b = bufs[i % 2]
            if when == "start":
                nxt = stage(i + 1) if i + 1 < n else None
                eng.train_step(b)
            else:
                eng.forward(b["inputs"], b["input_lengths"], None, b["mel_targets"], b["linear_targets"], b["loss_coeff"])
                if when == "after_fwd":
                    gate = torch.cuda.Event(); gate.record(torch.cuda.current_stream()); copy.wait_event(gate)
                    nxt = stage(i + 1) if i + 1 < n else None
                eng.backward()
                if when == "after_bwd":
                    gate = torch.cuda.Event(); gate.record(torch.cuda.current_stream()); copy.wait_event(gate)
                    nxt = stage(i + 1) if i + 1 < n else None
                eng.optimizer_step(True)
            free[i % 2] = torch.cuda.Event(); free[i % 2].record(torch.cuda.current_stream())
            ready = nxt
        else:
            eng.train_step(dev)
        if readback:
            cur = eng.scalars_async()
            if pending is not None:
                pending.get()
            pending = cur
    if pending is not None:
        pending.get()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n

for name, sc, rb, when in (("a resident, no read-back", False, False, "start"), ("c staged at step start (e2e)", True, True, "start"),
                           ("e staged after forward", True, True, "after_fwd"), ("f staged after backward", True, True, "after_bwd"),
                           ("a again", False, False, "start")):
    run(sc, rb, 3, when)
    print("%-32s %.3f ms/step" % (name, run(sc, rb, steps, when)))

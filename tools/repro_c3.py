import os, sys, torch
sys.path.insert(0, os.getcwd())
import tacotron_b200 as tb, bench
from importlib import import_module
Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
ranks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(8))
model_type = sys.argv[2] if len(sys.argv) > 2 else "deepvoice"
for r in ranks:
    hp = tb.hparams.override(reduction_factor=5, batch_size=32, **({"model_type": "deepvoice"} if model_type == "deepvoice" else {}))
    S = 3 if model_type == "deepvoice" else 1
    eng = Engine(hp, S, precision="bf16", device=0, seed=4321)
    b = bench.synth_batch(r)
    g = torch.Generator().manual_seed(99 + r)
    if S > 1: b["speaker_id"] = torch.randint(0, 3, (32,), generator=g, dtype=torch.int32)
    b["linear_targets"] = b["linear_targets"].to(torch.bfloat16)
    dev = {k: v.to(eng.dev) for k, v in b.items()}
    for i in range(8):
        eng.train_step(dev)
        torch.cuda.synchronize()
    print("rank-data", r, model_type, "ok loss", eng.scalars()["loss"], "lengths min", int(b["input_lengths"].min()), flush=True)
    eng.close()

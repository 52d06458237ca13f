#!/usr/bin/env python
"""bench.py — mel-frames/sec of one Tacotron training step (BASELINE.json metric, config C2).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference arm: CPU oracle restatement on host cores

A "step" = forward + L1 losses + backward (BPTT) + gradient all-reduce (N>1) + global-norm clip + Adam + BN update on
one synthetic batch: batch=32/GPU, text_len=128, mel_len=800, 80 mel bins, 1025 linear bins, r=5.
`value` times K steps with inputs resident in HBM; `e2e` repeats the run through the public Engine API with the
step's inputs copied from pinned host memory and the loss read back every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mel-frames/sec (train, batch=32/GPU)"
UNIT = "mel-frames/s"
CFG = dict(N=32, T_in=128, T_out=800, num_mels=80, num_freq=1025, r=5)
WORKLOAD = "C2: train step, batch=32/GPU, text_len=128, mel_len=800, 80-bin mel, 1025-bin linear, r=5, single-speaker"
# algorithmic HBM bytes of one training step per GPU (targets + params fwd/bwd + grads + clip/Adam)
ALGO_BYTES_PER_STEP = 486.6e6
# algorithmic FLOPs of one training step per GPU (3 x the 261.6 GFLOP forward; 30.66 MFLOP per mel frame)
ALGO_FLOPS_PER_STEP = 784.8e9
DTYPE_LABEL = {"bf16": "bf16", "tf32": "tf32+bf16-recurrent", "fp32": "fp32"}
DTYPE_NOTE = {
    "bf16": "bf16 operands (activations, gradients, weights) with fp32 accumulation in the CBHG / linear contractions (tcgen05 kind::f16), "
            "tf32 in the small decoder contractions, bf16 recurrent weights with fp32 state, MUFU tanh/sigmoid; parameters, Adam and batch-norm statistics fp32",
    "tf32": "fp32 operands rounded to tf32 inside the tensor core (tcgen05 kind::tf32), bf16 recurrent weights with fp32 state, MUFU tanh/sigmoid",
    "fp32": "exact-parity mode: fp32 FMA everywhere",
}


def gemm_traffic(precision):
    """DRAM bytes per GEMM launch from the committed ncu capture of this build (profiles/r2_gemm_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")) as f:
            t = json.load(f).get(precision)
        return (t["dram_bytes_per_launch"], t["source"]) if t else (None, "no ncu capture of this precision mode is committed")
    except (OSError, ValueError, KeyError):
        return None, "profiles/r2_gemm_traffic.json missing"


def synth_batch(rank: int, N=CFG["N"], Ti=CFG["T_in"], To=CFG["T_out"]):
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    L = torch.randint(96, Ti + 1, (N,), generator=g, dtype=torch.int32)
    L[0] = Ti
    inp = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
    for n in range(N):
        inp[n, L[n] - 1] = 1
        inp[n, L[n]:] = 0
    mel = torch.rand(N, To, CFG["num_mels"], generator=g)
    lin = torch.rand(N, To, CFG["num_freq"], generator=g)
    return dict(inputs=inp, input_lengths=L, mel_targets=mel, linear_targets=lin, loss_coeff=torch.ones(N))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), "fallback"


def run_ours(args):
    import torch
    import torch.distributed as dist
    import tacotron_b200 as tb
    from importlib import import_module
    Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner (and NCCL_DEBUG output) on the C-level stdout when the communicator is created; the
        # contract is ONE JSON line on stdout, so file descriptor 1 points at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hp = tb.hparams.override(reduction_factor=CFG["r"], batch_size=CFG["N"])
    eng = Engine(hp, 1, precision=args.precision, device=local, seed=4321)
    host = synth_batch(rank)
    if args.targets == "bf16":
        # linear targets travel (and stay) as bf16: 105 -> 52 MB of host->device copies per step; the loss is taken against them
        host["linear_targets"] = host["linear_targets"].to(torch.bfloat16)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    dev = {k: v.to(eng.dev) for k, v in host.items()}

    # the step's one collective: gradient all-reduce in two buckets, the larger one beside the encoder's backward pass (dist.py)
    OverlappedAllReduce = import_module("multi-speaker-tacotron-tensorflow_b200.dist").OverlappedAllReduce
    allreduce = OverlappedAllReduce(eng)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=eng.dev)   # > 126 MB L2

    # ---- device-resident timing ----
    # (the clock sampler is started before the warm-up so that nvidia-smi is already streaming when the short timed region
    #  begins; only rows stamped inside the timed region are used)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        eng.train_step(dev, allreduce=allreduce)
    barrier()
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    barrier()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        ev[i][0].record()
        eng.train_step(dev, allreduce=allreduce)
        ev[i][1].record()
    barrier()
    t_wall1 = time.time()
    launches = eng.launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=eng.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    sc = eng.scalars()

    # ---- roofline leg: per-class device time of the dominant kernels, CUDA events on the launching stream (outside the timed region) ----
    import ctypes as C
    prof_ms = (C.c_double * 4)(); prof_n = (C.c_int64 * 4)()
    PROF_STEPS = 3
    eng.lib.taco_profile(1, None, None)
    for _ in range(PROF_STEPS):
        eng.train_step(dev, allreduce=allreduce)
    eng.lib.taco_profile(0, prof_ms, prof_n)
    gemm_ms = prof_ms[0] / PROF_STEPS
    gemm_flops = prof_ms[3] * 1e9 / PROF_STEPS          # sum of 2*M*N*K over the step's GEMM problems, counted at launch
    barrier()

    # ---- end to end through the public API: pinned host inputs -> device every step, loss read back every step ----
    stream_copy = torch.cuda.Stream(device=eng.dev)
    bufs = [{k: torch.empty_like(v, device=eng.dev) for k, v in host.items()} for _ in range(2)]
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    free_ev = [None, None]       # compute-stream event after the last step that read bufs[k]

    def stage(i):
        with torch.cuda.stream(stream_copy):
            if free_ev[i % 2] is not None:
                stream_copy.wait_event(free_ev[i % 2])      # do not overwrite a batch a running step still reads
            for k in pinned:
                bufs[i % 2][k].copy_(pinned[k], non_blocking=True)
            e = torch.cuda.Event(); e.record(stream_copy)
        return e

    def e2e_loop(n):
        # the loop of train.py:217-226 with one step of run-ahead: step i's scalars are copied to pinned host memory behind
        # step i and read on the host once step i+1 has been enqueued (every step's loss is read inside the timed region).
        # The copy of batch i+1 is issued behind step i's FORWARD pass (an event gates the copy stream): H2D traffic beside
        # the forward pass slows it by 0.7 ms, beside the backward pass by 0.08 ms (tools/e2e_diag.py).
        free_ev[0] = free_ev[1] = None
        ready = stage(0)
        losses, pending = [], None
        for i in range(n):
            torch.cuda.current_stream().wait_event(ready)
            box = [None]

            def prefetch(i=i):
                if i + 1 < n:
                    gate = torch.cuda.Event(); gate.record(torch.cuda.current_stream())
                    stream_copy.wait_event(gate)
                    box[0] = stage(i + 1)
            eng.train_step(bufs[i % 2], allreduce=allreduce, after_forward=prefetch)
            nxt = box[0]
            cur = eng.scalars_async()
            free_ev[i % 2] = torch.cuda.Event(); free_ev[i % 2].record(torch.cuda.current_stream())
            if pending is not None:
                losses.append(pending.get()["loss"])
            pending = cur
            ready = nxt
        losses.append(pending.get()["loss"])
        return losses

    e2e_loop(max(1, min(2, args.warmup)))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=eng.dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    extra = {}
    if args.legs:
        eng.close()
        del dev, bufs, flush
        torch.cuda.empty_cache()
        if world == 1:
            extra["c3"] = leg_c3(tb, Engine, args.precision, local, rank, world, OverlappedAllReduce, barrier, args.targets)
            extra["c5"] = leg_c5(tb, Engine, args.precision, local, rank, world, barrier)
        else:
            # Multi-GPU: the two extra legs run as a SEPARATE job after this one has given its process group up, so that
            # whatever happens in them (an 8-GPU run of the C3 leg died once with a launch failure on one rank that no
            # single- or two-GPU run reproduces) cannot take the headline measurement with it.
            barrier()
            dist.destroy_process_group()
            if rank != 0:
                return
            extra.update(run_legs_job(args, world))

    frames = world * CFG["N"] * CFG["T_out"] * args.steps
    if rank == 0:
        peaks, which = measured_peaks()
        traffic, traffic_src = gemm_traffic(args.precision)
        n_gemm = int(prof_n[3] // PROF_STEPS)
        ms_step = total_ms / args.steps
        ach = ALGO_BYTES_PER_STEP / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": frames / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_LABEL[args.precision], "dtype_note": DTYPE_NOTE[args.precision], "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "global_batch": world * CFG["N"], "parallelism": "dp%d" % world, "l2": "flushed between timed steps (160 MB write)",
                       "linear_targets": args.targets + (" (stored and copied as bf16; the loss is taken against the bf16 values)" if args.targets == "bf16" else ""),
                       "collective": "one gradient all-reduce per step (NCCL), in two buckets: decoder/post-net/linear gradients beside the encoder's backward pass",
                       "timing": "CUDA events per step on the compute stream, max over ranks"},
            "e2e": {"value": frames / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 96,
                    "note": "Engine.train_step with inputs staged from pinned host memory (double-buffered copy stream, the next batch's copy issued behind the forward pass) + every step's loss read back to the host (pinned, one step of run-ahead)"},
            "gpu_launches": launches,
            "clocks": clocks,
            # dominant kernel class by device time: the tcgen05 GEMMs (all GEMM-shaped work of the step)
            "roofline": {"bound": "tensor",
                         "kernel": {"bf16": "gemm_bf16_kernel (tcgen05.mma kind::f16, bf16 operands, TMA-fed, persistent) + gemm_tc_kernel (kind::tf32) for the small decoder problems",
                                    "tf32": "gemm_tc_kernel (tcgen05.mma kind::tf32, TMA-fed)", "fp32": "gemm_simt_kernel (fp32 FMA)"}[args.precision],
                         "achieved": gemm_flops / (gemm_ms * 1e-3) / 1e12, "peak": peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
                         "unit": "TFLOP/s", "frac": gemm_flops / (gemm_ms * 1e-3) / 1e12 / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": which + " (dense bf16 cuBLAS, sustained)",
                         "launches_per_step": n_gemm, "ms_per_step": gemm_ms,
                         "algorithmic_flops_per_step": gemm_flops, "algorithmic_flops_per_launch": gemm_flops / max(1, n_gemm),
                         "avg_launch_us": 1e3 * gemm_ms / max(1, n_gemm),
                         "note": "one denominator: launches_per_step = GEMM problems of one step = GEMM kernel launches (each problem is one launch of a tensor-core kernel; "
                                 "the few sub-tile problems are batched into grouped fp32 launches); algorithmic FLOPs = sum of 2MNK over them, counted where they are launched; "
                                 "time = CUDA events around every GEMM call on its launching stream, summed, in a serialised window (the multi-stream schedule is off while profiling)"},
            "recurrence_ms_per_step": {"gru_fwd_bwd": prof_ms[1] / PROF_STEPS, "attention_fwd_bwd": prof_ms[2] / PROF_STEPS,
                                       "note": "serial chains: 2 656 dependent recurrence steps per training step (latency bound)"},
            "roofline_step_hbm": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                                  "note": "whole-step algorithmic bytes (486.6 MB, SURVEY.md 8d) / step time"},
            "loss": sc["loss"],
        }
        line.update(extra)
        if args.synth and world == 1:
            eng.close()
            line["synth_rtf"] = synth_rtf_ours(hp, local, args.precision)
        if args.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(sample_steps=3, warmup=1)
            if args.synth and world == 1:
                line["cpu_baseline"]["synth_rtf"] = synth_rtf_cpu()
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()


def run_legs_job(args, world):
    """Rank 0 of a multi-GPU run: the C3 / C5 legs as a child job (`bench.py --legs-only` under its own torch.distributed.run on
    the same GPUs), bounded by a timeout; returns {"c3": ..., "c5": ...} or an error record in their place."""
    env = {k: v for k, v in os.environ.items()
           if not (k.startswith("TORCHELASTIC") or k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE",
                                                         "ROLE_RANK", "ROLE_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS"))}
    port = int(os.environ.get("MASTER_PORT", "29500")) + 17
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.abspath(__file__), "--legs-only", "--gpus", str(world), "--precision", args.precision, "--targets", args.targets]
    try:
        p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        for line in reversed(p.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        lines = [l.strip() for l in (p.stderr or "").splitlines() if "rror" in l]
        err = "legs job exited with %d: %s" % (p.returncode, lines[-1][:300] if lines else "no error line on stderr")
    except subprocess.TimeoutExpired:
        err = "legs job timed out after 300 s"
    except Exception as e:      # noqa: BLE001 - the legs are optional extras of the line
        err = "legs job failed: %r" % (e,)
    return {"c3": {"error": err}, "c5": {"error": err}}


def run_legs_only(args):
    """`bench.py --legs-only` (internal): the C3 and C5 legs alone, one JSON object on rank 0's stdout."""
    import torch
    import torch.distributed as dist
    import tacotron_b200 as tb
    from importlib import import_module
    Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
    OverlappedAllReduce = import_module("multi-speaker-tacotron-tensorflow_b200.dist").OverlappedAllReduce
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)                       # NCCL's banner goes to stderr (see run_ours)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    out = {"c3": leg_c3(tb, Engine, args.precision, local, rank, world, OverlappedAllReduce, barrier, args.targets),
           "c5": leg_c5(tb, Engine, args.precision, local, rank, world, barrier)}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def leg_c3(tb, Engine, precision, local, rank, world, make_allreduce, barrier, targets, steps=5, warmup=3):
    """BASELINE.json configs[2]: the same training step with model_type=deepvoice, 3 speakers (batch 32 per GPU) - device-timed."""
    import torch
    import torch.distributed as dist
    hp = tb.hparams.override(reduction_factor=CFG["r"], batch_size=CFG["N"], model_type="deepvoice")
    eng = Engine(hp, 3, precision=precision, device=local, seed=4321)
    b = synth_batch(rank)
    g = torch.Generator().manual_seed(99 + rank)
    b["speaker_id"] = torch.randint(0, 3, (CFG["N"],), generator=g, dtype=torch.int32)
    if targets == "bf16":
        b["linear_targets"] = b["linear_targets"].to(torch.bfloat16)
    dev = {k: v.to(eng.dev) for k, v in b.items()}
    allreduce = make_allreduce(eng)
    trace = os.environ.get("TACO_BENCH_TRACE") == "1"      # per-step synchronisation + progress lines on stderr (diagnosis only)
    for i in range(warmup):
        eng.train_step(dev, allreduce=allreduce)
        if trace:
            torch.cuda.synchronize()
            print("[c3 trace] rank %d warm-up step %d done, loss %.6f" % (rank, i, eng.scalars()["loss"]), file=sys.stderr, flush=True)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        ev[i][0].record(); eng.train_step(dev, allreduce=allreduce); ev[i][1].record()
    barrier()
    t = torch.tensor([sum(a.elapsed_time(c) for a, c in ev)], dtype=torch.float64, device=eng.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    loss = eng.scalars()["loss"]
    eng.close()
    ms = float(t.item()) / steps
    return {"workload": "C3: train step, model_type=deepvoice, 3 speakers, batch=32/GPU, text_len=128, mel_len=800, r=5", "ms_per_step": ms,
            "value": world * CFG["N"] * CFG["T_out"] / (ms * 1e-3), "unit": UNIT, "steps": steps, "warmup": warmup, "loss": loss,
            "note": "device-timed (CUDA events, max over ranks), inputs resident, L2 not flushed"}


def leg_c5(tb, Engine, precision, local, rank, world, barrier, reps=5):
    """BASELINE.json configs[4]: batched inference, 64 utterances, 4 speakers, text_len=200, 200 free-running decoder steps ->
    1000 frames + post-net + linear projection.  The 64 rows are sharded over the ranks (replicas only, no collective)."""
    import torch
    import torch.distributed as dist
    from importlib import import_module
    shard_rows = import_module("multi-speaker-tacotron-tensorflow_b200.dist").shard_rows
    hp = tb.hparams.override(reduction_factor=CFG["r"], model_type="deepvoice")
    S, N, Ti, steps = 4, 64, 200, 200
    eng = Engine(hp, S, precision=precision, device=local, seed=4321, randomize_bn_state=True)
    g = torch.Generator().manual_seed(77)
    L = torch.randint(120, Ti + 1, (N,), generator=g, dtype=torch.int32); L[3] = Ti
    tok = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
    for n in range(N):
        tok[n, L[n] - 1] = 1; tok[n, L[n]:] = 0
    spk = torch.randint(0, S, (N,), generator=g, dtype=torch.int32)
    r0, r1 = shard_rows(N, rank, world)
    tok, L, spk = tok[r0:r1].to(eng.dev), L[r0:r1].to(eng.dev), spk[r0:r1].to(eng.dev)
    for _ in range(2):
        eng.forward(tok, L, spk, decoder_steps=steps)
    barrier()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = eng.forward(tok, L, spk, decoder_steps=steps); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64, device=eng.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    finite = bool(torch.isfinite(out["linear_outputs"]).all())
    eng.close()
    ms = float(t.item())
    return {"workload": "C5: batched inference, 64 utterances (sharded over the ranks), 4 speakers, text_len=200, 200 free-running steps -> 1000 frames, post-net, linear projection",
            "ms_per_batch": ms, "rows_per_s": N / (ms * 1e-3), "frames_per_s": N * steps * CFG["r"] / (ms * 1e-3), "rows_per_rank": r1 - r0, "finite": finite,
            "note": "device-timed median of %d calls, max over ranks; tokens resident" % reps}


SYNTH = dict(N=1, T_in=128, steps=200, gl_iters=60)      # config C4: 200 free-running steps -> 1000 frames


def synth_inputs():
    import torch
    g = torch.Generator().manual_seed(1234)
    tok = torch.randint(2, 80, (SYNTH["N"], SYNTH["T_in"]), generator=g, dtype=torch.int32)
    tok[:, -1] = 1
    L = torch.full((SYNTH["N"],), SYNTH["T_in"], dtype=torch.int32)
    phase = torch.rand(SYNTH["steps"] * CFG["r"], CFG["num_freq"], generator=g)
    return tok, L, phase


def synth_rtf_ours(hp, device, precision, reps=5):
    """Second half of BASELINE.json's metric: real-time factor of synthesis (C4): tokens -> free-running decoder (200 steps,
    1000 mel frames) -> post-net -> linear spectrogram -> 60-iteration Griffin-Lim -> 299 700 samples (12.49 s at 24 kHz).
    Timed through the public API with the tokens on the host and the waveform copied back (synthesizer.py:166-167,264)."""
    import torch
    from importlib import import_module
    Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
    GriffinLim = import_module("multi-speaker-tacotron-tensorflow_b200.audio").GriffinLim
    eng = Engine(hp, 1, precision=precision, device=device, seed=4321, randomize_bn_state=True)
    gl = GriffinLim(hp, max_frames=SYNTH["steps"] * CFG["r"], device=device)
    tok, L, phase = synth_inputs()
    tok, L = tok.pin_memory(), L.pin_memory()
    phase = phase.to(eng.dev)
    n0 = eng.launch_count()

    def once():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t0 = time.perf_counter()
        e[0].record()
        out = eng.forward(tok, L, decoder_steps=SYNTH["steps"])
        e[1].record()
        wav = gl.inv_spectrogram(out["linear_outputs"][0], phase, n_iters=SYNTH["gl_iters"])
        e[2].record()
        host = wav.cpu()
        wall = time.perf_counter() - t0
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), wall, host

    for _ in range(2):
        once()
    launches_per_call = (eng.launch_count() - n0) // 2
    rows = [once() for _ in range(reps)]
    host = rows[-1][3]
    audio_s = host.numel() / float(hp.sample_rate)
    med = lambda xs: sorted(xs)[len(xs) // 2]
    fwd, glt, wall = med([r[0] for r in rows]), med([r[1] for r in rows]), med([r[2] for r in rows])
    eng.close(); gl.close()
    return {"workload": "C4: batch=1, text_len=128, 200 free-running decoder steps -> 1000 frames, post-net, 60-iter Griffin-Lim",
            "rtf": wall / audio_s, "audio_s": audio_s, "samples": int(host.numel()), "ms_total_e2e": wall * 1e3,
            "ms_forward": fwd, "ms_griffin_lim": glt, "launches_per_call": int(launches_per_call), "reps": reps,
            "finite": bool(torch.isfinite(host).all()), "higher_is_better": False,
            "note": "rtf = wall time (host tokens in, host waveform out) / audio duration; lower is better"}


def synth_rtf_cpu(threads: int = 0):
    """The same C4 synthesis on the host cores with the oracle restatement (torch-CPU forward + numpy Griffin-Lim)."""
    import torch
    import tacotron_b200 as tb
    from oracle import tacotron_oracle as O
    from oracle import griffin_lim_oracle as G
    n = threads or os.cpu_count() or 1
    torch.set_num_threads(n)
    hp = tb.hparams.override(reduction_factor=CFG["r"])
    P = tb.params.init_params(hp, 1, seed=4321, randomize_bn_state=True)
    tok, L, phase = synth_inputs()
    t0 = time.perf_counter()
    with torch.no_grad():
        out = O.forward(P, hp, tok, L, 1, None, max_iters=SYNTH["steps"], speaker_mode="none")
    t1 = time.perf_counter()
    wav = G.inv_spectrogram(out["linear_outputs"][0].numpy(), phase.numpy(), n_iters=SYNTH["gl_iters"])
    t2 = time.perf_counter()
    audio_s = len(wav) / float(hp.sample_rate)
    return {"rtf": (t2 - t0) / audio_s, "s_forward": t1 - t0, "s_griffin_lim": t2 - t1, "audio_s": audio_s, "cores": n, "kind": "port",
            "sample": "one full C4 synthesis (oracle forward on torch-CPU + numpy/pocketfft Griffin-Lim, 60 iterations)"}


def cpu_baseline(sample_steps: int = 1, threads: int = 0, warmup: int = 0):
    """The CPU oracle restatement (torch-CPU, all host cores) timed on full C2 training steps: `warmup` untimed steps, then
    `sample_steps` timed ones; the MEDIAN step time is reported (the first step pays allocator / thread-pool start-up)."""
    import torch
    import tacotron_b200 as tb
    from oracle import tacotron_oracle as O
    n = threads or os.cpu_count() or 1
    torch.set_num_threads(n)
    hp = tb.hparams.override(reduction_factor=CFG["r"], batch_size=CFG["N"])
    P = tb.params.init_params(hp, 1, seed=4321)
    names = [k for k in P if not (k.endswith("moving_mean") or k.endswith("moving_var"))]
    m = {k: torch.zeros_like(P[k]) for k in names}
    v = {k: torch.zeros_like(P[k]) for k in names}
    b = synth_batch(0)
    times = []
    for i in range(warmup + sample_steps):
        t0 = time.perf_counter()
        res = O.train_step(P, m, v, hp, b, i, True, 1, "none")
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        P, m, v = res["params"], res["m"], res["v"]
    sec = sorted(times)[len(times) // 2]
    return {"value": CFG["N"] * CFG["T_out"] / sec, "unit": UNIT, "cores": n, "kind": "port", "steps": sample_steps, "warmup": warmup,
            "step_seconds": [round(t, 3) for t in times],
            "sample": "%d warm-up + %d timed full C2 training steps (fwd+bwd+clip+Adam, fp32) of the TF-semantics oracle on torch-CPU, median %.2f s/step"
                      % (warmup, sample_steps, sec)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use all host cores, so undo that BEFORE torch is imported
    ncpu = str(os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = ncpu
    os.environ["MKL_NUM_THREADS"] = ncpu
    # one warm-up step and >= 3 timed ones at every N (each step is a bounded sample: one full C2 batch, ~2.5 s on 16 cores)
    steps = max(3, min(args.steps, 5))
    base = cpu_baseline(sample_steps=steps, warmup=1)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": 1e3 * CFG["N"] * CFG["T_out"] / base["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": CFG["N"], "parallelism": "cpu",
                       "note": "the reference's train step restated on torch-CPU (oracle/tacotron_oracle.py, pinned to the reference's own "
                               "model code by tests/golden/ref_*.npz); TensorFlow 1.x itself cannot be installed here"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("TACO_PRECISION", "bf16"), choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--targets", default=None, choices=["fp32", "bf16"], help="storage type of the linear-spectrogram targets (default: bf16 in bf16 mode)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-synth", dest="synth", action="store_false", help="skip the C4 synthesis real-time-factor leg (N=1 only)")
    ap.add_argument("--no-legs", dest="legs", action="store_false", help="skip the C3 (deepvoice training) and C5 (batched inference) legs")
    ap.add_argument("--legs-only", action="store_true", help="(internal) run only the C3 / C5 legs and print them as one JSON object")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.targets is None:
        args.targets = "bf16" if args.precision == "bf16" else "fp32"
    if args.legs_only:
        run_legs_only(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

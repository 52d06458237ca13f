"""Training command line — mirror of the reference's ``train.py`` (flags: train.py:281-297; loop: train.py:100-276).

Same flags, same run-directory layout (``<log_dir>/<datasets>_<time>/{params.json, train.log, model.ckpt-<step>.pt}``), same
log lines; the TF session loop ``sess.run([global_step, loss, optimize])`` (train.py:217-219) becomes
``model.train_step(batch)`` on the CUDA engine.  New: data-parallel training under ``torchrun`` (one rank per GPU, one NCCL
all-reduce of the flat gradient per step), ``--precision`` and ``--max_steps``.

Not mirrored (out of scope): Slack reporting, TensorBoard summaries, alignment plots, the jamo text round
trip check.  The periodic test step still runs the free-running model (rnn_decoder_test_mode, train.py:158-166) and can
write the predicted audio with the GPU Griffin-Lim (``--test_audio``).
"""
from __future__ import annotations

import argparse
import math
import os
import time
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist

from . import dist as dp
from .datasets import DataFeeder
from .hparams import hparams, hparams_debug_string, load_hparams, save_hparams
from .models import create_model, get_most_recent_checkpoint
from .tf_checkpoint import load_any


class ValueWindow:                                              # utils/__init__.py:15-37
    def __init__(self, window_size=100):
        self._window_size, self._values = window_size, []

    def append(self, x):
        self._values = self._values[-(self._window_size - 1):] + [x]

    @property
    def average(self):
        return sum(self._values) / max(1, len(self._values))


def str2bool(v):                                                # utils/__init__.py
    return str(v).lower() in ("true", "1", "yes", "y")


def get_time():
    return datetime.now().strftime("%Y-%m-%d_%H-%M-%S")


class _Log:
    def __init__(self, path=None, enabled=True):
        self._f = open(path, "a", encoding="utf-8") if (path and enabled) else None
        self._enabled = enabled

    def __call__(self, msg, slack=False):                       # utils/infolog.py: print + file (+ slack, dropped)
        if not self._enabled:
            return
        print(msg, flush=True)
        if self._f:
            self._f.write("[%s]  %s\n" % (datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f")[:-3], msg))
            self._f.flush()


def prepare_dirs(config, hp, rank=0):
    """utils/__init__.py:39-61: resolve model_dir, write or load params.json."""
    config.datasets = [os.path.basename(os.path.normpath(p)) for p in config.data_paths]
    if config.load_path:
        config.model_dir = config.load_path
        load_hparams(hp, config.model_dir)
    else:
        config.model_name = "{}_{}".format("+".join(config.datasets), get_time())
        config.model_dir = os.path.join(config.log_dir, config.model_name)
        if rank == 0:
            os.makedirs(config.model_dir, exist_ok=True)
        hp.num_speakers = len(config.datasets)
        if rank == 0:
            save_hparams(config.model_dir, hp)


def save_checkpoint(model, model_dir, step, keep=5):
    """tf.train.Saver(max_to_keep=5) -> ``model.ckpt-<step>.pt`` (train.py:175,242-244)."""
    path = os.path.join(model_dir, "model.ckpt-%d.pt" % step)
    tmp = path + ".tmp"
    torch.save(model.state_dict(), tmp)
    os.replace(tmp, path)
    olds = sorted((int(f.split("-")[1].split(".")[0]), f) for f in os.listdir(model_dir) if f.startswith("model.ckpt-") and f.endswith(".pt"))
    for _, f in olds[:-keep]:
        os.remove(os.path.join(model_dir, f))
    return path


def train(log_dir, config, hp=hparams):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    log = _Log(os.path.join(log_dir, "train.log"), enabled=(rank == 0))
    data_dirs = [os.path.join(p, "data") for p in config.data_paths]                       # train.py:104-105
    num_speakers = len(data_dirs)
    config.num_test = config.num_test_per_speaker * num_speakers
    if num_speakers > 1 and hp.model_type not in ["deepvoice", "simple"]:
        raise Exception("[!] Unkown model_type for multi-speaker: {}".format(hp.model_type))   # train.py:109-110
    checkpoint_path = os.path.join(log_dir, "model.ckpt")
    log(" [*] Checkpoint path: %s" % checkpoint_path)
    log(" [*] Loading training data from: %s" % data_dirs)
    log(" [*] Using model: %s" % config.model_dir)
    log(hparams_debug_string(hp))
    device = torch.device("cuda", local)
    quiet = (lambda *a, **k: None)
    train_feeder = DataFeeder(data_dirs, hp, config, 32, data_type="train", batch_size=hp.batch_size, rank=rank, world=world,
                              device=device, log=log if rank == 0 else quiet)              # train.py:126-129
    test_feeder = DataFeeder(data_dirs, hp, config, 8, data_type="test", batch_size=config.num_test, device=device,
                             log=quiet) if config.test_interval > 0 else None              # train.py:130-132

    is_randomly_initialized = config.initialize_path is None                               # train.py:135
    model = create_model(hp)
    model._precision, model._device = config.precision, local
    model._seed = config.random_seed
    # build the engine with a first (tiny) dummy call is avoided: the engine is created on the first initialize()
    start_step = 0
    restore = None
    if config.load_path:
        restore = get_most_recent_checkpoint(config.model_dir)
    elif config.initialize_path:
        restore = get_most_recent_checkpoint(config.initialize_path)

    restored_state = None
    if restore and config.load_path:
        # the reference starts its feeders from the restored global_step (train.py:207-210): it decides the initial
        # data-greedy / equal-ratio sampling phase, so a resumed run must not replay it
        restored_state = load_any(restore, hp, num_speakers)
        start_step = int(restored_state.get("global_step", 0))
    train_feeder.start_in_session(None, start_step)
    if test_feeder is not None:
        test_feeder.start_in_session(None, start_step)
    time_window, loss_window = ValueWindow(100), ValueWindow(100)
    allreduce = None        # data parallel: created with the engine (two-bucket all-reduce, dist.OverlappedAllReduce)
    step = 0
    try:
        batch = train_feeder.next_device_batch()
        first = True
        while True:
            start_time = time.time()
            ready = batch.pop("_ready", None)
            if ready is not None:
                torch.cuda.current_stream().wait_event(ready)       # the batch staged behind the previous step's forward pass
            model.initialize(batch["inputs"], batch["input_lengths"], num_speakers, batch.get("speaker_id"), batch["mel_targets"],
                             batch["linear_targets"], batch["loss_coeff"], is_randomly_initialized=is_randomly_initialized)
            if first and restore:
                sd = restored_state if restored_state is not None else load_any(restore, hp, num_speakers)    # our .pt state or a reference TensorFlow checkpoint
                model.load_state_dict(sd, reset_step=bool(config.initialize_path))         # train.py:189-205
                log(("Resuming from checkpoint: %s" if config.load_path else "Initialized from checkpoint: %s") % restore, slack=True)
                if config.initialize_path:
                    log("=" * 50); log(" [*] Global step is reset to {}".format(model.engine.global_step)); log("=" * 50)
                model.initialize(batch["inputs"], batch["input_lengths"], num_speakers, batch.get("speaker_id"), batch["mel_targets"],
                                 batch["linear_targets"], batch["loss_coeff"], is_randomly_initialized=is_randomly_initialized)
            elif first:
                log("Starting new training run", slack=True)
            first = False
            fwd_done = torch.cuda.Event(); fwd_done.record(torch.cuda.current_stream())
            nxt = train_feeder.next_device_batch(after=fwd_done, defer_wait=True)   # the next batch's DMA runs beside the backward pass
            model.add_loss()
            if world > 1 and allreduce is None:
                allreduce = dp.OverlappedAllReduce(model.engine)
            model.add_optimizer(allreduce=allreduce)
            step = model.engine.global_step                 # value AFTER the update, as sess.run([global_step, ...]) returns
            loss = model.loss_without_coeff
            time_window.append(time.time() - start_time)
            loss_window.append(loss)
            log("Step %-7d [%.03f sec/step, loss=%.05f, avg_loss=%.05f]" % (step, time_window.average, loss, loss_window.average),
                slack=(step % config.checkpoint_interval == 0))                             # train.py:224-226
            if loss > 100 or math.isnan(loss):
                log("Loss exploded to %.05f at step %d!" % (loss, step), slack=True)
                raise Exception("Loss Exploded")                                            # train.py:228-230
            if step % config.checkpoint_interval == 0:
                dp.average_bn_state_(model.engine.bn_state)
                if rank == 0:
                    log("Saving checkpoint to: %s-%d" % (checkpoint_path, step))
                    save_checkpoint(model, log_dir, step)
            if test_feeder is not None and step % config.test_interval == 0 and rank == 0:
                log("Saving audio and alignment...")
                tb = test_feeder.next_device_batch()
                out = model.engine.forward(tb["inputs"], tb["input_lengths"], tb.get("speaker_id"), tb["mel_targets"], tb["linear_targets"],
                                           tb["loss_coeff"], rnn_decoder_test_mode=True)    # test_model, train.py:158-166
                np.save(os.path.join(log_dir, "step-%d-test-align.npy" % step), out["alignments"][:1].cpu().numpy())
                if config.test_audio:
                    from .audio import inv_spectrogram
                    from scipy.io import wavfile
                    wav = inv_spectrogram(out["linear_outputs"][0].t().cpu().numpy(), hp)
                    wavfile.write(os.path.join(log_dir, "step-%d-test-audio.wav" % step), hp.sample_rate, wav.astype(np.float32))
                log("Test finished for step {}.".format(step))
            batch = nxt
            if config.max_steps and step >= config.max_steps:
                break
    finally:
        train_feeder.stop()
        if test_feeder is not None:
            test_feeder.stop()
    return step


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--log_dir", default="logs")
    parser.add_argument("--data_paths", default="datasets/kr_example")
    parser.add_argument("--load_path", default=None)
    parser.add_argument("--initialize_path", default=None)
    parser.add_argument("--num_test_per_speaker", type=int, default=2)
    parser.add_argument("--random_seed", type=int, default=123)
    parser.add_argument("--summary_interval", type=int, default=100)
    parser.add_argument("--test_interval", type=int, default=500)
    parser.add_argument("--checkpoint_interval", type=int, default=1000)
    parser.add_argument("--skip_path_filter", type=str2bool, default=False, help="Use only for debugging")
    parser.add_argument("--slack_url", help="(accepted for compatibility; Slack reporting is not implemented)")
    parser.add_argument("--git", action="store_true", help="(accepted for compatibility)")
    # additions of this implementation
    parser.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "bf16"])
    parser.add_argument("--max_steps", type=int, default=0, help="stop after this many steps (0 = run until interrupted)")
    parser.add_argument("--test_audio", type=str2bool, default=False, help="write Griffin-Lim audio of the periodic test step")
    parser.add_argument("--hparams", default="", help="comma separated name=value overrides")
    return parser


def main(argv=None, hp=None):
    config = build_parser().parse_args(argv)
    config.data_paths = config.data_paths.split(",")
    hp = hp if hp is not None else hparams
    if config.hparams:
        hp.parse(config.hparams)
    hp.num_speakers = len(config.data_paths)                                               # train.py:301
    rank = int(os.environ.get("RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # every rank must resolve the same run directory: rank 0 decides, the others read it from the store
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prepare_dirs(config, hp, rank)
    if dist.is_initialized():
        box = [config.model_dir]
        dist.broadcast_object_list(box, src=0)
        config.model_dir = box[0]
    torch.manual_seed(config.random_seed)                                                  # tf.set_random_seed, train.py:308
    if any("krbook" not in p for p in config.data_paths) and hp.sample_rate != 20000:
        print(" [!] sample_rate is {} (the reference warns unless it is 20000 for non-krbook data, train.py:310-313)".format(hp.sample_rate))
    return train(config.model_dir, config, hp)


if __name__ == "__main__":
    main()

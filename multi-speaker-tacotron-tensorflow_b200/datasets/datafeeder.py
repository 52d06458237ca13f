"""Batch assembly for training — mirror of the reference's ``datasets/datafeeder.py`` (reference lines cited inline).

What the reference does (datafeeder.py:76-328): a thread reads ``.npz`` examples ``{tokens i32[L], mel f32[T,80],
linear f32[T,1025], loss_coeff}`` (datasets/generate_data.py:156-172), draws ``batch_size * batches_per_group`` of them
per group, sorts the group by frame count, cuts it into batches, shuffles the batches, pads each batch (tokens with 0
to the longest; targets with 0 to ``round_up(max_len + 1, r)``) and pushes it through a TF FIFOQueue of depth 8.

What this module does differently (same batches, B200-first plumbing):
  * batches are assembled straight into **pinned** host buffers (one ring slot per in-flight batch), so the
    host->device copy of the 113 MB/step of targets is a single async DMA per tensor on a copy stream that overlaps the
    previous training step (``DataFeeder.next_device_batch``);
  * data parallel: one feeder per rank, each with its own example stream (seed offset by rank) and ``batch_size`` rows per
    GPU — the weak-scaling convention of SURVEY.md §8e; no rank ever sees another rank's rows;
  * no TensorFlow: placeholders/queue/coordinator become a bounded ``queue.Queue`` and a daemon thread.
"""
from __future__ import annotations

import os
import queue
import threading
import time
from collections import defaultdict
from glob import glob
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_pad = 0                                                        # datafeeder.py:19


def _round_up(x: int, multiple: int) -> int:                    # datafeeder.py:326-328
    rem = x % multiple
    return x if rem == 0 else x + multiple - rem


def write_example(path: str, tokens, mel, linear, loss_coeff: float = 1.0) -> None:
    """One training example in the reference's on-disk schema (datasets/generate_data.py:156-172)."""
    np.savez(path, tokens=np.asarray(tokens, np.int32), mel=np.asarray(mel, np.float32),
             linear=np.asarray(linear, np.float32), loss_coeff=np.float32(loss_coeff))


def _frame_info(path: str) -> Tuple[str, int, int]:             # datafeeder.py:21-25
    with np.load(path) as data:
        return path, int(data["linear"].shape[0]), int(len(data["tokens"]))


def get_path_dict(data_dirs: Sequence[str], hparams, config, data_type: str, n_test: Optional[int] = None,
                  rng: Optional[np.random.RandomState] = None, log=print) -> Dict[str, List[str]]:
    """datafeeder.py:27-74: per data dir, the (shuffled, length-filtered) example paths; the last ``n_test`` are the test
    split.  ``glob`` order is made deterministic (sorted) before the shuffle so that ranks agree on the split."""
    rng = rng if rng is not None else np.random.RandomState(123)
    path_dict = {}
    for data_dir in data_dirs:
        paths = sorted(glob("{}/*.npz".format(data_dir)))
        if data_type == "train":
            rng.shuffle(paths)
        if not getattr(config, "skip_path_filter", False):
            items = [_frame_info(p) for p in paths]
            min_n_frame = hparams.reduction_factor * hparams.min_iters                               # :44
            max_n_frame = hparams.reduction_factor * hparams.max_iters - hparams.reduction_factor    # :45
            kept = [(p, n) for p, n, n_tok in items if min_n_frame <= n <= max_n_frame and n_tok >= hparams.min_tokens]
            new_paths = [p for p, _ in kept]
            frames = [n for _, n in kept]
            if frames:
                hours = sum(frames) * hparams.frame_shift_ms / (3600 * 1000)                        # audio.frames_to_hours
                log(" [{}] Loaded metadata for {} examples ({:.2f} hours)".format(data_dir, len(frames), hours))
                log(" [{}] Max length: {}".format(data_dir, max(frames)))
                log(" [{}] Min length: {}".format(data_dir, min(frames)))
        else:
            new_paths = paths
        if data_type == "train":
            new_paths = new_paths[:-n_test] if n_test else new_paths
        elif data_type == "test":
            new_paths = new_paths[-n_test:] if n_test else new_paths
        else:
            raise Exception(" [!] Unkown data_type: {}".format(data_type))                          # :70
        path_dict[data_dir] = new_paths
    return path_dict


def batch_shapes(batch, reduction_factor: int) -> Tuple[int, int]:
    """(T_in, T_out) of the padded batch: longest token row; ``round_up(longest target + 1, r)`` (datafeeder.py:308-315)."""
    t_in = max(len(x[0]) for x in batch)
    t_out = _round_up(max(len(x[3]) for x in batch) + 1, reduction_factor)
    return t_in, t_out


def prepare_batch(batch, reduction_factor: int, rng, data_type: Optional[str] = None, out: Optional[dict] = None,
                  multi_speaker: bool = True) -> Dict[str, np.ndarray]:
    """datafeeder.py:289-305.  ``batch``: list of (tokens, loss_coeff, mel, linear, speaker_id, n_frames) tuples.
    Returns the feed as a dict keyed by the reference's placeholder names.  With ``out`` (arrays at least as large,
    e.g. pinned staging buffers) the batch is written in place and the returned arrays are views into it."""
    if data_type == "train":
        rng.shuffle(batch)                                                                          # :290-291
    n = len(batch)
    t_in, t_out = batch_shapes(batch, reduction_factor)
    n_mel, n_lin = batch[0][2].shape[1], batch[0][3].shape[1]

    def buf(name, shape, dtype):
        if out is None:
            return np.zeros(shape, dtype)
        flat = out[name].reshape(-1)
        need = int(np.prod(shape))
        if flat.size < need:
            raise ValueError("staging buffer '%s' too small: %d < %d elements" % (name, flat.size, need))
        v = flat[:need].reshape(shape)
        v[...] = _pad
        return v

    inputs = buf("inputs", (n, t_in), np.int32)
    mel = buf("mel_targets", (n, t_out, n_mel), np.float32)
    lin = buf("linear_targets", (n, t_out, n_lin), np.float32)
    lengths = buf("input_lengths", (n,), np.int32)
    coeff = buf("loss_coeff", (n,), np.float32)
    for i, x in enumerate(batch):
        inputs[i, :len(x[0])] = x[0]                                                                # _pad_input :318-319
        lengths[i] = len(x[0])
        coeff[i] = x[1]
        mel[i, :len(x[2])] = x[2]                                                                   # _pad_target :322-323
        lin[i, :len(x[3])] = x[3]
    feed = dict(inputs=inputs, input_lengths=lengths, loss_coeff=coeff, mel_targets=mel, linear_targets=lin)
    if multi_speaker:
        spk = buf("speaker_id", (n,), np.int32)
        for i, x in enumerate(batch):
            spk[i] = x[4]
        feed["speaker_id"] = spk
    return feed


class DataFeeder(threading.Thread):
    """Feeds batches on a background thread (datafeeder.py:76-286).

    ``config`` needs ``random_seed`` and ``skip_path_filter`` (train.py:287-291).  ``rank``/``world`` shard the example
    stream for data-parallel training; ``device`` (a torch.device) enables pinned staging + async copies."""

    def __init__(self, data_dirs: Sequence[str], hparams, config, batches_per_group: int, data_type: str, batch_size: int,
                 rank: int = 0, world: int = 1, device=None, depth: Optional[int] = None, log=print):
        super().__init__(daemon=True)
        self._hp = hparams
        self._step = 0
        self._offset = defaultdict(lambda: 2)                                                       # :87 (sic)
        self._batches_per_group = batches_per_group
        self.rank, self.world = rank, world
        self.rng = np.random.RandomState(config.random_seed + rank)
        self.data_type = data_type
        self.batch_size = batch_size
        self.min_tokens = hparams.min_tokens
        self.min_n_frame = hparams.reduction_factor * hparams.min_iters
        self.max_n_frame = hparams.reduction_factor * hparams.max_iters - hparams.reduction_factor
        self.skip_path_filter = getattr(config, "skip_path_filter", False)
        self._log = log
        # the train/test split must be the same on every rank: it uses the un-offset seed
        self.path_dict = get_path_dict(data_dirs, hparams, config, data_type, n_test=batch_size,
                                       rng=np.random.RandomState(config.random_seed), log=log if rank == 0 else (lambda *_: None))
        self.data_dirs = list(self.path_dict.keys())
        self.data_dir_to_id = {d: i for i, d in enumerate(self.data_dirs)}
        weight = {d: 1.0 for d in self.data_dirs}                                                   # :109-125
        if hparams.main_data_greedy_factor > 0 and any(md in d for d in self.data_dirs for md in hparams.main_data):
            for md in hparams.main_data:
                for d in self.data_dirs:
                    if md in d:
                        weight[d] += hparams.main_data_greedy_factor
        z = sum(weight.values())
        self.data_ratio = {d: w / z for d, w in weight.items()}
        self.is_multi_speaker = len(self.data_dirs) > 1                                             # :149
        depth = depth or (8 if data_type == "train" else 1)                                         # :156-157
        self._queue: "queue.Queue" = queue.Queue(maxsize=depth)
        self._stop_event = threading.Event()
        self._error: Optional[BaseException] = None
        self._device = device
        self._slots: List[dict] = []
        self._free: "queue.Queue" = queue.Queue()
        self._depth = depth
        self._copy_stream = None
        self._inflight = None
        if data_type == "test":                                                                     # :179-191
            examples = []
            while len(examples) < batch_size:
                for d in self.data_dirs:
                    examples.append(self._get_next_example(d))
                    if len(examples) >= batch_size:
                        break
            self.static_batches = [examples for _ in range(batches_per_group)]
        else:
            self.static_batches = None

    # ---- staging ring ------------------------------------------------------------------------------------
    def _alloc_slots(self):
        """Pinned ring sized for the largest batch the length filter admits (allocated once; reused for every batch)."""
        import torch
        hp, n = self._hp, self.batch_size
        t_out = _round_up(self.max_n_frame + 1, hp.reduction_factor) if not self.skip_path_filter else None
        t_in = None
        if t_out is None or t_in is None:
            # without the filter (or for tokens) the bound comes from the data itself
            n_frames, n_tokens = [], []
            for paths in self.path_dict.values():
                for p in paths:
                    _, f, t = _frame_info(p)
                    n_frames.append(f); n_tokens.append(t)
            t_in = max(n_tokens)
            t_out = t_out or _round_up(max(n_frames) + 1, hp.reduction_factor)
        pin = self._device is not None and torch.cuda.is_available()
        mk = lambda shape, dt: (torch.empty(shape, dtype=dt).pin_memory() if pin else torch.empty(shape, dtype=dt))
        for _ in range(self._depth + 2):
            t = dict(inputs=mk((n, t_in), torch.int32), input_lengths=mk((n,), torch.int32), loss_coeff=mk((n,), torch.float32),
                     mel_targets=mk((n, t_out, hp.num_mels), torch.float32), linear_targets=mk((n, t_out, hp.num_freq), torch.float32),
                     speaker_id=mk((n,), torch.int32))
            slot = dict(torch=t, numpy={k: v.numpy() for k, v in t.items()})
            self._slots.append(slot)
            self._free.put(slot)

    # ---- thread ------------------------------------------------------------------------------------------
    def start_in_session(self, session=None, start_step: int = 0):                                  # :196-199
        self._step = start_step
        if not self._slots:
            self._alloc_slots()
        self.start()

    def stop(self):
        self._stop_event.set()

    def run(self):                                                                                  # :202-208
        try:
            while not self._stop_event.is_set():
                self._enqueue_next_group()
        except BaseException as e:      # surfaced to the consumer by next_batch()
            self._error = e
            self._queue.put(None)

    def _next_group(self) -> List[list]:
        """One group of batches (datafeeder.py:211-242)."""
        n, hp = self.batch_size, self._hp
        if self.static_batches is not None:
            return [list(b) for b in self.static_batches]
        examples = []
        for data_dir in self.data_dirs:
            if hp.initial_data_greedy:                                                              # :224-227
                if self._step < hp.initial_phase_step and any("krbook" in d for d in self.data_dirs):
                    data_dir = [d for d in self.data_dirs if "krbook" in d][0]
            if self._step < hp.initial_phase_step:                                                  # :229-234
                count = int(n * self._batches_per_group // len(self.data_dirs))
            else:
                count = int(n * self._batches_per_group * self.data_ratio[data_dir])
            examples.extend(self._get_next_example(data_dir) for _ in range(count))
        examples.sort(key=lambda x: x[-1])                                                          # :236 bucket by frame count
        batches = [examples[i:i + n] for i in range(0, len(examples), n)]
        self.rng.shuffle(batches)                                                                   # :239
        return batches

    def _enqueue_next_group(self):
        start = time.time()
        batches = self._next_group()
        self._log("Generated %d batches of size %d in %.03f sec" % (len(batches), self.batch_size, time.time() - start))
        r = self._hp.reduction_factor
        for batch in batches:
            slot = None
            while slot is None and not self._stop_event.is_set():
                try:
                    slot = self._free.get(timeout=0.1)
                except queue.Empty:
                    pass
            if slot is None:
                return
            feed = prepare_batch(batch, r, self.rng, self.data_type, out=slot["numpy"], multi_speaker=self.is_multi_speaker)
            shapes = {k: v.shape for k, v in feed.items()}
            while not self._stop_event.is_set():
                try:
                    self._queue.put((slot, shapes), timeout=0.1)
                    break
                except queue.Full:
                    pass
            self._step += 1

    def _get_next_example(self, data_dir: str):
        """datafeeder.py:245-286: (tokens, loss_coeff, mel, linear, speaker id, n_frames) of the next readable example."""
        data_paths = self.path_dict[data_dir]
        if not data_paths:
            raise RuntimeError("no training examples under %s" % data_dir)
        misses = 0
        while True:
            if self._offset[data_dir] >= len(data_paths):
                self._offset[data_dir] = 0
                if self.data_type == "train":
                    self.rng.shuffle(data_paths)
            data_path = data_paths[self._offset[data_dir]]
            self._offset[data_dir] += 1
            try:
                with np.load(data_path) as z:
                    data = {k: z[k] for k in z.files}
            except Exception:
                misses += 1
                if misses > 2 * len(data_paths) + 2:
                    raise RuntimeError("no readable example under %s" % data_dir)
                continue                                                                            # the reference deletes the file (:262-264)
            if not self.skip_path_filter:
                break
            if self.min_n_frame <= data["linear"].shape[0] <= self.max_n_frame and len(data["tokens"]) > self.min_tokens:
                break
        loss_coeff = data["loss_coeff"] if "loss_coeff" in data else 1
        return (data["tokens"], loss_coeff, data["mel"], data["linear"], self.data_dir_to_id[data_dir], len(data["linear"]))

    # ---- consumer side -----------------------------------------------------------------------------------
    def _take(self):
        item = self._queue.get()
        if item is None:
            raise RuntimeError("DataFeeder thread failed") from self._error
        return item

    def next_batch(self) -> Dict[str, "object"]:
        """Next padded batch as host tensors (pinned when a device was given), keyed like the reference's placeholders.
        The tensors are copies-free views of a ring slot that is recycled two calls later."""
        slot, shapes = self._take()
        out = self._views(slot, shapes)
        self._recycle(slot)
        return out

    @staticmethod
    def _views(slot, shapes):
        out = {}
        for k, shp in shapes.items():
            n = int(np.prod(shp))
            out[k] = slot["torch"][k].reshape(-1)[:n].view(*shp)
        return out

    def _recycle(self, slot):
        # a slot goes back to the producer one call later, so the views handed out stay valid until the next call
        prev, self._inflight = self._inflight, slot
        if prev is not None:
            self._free.put(prev)

    def next_device_batch(self, compute_stream=None, after=None, defer_wait: bool = False) -> Dict[str, "object"]:
        """Next batch on the device: async copies from the pinned slot on a dedicated copy stream; the compute stream is
        made to wait on them (so the copy of batch k+1 overlaps step k when called right after launching step k).
        ``after``: a CUDA event the copies are ordered behind (train.py passes "forward pass of step k enqueued": the DMA then
        runs beside the backward pass, where it costs 0.08 ms instead of 0.7 ms beside a forward pass).  ``defer_wait``: do not
        make the compute stream wait now — the event comes back as ``batch["_ready"]`` for the caller to wait on when it
        starts using the batch (otherwise work enqueued after this call would stall behind the copy)."""
        import torch
        if self._device is None:
            raise RuntimeError("DataFeeder was created without a device")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self._device)
        slot, shapes = self._take()
        host = self._views(slot, shapes)
        dev = {}
        with torch.cuda.stream(self._copy_stream):
            if after is not None:
                self._copy_stream.wait_event(after)
            for k, v in host.items():
                dev[k] = v.to(self._device, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        cs = compute_stream or torch.cuda.current_stream(self._device)
        if not defer_wait:
            cs.wait_event(done)
        for v in dev.values():
            v.record_stream(cs)
        if defer_wait:
            dev["_ready"] = done
        slot["event"] = done
        prev, self._inflight = self._inflight, slot
        if prev is not None:
            if prev.get("event") is not None:
                prev["event"].synchronize()      # its DMA has long finished; makes the recycle safe
            self._free.put(prev)
        return dev

"""Data preparation on the GPU — mirror of the reference's ``datasets/generate_data.py``.

    python -m ... datasets.generate_data <metadata.json|metadata.csv> [--data_dirname data]

Same inputs (``recognition.json`` / ``metadata.csv`` beside the audio, generate_data.py:38-50), same skip rules
(``ignore_recognition_level`` / ``recognition_loss_coeff``, :66-71,83-85), same output: one ``<audio name>.npz`` per
utterance holding ``linear[T,1025] f32, mel[T,80] f32, tokens i32, loss_coeff`` (:156-172), and the same report lines.
What changes is where the work happens: the reference fans utterances out to a ``ProcessPoolExecutor`` of librosa STFTs
(:30,95-99); here both spectrograms of an utterance come from ONE cuFFT-backed analysis call
(``taco_audio_spectrogram``: pre-emphasis, reflect-padded STFT, magnitude, mel projection, dB, normalise) and the worker
pool is gone (``--num_workers`` is accepted and ignored).

Not reproduced: the matplotlib histograms (:109-140) and the normalisation half of the Korean text front end — the tokenizer
is a parameter (``text_to_sequence=``; the command line passes ``text.text_to_sequence``, the jamo decomposition without the
reference's dictionary / number / English read-outs); metadata whose values are already token-id lists is used as is.  Audio decoding: 16-bit / 32-bit / float PCM ``.wav`` through ``scipy.io.wavfile`` (librosa's audioread
back ends are not here); a file at another rate is resampled with a polyphase filter, which is NOT sample-identical to
librosa 0.5.1's resampy kernel — prepare audio at ``hparams.sample_rate`` when exact parity with the reference's features
matters.
"""
from __future__ import annotations

import argparse
import json
import os
from collections import defaultdict
from fractions import Fraction
from typing import Callable, Optional

import numpy as np

from ..hparams import hparams as _default_hp
from .datafeeder import write_example


def load_audio(path: str, sample_rate: int) -> np.ndarray:
    """audio/__init__.py:12-13 (``librosa.core.load(path, sr=...)[0]``): mono float32 in [-1, 1] at ``sample_rate``."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(np.iinfo(x.dtype).max + 1)
    elif x.dtype.kind == "u":                                  # 8-bit PCM is unsigned
        x = (x.astype(np.float32) - 128.0) / 128.0
    else:
        x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)                                     # librosa.to_mono
    if sr != sample_rate:
        from scipy.signal import resample_poly
        fr = Fraction(sample_rate, sr).limit_denominator(1000)
        x = resample_poly(x, fr.numerator, fr.denominator).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def frames_to_hours(n_frames, hp) -> float:                   # audio/__init__.py:38-40
    return sum(n_frames) * hp.frame_shift_ms / (3600 * 1000)


def _read_metadata(path: str):
    if path.endswith("json"):
        with open(path) as f:
            return json.load(f)
    if path.endswith("csv"):
        info = {}
        with open(path) as f:
            for line in f:
                p, text = line.strip().split("|")
                info[p] = text
        return info
    raise Exception(" [!] Unkown metadata format: {}".format(path))


def build_from_path(config, hp=None, text_to_sequence: Optional[Callable] = None, device: int = 0, log=print):
    """generate_data.py:26-127.  Returns the list of frame counts of the examples written / found."""
    from ..audio import GriffinLim
    hp = hp or _default_hp
    log(" [!] Sampling rate: {}".format(hp.sample_rate))
    base_dir = os.path.dirname(config.metadata_path)
    data_dir = os.path.join(base_dir, config.data_dirname)
    os.makedirs(data_dir, exist_ok=True)
    info = {}
    for path, value in _read_metadata(config.metadata_path).items():
        new_path = path if os.path.exists(path) else os.path.join(base_dir, path)
        if not os.path.exists(new_path):
            log(" [!] Audio not found: {}".format([path, new_path]))
            continue
        info[new_path] = value
    loss_coeff = defaultdict(lambda: 1)
    for path, value in list(info.items()):
        if isinstance(value, list) and not (value and isinstance(value[0], int)):      # [text] or [text, recognised text]
            if hp.ignore_recognition_level == 1 and len(value) == 1 or hp.ignore_recognition_level == 2:
                loss_coeff[path] = hp.recognition_loss_coeff
            info[path] = value[0]
    log(" [!] Skip recognition level: {}".format(hp.ignore_recognition_level))

    front = None
    n_frames = []
    for audio_path, text in info.items():
        if hp.ignore_recognition_level > 0 and loss_coeff[audio_path] != 1:
            continue
        if isinstance(text, list):
            tokens = np.asarray(text, dtype=np.int32)
        else:
            if text_to_sequence is None:
                raise RuntimeError("metadata holds text: pass text_to_sequence= (the reference's text front end needs `jamo`, "
                                   "which this build does not ship) or store token-id lists in the metadata")
            try:
                tokens = np.asarray(text_to_sequence(text), dtype=np.int32)
            except Exception:
                continue                                                               # generate_data.py:90-93
        numpy_path = os.path.join(data_dir, os.path.basename(audio_path).rsplit(".", 1)[0] + ".npz")
        if os.path.exists(numpy_path):
            try:
                with np.load(numpy_path) as data:
                    n_frames.append(int(data["linear"].shape[0]))
                continue
            except Exception:
                os.remove(numpy_path)                                                  # unreadable: rebuild (:173-179)
        wav = load_audio(audio_path, hp.sample_rate)
        frames = 1 + len(wav) // int(hp.frame_shift_ms / 1000 * hp.sample_rate)
        if front is None or front.max_frames < frames:
            if front is not None:
                front.close()
            front = GriffinLim(hp, max_frames=max(2048, frames), device=device)
        linear, mel = front.spectrograms(wav)                                          # [T, 1025], [T, 80] on the device
        n_frame = int(linear.shape[0])
        if hp.skip_inadequate:                                                          # :163-168 (as written in the reference)
            lo = hp.reduction_factor * hp.min_iters
            hi = hp.reduction_factor * hp.max_iters - hp.reduction_factor
            if lo <= n_frame <= hi and len(tokens) >= hp.min_tokens:
                continue
        write_example(numpy_path, tokens, mel.cpu().numpy(), linear.cpu().numpy(), float(loss_coeff[audio_path]))
        n_frames.append(n_frame)
    if front is not None:
        front.close()
    if n_frames:
        log(" [*] Loaded metadata for {} examples ({:.2f} hours)".format(len(n_frames), frames_to_hours(n_frames, hp)))
        log(" [*] Max length: {}".format(max(n_frames)))
        log(" [*] Min length: {}".format(min(n_frames)))
        lo = hp.reduction_factor * hp.min_iters
        hi = hp.reduction_factor * hp.max_iters - hp.reduction_factor
        kept = [n for n in n_frames if lo <= n <= hi]
        log(" [*] After filtered: {} examples ({:.2f} hours)".format(len(kept), frames_to_hours(kept, hp)))
    return n_frames


def main(argv=None):
    parser = argparse.ArgumentParser(description="spectrogram")
    parser.add_argument("metadata_path", type=str)
    parser.add_argument("--data_dirname", type=str, default="data")
    parser.add_argument("--num_workers", type=int, default=None)       # accepted for command-line compatibility; unused
    from ..text import text_to_sequence           # jamo decomposition without the reference's text normalisation (see text.py)
    build_from_path(parser.parse_args(argv), text_to_sequence=text_to_sequence)


if __name__ == "__main__":
    main()

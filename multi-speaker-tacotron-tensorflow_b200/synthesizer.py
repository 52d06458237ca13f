"""Synthesis front door — mirror of the reference's ``synthesizer.py`` ``Synthesizer`` (synthesizer.py:24-207) and its
post-processing (``plot_graph_and_save_audio``, :209-288) on the CUDA engine + GPU Griffin-Lim.

    synthesizer = Synthesizer()
    synthesizer.load(checkpoint_path, num_speakers, checkpoint_step)
    audio = synthesizer.synthesize(texts=[...] | tokens=[[...]], base_path=..., speaker_ids=[...], attention_trim=False)

Text front end: the reference tokenises Korean/English text with ``text.text_to_sequence`` (needs ``jamo``; out of scope,
SURVEY.md §2).  ``tokens=`` is always accepted; ``texts=`` works when a ``text_to_sequence`` callable is supplied
(``Synthesizer(text_to_sequence=fn)``) and otherwise raises, naming what is missing.
"""
from __future__ import annotations

import io
import os
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .hparams import hparams as _default_hparams, load_hparams
from .models import create_model, get_most_recent_checkpoint
from .tf_checkpoint import load_any

EOS = 1                                                          # text/symbols.py: _pad=0, _eos=1


def attention_trim_frames(alignment: np.ndarray, sequence_len: int, reduction_factor: int) -> int:
    """Number of spectrogram frames to keep (synthesizer.py:242-262): walk the per-step attention argmax until the last
    input position has been attended ``min(count, 5)`` times or attention moves past it.  alignment: [T_in, T_dec]."""
    attention_argmax = alignment.argmax(0)
    end_idx = min(sequence_len - 1, int(max(attention_argmax)))
    max_counter = min(int((attention_argmax == end_idx).sum()), 5)
    end_idx_counter = 0
    jdx = 0
    for jdx, attend_idx in enumerate(attention_argmax):
        if len(attention_argmax) > jdx + 1:
            if attend_idx == end_idx:
                end_idx_counter += 1
            if attend_idx == end_idx and attention_argmax[jdx + 1] > end_idx:
                break
            if end_idx_counter >= max_counter:
                break
        else:
            break
    return reduction_factor * jdx + 3


def librosa_trim_end(audio: np.ndarray, top_db: float = 50.0, frame_length: int = 5120, hop_length: int = 256) -> int:
    """End index of ``librosa.effects.trim(audio, frame_length=5120, hop_length=256, top_db=50)`` — the only part of its result
    the reference uses (synthesizer.py:266-269: ``audio_out[:index[-1]]``).  librosa 0.5.1: frame-wise mean-square energy over
    uncentred frames (``feature.rmse(y=, n_fft=frame_length, hop_length=)``), in dB relative to the loudest frame
    (``logamplitude(mse, ref_power=np.max, top_db=None)``, floor 1e-10); frames above ``-top_db`` are non-silent; the end is
    ``min(len, (last_non_silent_frame + 1) * hop_length)``.  (librosa >= 0.6 centres the frames, which moves the cut by up to
    frame_length/2 samples; the reference pins 0.5.1.)  Host post-processing of a finished waveform, not GPU work."""
    y = np.asarray(audio, dtype=np.float64)
    if len(y) < frame_length:
        return len(y)
    n_frames = 1 + (len(y) - frame_length) // hop_length
    csum = np.concatenate([[0.0], np.cumsum(y * y)])
    starts = np.arange(n_frames) * hop_length
    mse = (csum[starts + frame_length] - csum[starts]) / frame_length
    db = 10.0 * np.log10(np.maximum(1e-10, mse)) - 10.0 * np.log10(max(1e-10, float(mse.max())))
    nz = np.flatnonzero(db > -top_db)
    if len(nz) == 0:
        return 0
    return int(min(len(y), (nz[-1] + 1) * hop_length))


def wav_bytes(wav: np.ndarray, sample_rate: int) -> bytes:
    """audio.save_audio to an in-memory file (synthesizer.py:283-288): peak-normalised 16-bit PCM."""
    from scipy.io import wavfile
    out = wav * (32767 / max(0.01, float(np.max(np.abs(wav)))))          # audio/__init__.py:26-28
    buf = io.BytesIO()
    wavfile.write(buf, sample_rate, out.astype(np.int16))
    return buf.getvalue()


def manual_alignments_from(alignments: np.ndarray, mode: int) -> np.ndarray:
    """The second-pass alignments of ``manual_attention_mode`` (reference: synthesizer.py:171-196), [N, T_dec, T_in].

    ``alignments`` is the first pass's [N, T_in, T_dec] history.  The reference takes ``alignments[idx].argmax(1)`` - for every
    INPUT position the decoder step that attended it most - and writes a one at (that step, position):
      mode 1 ("argmax one hot"): into zeros;  mode 3 ("prunning"): into the transposed soft alignments themselves.
    Mode 2 ("sharpening") calls ``np.pow``, which does not exist, and overwrites its own accumulator; it cannot run there."""
    if mode not in (1, 3):
        raise NotImplementedError("manual_attention_mode=2 calls np.pow in the reference (synthesizer.py:188), which does not exist")
    alignments = np.asarray(alignments, np.float32)
    alignments_T = np.transpose(alignments, [0, 2, 1])
    new = np.zeros_like(alignments_T) if mode == 1 else np.array(alignments_T, copy=True)
    for idx in range(len(alignments)):
        argmax = alignments[idx].argmax(1)                       # [T_in]: decoder step per input position
        new[idx][(argmax, np.arange(len(argmax)))] = 1
    return np.ascontiguousarray(new)


class Synthesizer:
    def __init__(self, hparams=None, precision: str = "tf32", device: int = 0,
                 text_to_sequence: Optional[Callable[[str], Sequence[int]]] = None):
        self.hparams = hparams or _default_hparams
        self._precision, self._device = precision, device
        self._text_to_sequence = text_to_sequence
        self.model = None
        self._gl = None

    def close(self):                                                     # synthesizer.py:25-27
        if self.model is not None and self.model.engine is not None:
            self.model.engine.close()
        if self._gl is not None:
            self._gl.close()
        self.model, self._gl = None, None

    def load(self, checkpoint_path, num_speakers=2, checkpoint_step=None, model_name="tacotron"):
        """synthesizer.py:29-68: resolve the checkpoint, load params.json beside it, build the model, restore the weights."""
        self.num_speakers = num_speakers
        if os.path.isdir(checkpoint_path):
            load_path = checkpoint_path
            if checkpoint_step is not None:
                checkpoint_path = os.path.join(load_path, "model.ckpt-{}.pt".format(checkpoint_step))
                if not os.path.exists(checkpoint_path):                  # a reference (TensorFlow) checkpoint of that step
                    checkpoint_path = checkpoint_path[:-3]
            else:
                checkpoint_path = get_most_recent_checkpoint(load_path)
        else:
            load_path = os.path.dirname(checkpoint_path)
        print("Constructing model: %s" % model_name)
        if os.path.exists(os.path.join(load_path, "params.json")):
            load_hparams(self.hparams, load_path)
        self.model = create_model(self.hparams)
        self.model._precision, self.model._device = self._precision, self._device
        print("Loading checkpoint: %s" % checkpoint_path)
        self._state = load_any(checkpoint_path, self.hparams, num_speakers)
        self._restored = False
        return self

    # ------------------------------------------------------------------------------------------------------
    def _run(self, sequences, input_lengths, speaker_ids, manual_alignments):
        hp = self.hparams
        m = self.model
        m.is_manual_attention = manual_alignments is not None
        m.manual_alignments = None if manual_alignments is None else torch.as_tensor(manual_alignments, dtype=torch.float32)
        inputs = torch.as_tensor(np.asarray(sequences), dtype=torch.int32)
        lengths = torch.as_tensor(np.asarray(input_lengths), dtype=torch.int32)
        spk = None
        if self.num_speakers > 1:
            spk = torch.as_tensor(np.asarray(speaker_ids if speaker_ids is not None else [0] * len(sequences)), dtype=torch.int32)
        if not self._restored:
            from .engine import Engine
            m.engine = Engine(hp, self.num_speakers, precision=self._precision, device=self._device)
            m.num_speakers = self.num_speakers
            m.load_state_dict(self._state)                                                        # saver.restore, synthesizer.py:66-68
            self._restored = True
        m.initialize(inputs, lengths, self.num_speakers, spk)                                     # synthesizer.py:47-54,166-167
        return m.linear_outputs.cpu().numpy(), m.alignments.cpu().numpy()

    def synthesize(self, texts=None, tokens=None, base_path=None, paths=None, speaker_ids=None,
                   start_of_sentence=None, end_of_sentence=True, pre_word_num=0, post_word_num=0,
                   pre_surplus_idx=0, post_surplus_idx=1, use_short_concat=False, manual_attention_mode=0,
                   base_alignment_path=None, librosa_trim=False, attention_trim=True) -> List[object]:
        """synthesizer.py:70-207.  Returns, per utterance, wav bytes (no path given) or True (written to disk)."""
        if isinstance(texts, str):
            texts = [texts]
        if texts is not None and tokens is None:
            if self._text_to_sequence is None:
                raise RuntimeError("texts= needs a text front end: pass Synthesizer(text_to_sequence=fn) (the reference's "
                                   "text.text_to_sequence depends on `jamo`, which this build does not ship) or call with tokens=")
            sequences = [list(self._text_to_sequence(t)) for t in texts]
        elif tokens is not None:
            sequences = [list(t) for t in tokens]
        else:
            raise ValueError("synthesize needs texts= or tokens=")
        if use_short_concat:
            # synthesizer.py:301-389 cuts at word boundaries found through the Korean text front end (decompose_ko_text) and reads
            # an undefined name (`decomposed_text`, :324,:334) on most of its branches; it is not reproduced
            raise NotImplementedError("use_short_concat depends on the reference's Korean text front end (out of scope, SURVEY.md §8f)")
        n = len(sequences)
        paths = paths if paths is not None else [None] * n
        texts = texts if texts is not None else [None] * n
        max_len = max(len(s) for s in sequences)
        seq = np.zeros((n, max_len), np.int32)
        for i, s in enumerate(sequences):
            seq[i, :len(s)] = s
        input_lengths = np.argmax(seq == EOS, 1)                         # synthesizer.py:118 (position of the EOS token)
        if (input_lengths == 0).any():
            raise ValueError("every token sequence must contain the EOS token (1) after its first symbol")
        manual = None
        if base_alignment_path is not None:                              # synthesizer.py:131-147
            apath = os.path.join(base_alignment_path, os.path.basename(base_path))
            manual = np.transpose(np.stack([np.load("{}.{}.npy".format(apath, i)) for i in range(n)]), [0, 2, 1])
        spectrograms, alignments = self._run(seq, input_lengths, speaker_ids, manual)
        self._librosa_trim = bool(librosa_trim)
        results = self._save(spectrograms, alignments, paths, sequences, base_path, end_of_sentence, attention_trim, manual is not None)
        if manual_attention_mode > 0:
            new = manual_alignments_from(alignments, manual_attention_mode)
            spectrograms, alignments = self._run(seq, input_lengths, speaker_ids, new)
            results = self._save(spectrograms, alignments, paths, sequences, base_path, end_of_sentence, attention_trim, True)
        return results

    def _save(self, spectrograms, alignments, paths, sequences, base_path, end_of_sentence, attention_trim, manual):
        from .audio import GriffinLim
        hp = self.hparams
        results = []
        for idx, (spec, alignment, path, sequence) in enumerate(zip(spectrograms, alignments, paths, sequences)):
            if attention_trim and end_of_sentence:
                spec = spec[:attention_trim_frames(alignment, len(sequence), hp.reduction_factor)]
            if self._gl is None or self._gl.max_frames < spec.shape[0]:
                self._gl = GriffinLim(hp, max_frames=max(1024, spec.shape[0]), device=self._device)
            audio_out = self._gl.inv_spectrogram(torch.from_numpy(np.ascontiguousarray(spec))).cpu().numpy()    # synthesizer.py:264
            if getattr(self, "_librosa_trim", False) and end_of_sentence:                                      # synthesizer.py:266-269
                audio_out = audio_out[:librosa_trim_end(audio_out)]
            if path or base_path:
                if path is None:
                    tag = ".manual" if manual else ""
                    path = os.path.join(base_path, "{}{}.wav".format(idx, tag))
                os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
                with open(path, "wb") as f:
                    f.write(wav_bytes(audio_out, hp.sample_rate))
                np.save(os.path.splitext(path)[0] + ".npy", alignment)
                results.append(True)
            else:
                results.append(wav_bytes(audio_out, hp.sample_rate))
        return results


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("--load_path", required=True)
    parser.add_argument("--sample_path", default="samples")
    parser.add_argument("--text", default=None)
    parser.add_argument("--tokens", default=None, help="space separated symbol ids ending with EOS=1 (replaces --text without a tokenizer)")
    parser.add_argument("--num_speakers", default=1, type=int)
    parser.add_argument("--speaker_id", default=0, type=int)
    parser.add_argument("--checkpoint_step", default=None, type=int)
    parser.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "bf16"])
    config = parser.parse_args(argv)
    os.makedirs(config.sample_path, exist_ok=True)
    from .text import text_to_sequence            # jamo decomposition without the reference's text normalisation (see text.py)
    synthesizer = Synthesizer(precision=config.precision, text_to_sequence=text_to_sequence)
    synthesizer.load(config.load_path, config.num_speakers, config.checkpoint_step)
    kw = dict(tokens=[[int(t) for t in config.tokens.split()]]) if config.tokens else dict(texts=[config.text])
    return synthesizer.synthesize(base_path=config.sample_path, speaker_ids=[config.speaker_id], attention_trim=False, **kw)[0]


if __name__ == "__main__":
    main()

"""GPU spectrogram inversion (Griffin-Lim) and analysis front end behind the C ABI — mirrors the functions of the
reference's audio/__init__.py that sit either side of the model: inv_spectrogram (:54-56), _griffin_lim (:76-84),
_stft_parameters (:118-122) on the synthesis side; spectrogram (:48-51), melspectrogram (:64-67), _build_mel_basis
(:141-143) on the data-preparation side (datasets/generate_data.py:57-58 calls them per utterance)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import capi
from .hparams import hparams as _default_hp, stft_parameters


def build_mel_basis(hp=None) -> np.ndarray:
    """[num_mels, 1 + n_fft/2] float32 — what librosa 0.5.1 ``filters.mel(sample_rate, n_fft, n_mels=num_mels)`` returns with
    its defaults (audio/__init__.py:141-143): triangular filters on the Slaney mel scale (linear below 1 kHz, log above),
    fmin 0, fmax sr/2, each triangle normalised to unit area.  A constant table, built once on the host."""
    hp = hp or _default_hp
    n_fft = (hp.num_freq - 1) * 2
    f_sp, brk, step = 200.0 / 3.0, 1000.0, np.log(6.4) / 27.0
    hz2mel = lambda f: np.where(f < brk, f / f_sp, brk / f_sp + np.log(np.maximum(f, 1e-30) / brk) / step)  # noqa: E731
    mel2hz = lambda m: np.where(m < brk / f_sp, m * f_sp, brk * np.exp(step * (m - brk / f_sp)))            # noqa: E731
    edges = mel2hz(np.linspace(hz2mel(np.float64(0.0)), hz2mel(np.float64(hp.sample_rate / 2.0)), hp.num_mels + 2))
    freqs = np.linspace(0.0, hp.sample_rate / 2.0, 1 + n_fft // 2)
    lo, ce, hi = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    tri = np.maximum(0.0, np.minimum((freqs[None] - lo) / (ce - lo), (hi - freqs[None]) / (hi - ce)))
    return (tri * (2.0 / (hi - lo))).astype(np.float32)


class GriffinLim:
    def __init__(self, hp=None, max_frames: int = 1024, device: int = 0):
        self.hp = hp or _default_hp
        if not torch.cuda.is_available():
            raise capi.TacoError("Griffin-Lim needs a CUDA device; there is no CPU fallback")
        self.lib = capi.load()
        self.dev = torch.device("cuda", device)
        self.n_fft, self.hop, self.win = stft_parameters(self.hp)
        self.max_frames = max_frames
        self._h = C.c_void_p()
        capi.check(self.lib.taco_gl_create(C.byref(self._h), self.n_fft, self.hop, self.win, max_frames, device))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.taco_gl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def inv_spectrogram(self, spectrogram: torch.Tensor, init_phase: Optional[torch.Tensor] = None, n_iters: Optional[int] = None):
        """spectrogram: [T, num_freq] normalised-dB linear spectrogram (device or host tensor) -> waveform [hop*(T-1)] (device).
        init_phase in [0,1) plays the role of the reference's np.random.rand(*S.shape); None = torch.rand."""
        hp = self.hp
        spec = spectrogram.to(device=self.dev, dtype=torch.float32).contiguous()
        T, F = spec.shape
        if F != hp.num_freq:
            raise capi.TacoError("spectrogram has %d bins, expected %d" % (F, hp.num_freq))
        if init_phase is None:
            init_phase = torch.rand(T, F, device=self.dev)
        ph = init_phase.to(device=self.dev, dtype=torch.float32).contiguous()
        wav = torch.empty(self.hop * (T - 1), dtype=torch.float32, device=self.dev)
        capi.check(self.lib.taco_gl_inv_spectrogram(
            self._h, spec.data_ptr(), ph.data_ptr(), T, hp.griffin_lim_iters if n_iters is None else n_iters,
            float(hp.power), float(hp.min_level_db), float(hp.ref_level_db), float(hp.preemphasis), wav.data_ptr(), None,
            torch.cuda.current_stream(self.dev).cuda_stream))
        return wav


    def _analyse(self, wav, want_linear: bool, want_mel: bool):
        hp = self.hp
        y = torch.as_tensor(wav).to(device=self.dev, dtype=torch.float32).contiguous()
        if y.dim() != 1:
            raise capi.TacoError("expected a mono waveform [n_samples]")
        T = 1 + y.numel() // self.hop
        lin = torch.empty(T, hp.num_freq, dtype=torch.float32, device=self.dev) if want_linear else None
        mel = torch.empty(T, hp.num_mels, dtype=torch.float32, device=self.dev) if want_mel else None
        if want_mel and getattr(self, "_mel_basis", None) is None:
            self._mel_basis = torch.from_numpy(build_mel_basis(hp)).to(self.dev)
        capi.check(self.lib.taco_audio_spectrogram(
            self._h, y.data_ptr(), y.numel(), float(hp.preemphasis), float(hp.ref_level_db), float(hp.min_level_db),
            self._mel_basis.data_ptr() if want_mel else None, hp.num_mels if want_mel else 0,
            None if lin is None else lin.data_ptr(), None if mel is None else mel.data_ptr(),
            torch.cuda.current_stream(self.dev).cuda_stream))
        return lin, mel

    def spectrogram(self, wav) -> torch.Tensor:
        """waveform [n] -> normalised-dB linear spectrogram [T, num_freq] on the device (T = 1 + n // hop); the layout the
        model's ``linear_targets`` use (the reference stores ``audio.spectrogram(wav).T``, datasets/generate_data.py)."""
        return self._analyse(wav, True, False)[0]

    def melspectrogram(self, wav) -> torch.Tensor:
        """waveform [n] -> normalised-dB mel spectrogram [T, num_mels] on the device."""
        return self._analyse(wav, False, True)[1]

    def spectrograms(self, wav):
        """Both targets of one utterance from a single STFT: (linear [T, num_freq], mel [T, num_mels])."""
        return self._analyse(wav, True, True)


_gl: Optional[GriffinLim] = None


def _shared(hp, frames: int) -> GriffinLim:
    global _gl
    if _gl is None or _gl.max_frames < frames or (hp is not None and hp is not _gl.hp):
        _gl = GriffinLim(hp, max_frames=max(1024, frames))
    return _gl


def spectrogram(y, hp=None):
    """Drop-in for the reference's audio.spectrogram(y): numpy [num_freq, T] (audio/__init__.py:48-51)."""
    y = np.asarray(y, dtype=np.float32)
    g = _shared(hp, 1 + len(y) // stft_parameters(hp or _default_hp)[1])
    return g.spectrogram(y).t().cpu().numpy()


def melspectrogram(y, hp=None):
    """Drop-in for the reference's audio.melspectrogram(y): numpy [num_mels, T] (audio/__init__.py:64-67)."""
    y = np.asarray(y, dtype=np.float32)
    g = _shared(hp, 1 + len(y) // stft_parameters(hp or _default_hp)[1])
    return g.melspectrogram(y).t().cpu().numpy()


def inv_spectrogram(spectrogram, hp=None):
    """Drop-in for audio.inv_spectrogram(spectrogram) of the reference, which takes [num_freq, T] (audio/__init__.py:54,
    called with linear_output.T at synthesizer.py:264) and returns a numpy waveform."""
    spec = torch.as_tensor(spectrogram).t()
    return _shared(hp, spec.shape[0]).inv_spectrogram(spec).cpu().numpy()

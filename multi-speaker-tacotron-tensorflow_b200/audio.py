"""GPU spectrogram inversion (Griffin-Lim) behind the C ABI — mirrors the functions of the reference's audio/__init__.py
that sit on the synthesis hot path: inv_spectrogram (:54-56), _griffin_lim (:76-84), _stft_parameters (:118-122)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import capi
from .hparams import hparams as _default_hp, stft_parameters


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


_gl: Optional[GriffinLim] = None


def inv_spectrogram(spectrogram, hp=None):
    """Drop-in for audio.inv_spectrogram(spectrogram) of the reference, which takes [num_freq, T] (audio/__init__.py:54,
    called with linear_output.T at synthesizer.py:264) and returns a numpy waveform."""
    global _gl
    spec = torch.as_tensor(spectrogram).t()
    if _gl is None or _gl.max_frames < spec.shape[0] or (hp is not None and hp is not _gl.hp):
        _gl = GriffinLim(hp, max_frames=max(1024, spec.shape[0]))
    return _gl.inv_spectrogram(spec).cpu().numpy()

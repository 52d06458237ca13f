"""CPU ORACLE (test infrastructure only) for the reference's spectrogram inversion.

PARITY PINNING: the arithmetic lives in librosa 0.5.1 (requirements.txt:61, not installable here) and scipy (installed).
This file restates librosa 0.5.1's ``stft``/``istft``/``filters.mel`` from its published source as called by the reference:
    audio/__init__.py:54-56  inv_spectrogram     audio/__init__.py:76-84   _griffin_lim (60 iterations)
    audio/__init__.py:99-106 _stft/_istft        audio/__init__.py:118-122 _stft_parameters
    audio/__init__.py:149    _db_to_amp          :158-159 inv_preemphasis  :164-165 _denormalize
    audio/__init__.py:48-51  spectrogram         :64-67 melspectrogram     :141-143 _build_mel_basis
It is pinned by (1) tests/golden/ref_audio_small.npz — the reference's OWN audio/__init__.py executed unmodified over a
librosa stand-in (oracle/tf1_shim/librosa, written independently of this file) and the real scipy.signal
(tools/make_reference_golden.py; tests/test_reference_golden.py requires this oracle to reproduce its spectrogram,
melspectrogram and seeded 4-iteration inv_spectrogram), which pins the reference's glue (dB / normalise / power / loop /
pre-emphasis filters through scipy) but not librosa's own kernels, restated twice and cross-checked; and (2) tests/test_oracle.py:
STFT->iSTFT round trip, spectral convergence, the scipy.signal.lfilter recurrence.

librosa 0.5.x quirk kept on purpose: ``stft`` stores conj(FFT) and ``istft`` conjugates again (the "match phase from
DPWE code" comment in its source).  The reference draws its initial phase from an UNSEEDED np.random.rand
(audio/__init__.py:77); here the phase is an explicit input so that the GPU path can be compared sample for sample.
"""
from __future__ import annotations

import numpy as np


def stft_parameters(num_freq=1025, frame_shift_ms=12.5, frame_length_ms=50, sample_rate=24000):
    n_fft = (num_freq - 1) * 2
    hop = int(frame_shift_ms / 1000 * sample_rate)
    win = int(frame_length_ms / 1000 * sample_rate)
    return n_fft, hop, win


def _window(n_fft, win):
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win) / win)      # scipy.signal.get_window('hann', win, fftbins=True)
    lpad = (n_fft - win) // 2                                         # librosa.util.pad_center
    return np.pad(w, (lpad, n_fft - win - lpad)).astype(np.float32)


def stft(y, n_fft, hop, win):
    """librosa 0.5.1 stft(center=True, pad_mode='reflect') -> [1 + n_fft/2, n_frames] complex64 (conjugated FFT)."""
    w = _window(n_fft, win)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = yp[idx] * w[:, None]
    return np.fft.fft(frames, axis=0)[: n_fft // 2 + 1].conj().astype(np.complex64)


def istft(S, hop, win):
    """librosa 0.5.1 istft(center=True): overlap-add with window-sum-square normalisation, trims n_fft/2 each side."""
    n_fft = 2 * (S.shape[0] - 1)
    w = _window(n_fft, win)
    n_frames = S.shape[1]
    out_len = n_fft + hop * (n_frames - 1)
    y = np.zeros(out_len, dtype=np.float32)
    wsum = np.zeros(out_len, dtype=np.float32)
    wsq = w * w
    for i in range(n_frames):
        spec = S[:, i]
        spec = np.concatenate((spec.conj(), spec[-2:0:-1]), 0)
        ytmp = w * np.fft.ifft(spec).real
        y[i * hop:i * hop + n_fft] += ytmp.astype(np.float32)
        wsum[i * hop:i * hop + n_fft] += wsq
    nz = wsum > np.finfo(np.float32).tiny
    y[nz] /= wsum[nz]
    return y[n_fft // 2: -(n_fft // 2)]


def lfilter_inv_preemphasis(x, coef=0.97):
    """scipy.signal.lfilter([1], [1, -coef], x): y[n] = x[n] + coef*y[n-1]."""
    y = np.empty_like(x, dtype=np.float64)
    acc = 0.0
    for n in range(len(x)):
        acc = float(x[n]) + coef * acc
        y[n] = acc
    return y.astype(np.float32)


def inv_spectrogram(spec_TF, init_phase_TF=None, n_iters=60, power=1.5, min_level_db=-100.0, ref_level_db=20.0,
                    preemphasis=0.97, num_freq=1025, frame_shift_ms=12.5, frame_length_ms=50, sample_rate=24000):
    """spec_TF: [T, num_freq] normalised-dB spectrogram (the model's linear_outputs for one utterance).
    init_phase_TF: [T, num_freq] in [0,1) (the reference uses np.random.rand) or None for zero phase."""
    n_fft, hop, win = stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate)
    S = np.power(10.0, ((np.clip(spec_TF, 0, 1) * -min_level_db) + min_level_db + ref_level_db) * 0.05)   # _denormalize, _db_to_amp
    S = (S ** power).T.astype(np.float32)                                                                 # [F, T]
    if init_phase_TF is None:
        angles = np.ones_like(S, dtype=np.complex64)
    else:
        angles = np.exp(2j * np.pi * init_phase_TF.T).astype(np.complex64)
    Sc = S.astype(np.complex64)
    y = istft(Sc * angles, hop, win)
    for _ in range(n_iters):
        est = stft(y, n_fft, hop, win)
        mag = np.abs(est)
        angles = np.where(mag > 0, est / np.maximum(mag, 1e-30), 1.0).astype(np.complex64)   # exp(1j*angle(est)), angle(0)=0
        y = istft(Sc * angles, hop, win)
    return lfilter_inv_preemphasis(y, preemphasis)


# --------------------------------------------------------------------------
# analysis front end (reference: audio/__init__.py:48-51 spectrogram, :64-67 melspectrogram)
# --------------------------------------------------------------------------
def preemphasis(x, coef=0.97):
    """scipy.signal.lfilter([1, -coef], [1], x): y[n] = x[n] - coef*x[n-1], y[0] = x[0]  (audio/__init__.py:155-156)."""
    x = np.asarray(x, dtype=np.float64)
    y = x.copy()
    y[1:] -= coef * x[:-1]
    return y


def slaney_mel_basis(sample_rate=24000, n_fft=2048, n_mels=80):
    """librosa 0.5.1 filters.mel(sr, n_fft, n_mels) with its defaults (fmin 0, fmax sr/2, Slaney scale, area norm) as
    called at audio/__init__.py:141-143.  Written bin by bin (triangles) rather than with librosa's ramp matrices."""
    f_sp, brk = 200.0 / 3.0, 1000.0
    step = np.log(6.4) / 27.0

    def to_mel(f):
        return f / f_sp if f < brk else brk / f_sp + np.log(f / brk) / step

    def to_hz(m):
        return m * f_sp if m < brk / f_sp else brk * np.exp(step * (m - brk / f_sp))

    nb = 1 + n_fft // 2
    edges = [to_hz(m) for m in np.linspace(to_mel(0.0), to_mel(sample_rate / 2.0), n_mels + 2)]
    B = np.zeros((n_mels, nb))
    for i in range(n_mels):
        lo, ce, hi = edges[i], edges[i + 1], edges[i + 2]
        for k in range(nb):
            f = k * (sample_rate / 2.0) / (nb - 1)
            if lo < f < hi:
                B[i, k] = min((f - lo) / (ce - lo), (hi - f) / (hi - ce)) * 2.0 / (hi - lo)
    return B


def _normalize(S, min_level_db):
    return np.clip((S - min_level_db) / -min_level_db, 0, 1)


def spectrogram(y, ref_level_db=20.0, min_level_db=-100.0, coef=0.97, num_freq=1025, frame_shift_ms=12.5, frame_length_ms=50,
                sample_rate=24000):
    """-> [num_freq, T] normalised dB linear spectrogram (audio/__init__.py:48-51; T = 1 + len(y)//hop)."""
    n_fft, hop, win = stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate)
    D = np.abs(stft(preemphasis(y, coef), n_fft, hop, win))
    return _normalize(20 * np.log10(np.maximum(1e-5, D)) - ref_level_db, min_level_db)


def melspectrogram(y, min_level_db=-100.0, coef=0.97, num_mels=80, num_freq=1025, frame_shift_ms=12.5, frame_length_ms=50,
                   sample_rate=24000):
    """-> [num_mels, T] (audio/__init__.py:64-67).  Unlike spectrogram() the reference does NOT subtract ref_level_db here."""
    n_fft, hop, win = stft_parameters(num_freq, frame_shift_ms, frame_length_ms, sample_rate)
    D = np.abs(stft(preemphasis(y, coef), n_fft, hop, win))
    M = slaney_mel_basis(sample_rate, n_fft, num_mels) @ D
    return _normalize(20 * np.log10(np.maximum(1e-5, M)), min_level_db)

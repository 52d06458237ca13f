"""TEST INFRASTRUCTURE — stand-in for the four librosa 0.5.1 entry points the reference's audio/__init__.py calls on the
spectrogram path (requirements.txt:61 pins librosa==0.5.1; the package is not installable here): ``stft``, ``istft``,
``filters.mel`` and the sub-module names the reference mentions at import time.  Restated from the librosa 0.5.1 sources
(core/spectrum.py, filters.py, core/time_frequency.py) independently of oracle/griffin_lim_oracle.py; used only by
tools/make_reference_golden.py to execute the reference's audio code unmodified."""
import numpy as np
import scipy.signal

from . import filters  # noqa: F401


class _Absent:
    def __getattr__(self, name):
        raise NotImplementedError("librosa.%s is outside the spectrogram path and not provided by the stand-in" % name)


core = output = _Absent()


def _pad_center(data, size):
    n = len(data)
    lpad = int((size - n) // 2)
    return np.pad(data, (lpad, int(size - n - lpad)), mode="constant")


def _window(win_length, n_fft):
    return _pad_center(scipy.signal.get_window("hann", win_length, fftbins=True), n_fft)


def stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True, dtype=np.complex64, pad_mode="reflect"):
    """core/spectrum.py stft: centred (reflect pad n_fft/2), periodic Hann of win_length zero-padded to n_fft, frames as
    columns, ``fft(...)[:1+n_fft/2].conj()`` (the 'match phase from DPWE code' conjugate of 0.5.x)."""
    win_length = win_length or n_fft
    hop_length = hop_length or win_length // 4
    w = _window(win_length, n_fft).reshape(-1, 1)
    if center:
        y = np.pad(y, int(n_fft // 2), mode=pad_mode)
    n_frames = 1 + int((len(y) - n_fft) / hop_length)
    frames = np.stack([y[i * hop_length:i * hop_length + n_fft] for i in range(n_frames)], axis=1)
    out = np.empty((1 + n_fft // 2, n_frames), dtype=dtype, order="F")
    out[:] = np.fft.fft(w * frames, axis=0)[:1 + n_fft // 2].conj()
    return out


def istft(stft_matrix, hop_length=None, win_length=None, window="hann", center=True, dtype=np.float32, length=None):
    """core/spectrum.py istft: per frame ifft of the Hermitian extension of conj(spec), times the window, overlap-added;
    divided by the window sum-square where it exceeds tiny; n_fft/2 trimmed from both ends."""
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    win_length = win_length or n_fft
    hop_length = hop_length or win_length // 4
    w = _window(win_length, n_fft)
    n_frames = stft_matrix.shape[1]
    expected = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected, dtype=dtype)
    ifft_window_sum = np.zeros(expected, dtype=dtype)
    ifft_window_square = w * w
    for i in range(n_frames):
        sample = i * hop_length
        spec = stft_matrix[:, i].flatten()
        spec = np.concatenate((spec.conj(), spec[-2:0:-1]), 0)
        ytmp = w * np.fft.ifft(spec).real
        y[sample:sample + n_fft] = y[sample:sample + n_fft] + ytmp
        ifft_window_sum[sample:sample + n_fft] += ifft_window_square
    approx_nonzero = ifft_window_sum > np.finfo(ifft_window_sum.dtype).tiny
    y[approx_nonzero] /= ifft_window_sum[approx_nonzero]
    if center:
        y = y[int(n_fft // 2):-int(n_fft // 2)]
    return y

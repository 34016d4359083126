"""Slaney mel filterbank (host, float64 design -> float32 matrix) and its band table.

The reference takes the matrix from ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)``
with librosa's defaults ``htk=False, norm="slaney"`` (ref:audiotools/core/audio_signal.py:1298-1331).
Filters are triangles on the Slaney mel axis (linear below 1 kHz, log above), area-normalised.
The device kernel (``csrc/spectral.cu``) consumes the dense matrix together with the first /
one-past-last non-zero column of every row, derived here from the actual non-zeros."""
import functools

import numpy as np

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    mel = f / _F_SP
    hi = f >= _MIN_LOG_HZ
    mel[hi] = _MIN_LOG_MEL + np.log(f[hi] / _MIN_LOG_HZ) / _LOGSTEP
    return mel


def mel_to_hz(m):
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    f = _F_SP * m
    hi = m >= _MIN_LOG_MEL
    f[hi] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m[hi] - _MIN_LOG_MEL))
    return f


@functools.lru_cache(None)
def mel_filters(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: float = None) -> np.ndarray:
    """float32 ``[n_mels, 1 + n_fft // 2]``."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    bins = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin)[0], hz_to_mel(fmax)[0], n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - bins[None, :]
    w = np.zeros((n_mels, bins.size), dtype=np.float32)
    for i in range(n_mels):
        rising = -ramps[i] / width[i]
        falling = ramps[i + 2] / width[i + 1]
        w[i] = np.maximum(0, np.minimum(rising, falling))
    w *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    w.setflags(write=False)
    return w


def band_table(fb: np.ndarray):
    """(lo[n_rows], hi[n_rows]) int32 with fb[m, k] == 0 for k outside [lo, hi)."""
    nz = fb != 0
    any_ = nz.any(axis=1)
    lo = np.where(any_, nz.argmax(axis=1), 0).astype(np.int32)
    hi = np.where(any_, fb.shape[1] - nz[:, ::-1].argmax(axis=1), 0).astype(np.int32)
    return lo, hi

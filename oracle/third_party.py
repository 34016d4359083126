"""Restated arithmetic of the reference's ABSENT third-party dependencies.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

None of this source is under /root/reference; the packages are unpinned in
``ref:setup.py:35-61`` and not installable here (no network).  Each function
restates the published algorithm of the release current at the reference
snapshot (julius 0.2.7, pyloudnorm 0.1.1, librosa 0.10.x) and names the
reference call site that reaches it.  They are written with the same torch /
numpy primitives and dtypes the originals use, so that rounding behaviour
follows the originals as closely as a restatement can.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# julius.core
# ----------------------------------------------------------------------------
def sinc(x: torch.Tensor) -> torch.Tensor:
    """julius.core.sinc: sin(x)/x with sinc(0)=1 (un-normalised sinc)."""
    return torch.where(
        x == 0, torch.tensor(1.0, device=x.device, dtype=x.dtype), torch.sin(x) / x
    )


def unfold(x: torch.Tensor, kernel_size: int, stride: int) -> torch.Tensor:
    """julius.core.unfold -- used at ref:audiotools/core/loudness.py:171.

    ``n = ceil((max(T, K) - K) / stride) + 1`` frames; the tail is ZERO padded
    to ``(n-1)*stride + K`` (this differs from pyloudnorm, which drops partial
    blocks).  Returns ``[..., n, K]``.
    """
    shape = list(x.shape)
    length = shape.pop(-1)
    n_frames = math.ceil((max(length, kernel_size) - kernel_size) / stride) + 1
    tgt_length = (n_frames - 1) * stride + kernel_size
    padded = F.pad(x, (0, tgt_length - length)).contiguous()
    strides = [padded.stride(d) for d in range(padded.dim())]
    last = strides.pop(-1)
    assert last == 1
    strides = strides + [stride, 1]
    return padded.as_strided(shape + [n_frames, kernel_size], strides)


def unfold_num_frames(length: int, kernel_size: int, stride: int) -> int:
    return math.ceil((max(length, kernel_size) - kernel_size) / stride) + 1


# ----------------------------------------------------------------------------
# julius.fftconv
# ----------------------------------------------------------------------------
def fft_conv1d(x: torch.Tensor, weight: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """julius.fftconv.fft_conv1d (cross-correlation, no padding), used at
    ref:audiotools/core/loudness.py:94 and inside julius low-pass banks when
    ``half_size > 32``.  julius does block-wise overlap-save with block =
    min(5*kernel, length); in exact arithmetic this equals ``F.conv1d`` and the
    oracle computes it with one full-length FFT in float64 (a *more* exact
    evaluation of the same sum), returned as float32.

    x: [B, C, T]; weight: [D, C, K] -> [B, D, (T-K)//stride + 1]
    """
    B, C, T = x.shape
    D, C2, K = weight.shape
    assert C == C2
    n = T
    X = torch.fft.rfft(x.double(), n)
    W = torch.fft.rfft(weight.double(), n)
    # correlation = conj(W) * X, summed over input channels
    Y = torch.einsum("bcf,dcf->bdf", X, W.conj())
    y = torch.fft.irfft(Y, n)[..., : T - K + 1]
    if stride != 1:
        y = y[..., ::stride]
    return y.to(x.dtype)


# ----------------------------------------------------------------------------
# julius.resample
# ----------------------------------------------------------------------------
def resample_kernel(old_sr: int, new_sr: int, zeros: int = 24, rolloff: float = 0.945):
    """julius.ResampleFrac._init_kernels.  Returns (kernel[new, K], width, old, new)
    with old/new gcd-reduced and K = 2*width + old."""
    gcd = math.gcd(old_sr, new_sr)
    old = old_sr // gcd
    new = new_sr // gcd
    sr = min(new, old)
    sr *= rolloff
    width = math.ceil(zeros * old / sr)
    idx = torch.arange(-width, width + old).float()
    kernels = []
    for i in range(new):
        t = (-i / new + idx / old) * sr
        t = t.clamp_(-zeros, zeros)
        t *= math.pi
        window = torch.cos(t / zeros / 2) ** 2
        kernel = sinc(t) * window
        kernel.div_(kernel.sum())
        kernels.append(kernel)
    return torch.stack(kernels), width, old, new


def resample_frac(x: torch.Tensor, old_sr: int, new_sr: int, zeros: int = 24,
                  rolloff: float = 0.945) -> torch.Tensor:
    """julius.resample_frac -- used at ref:audiotools/core/audio_signal.py:732-734.

    Replicate-pad (width, width+old), strided conv1d with one kernel per output
    phase, interleave the phases, keep floor(new*T/old) samples.
    """
    if old_sr == new_sr:
        return x
    kernel, width, old, new = resample_kernel(old_sr, new_sr, zeros, rolloff)
    if old == new:
        return x
    shape = x.shape
    length = x.shape[-1]
    x = x.reshape(-1, length)
    x = F.pad(x[:, None], (width, width + old), mode="replicate")
    ys = F.conv1d(x, kernel.to(x)[:, None, :], stride=old)
    y = ys.transpose(1, 2).reshape(list(shape[:-1]) + [-1])
    out_len = int(math.floor(new * length / old))
    return y[..., :out_len]


# ----------------------------------------------------------------------------
# julius.lowpass / julius.bands
# ----------------------------------------------------------------------------
class LowPassFilters:
    """julius.LowPassFilters(cutoffs, zeros) with stride=1, pad=True.

    ``cutoffs`` are normalised (Hz / sample_rate), in [0, 0.5]; elements may be
    python floats, numpy scalars or 1-element tensors exactly as the reference
    passes them (ref:audiotools/core/dsp.py:174-178 passes a 1-element tensor,
    ref:audiotools/core/effects.py:399 passes numpy float64 via SplitBands).
    """

    def __init__(self, cutoffs, zeros: float = 8):
        self.cutoffs = list(cutoffs)
        if min(self.cutoffs) < 0:
            raise ValueError("Minimum cutoff must be larger than zero.")
        if max(self.cutoffs) > 0.5:
            raise ValueError("A cutoff above 0.5 does not make sense.")
        self.zeros = zeros
        self.half_size = int(zeros / min([c for c in self.cutoffs if c > 0]) / 2)
        self.fft = self.half_size > 32
        window = torch.hann_window(2 * self.half_size + 1, periodic=False)
        time = torch.arange(-self.half_size, self.half_size + 1)
        filters = []
        for cutoff in self.cutoffs:
            if cutoff == 0:
                filter_ = torch.zeros_like(time).float()
            else:
                filter_ = 2 * cutoff * window * sinc(2 * cutoff * math.pi * time)
                filter_ = filter_ / filter_.sum()
            filters.append(filter_.float())
        self.filters = torch.stack(filters)[:, None]  # [n, 1, 2*half+1]

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        shape = list(x.shape)
        x = x.reshape(-1, 1, shape[-1])
        x = F.pad(x, (self.half_size, self.half_size), mode="replicate")
        if self.fft:
            out = fft_conv1d(x, self.filters.to(x))
        else:
            out = F.conv1d(x, self.filters.to(x))
        shape.insert(0, len(self.cutoffs))
        shape[-1] = out.shape[-1]
        return out.permute(1, 0, 2).reshape(shape)


class LowPassFilter:
    """julius.LowPassFilter -- used at ref:audiotools/core/dsp.py:178."""

    def __init__(self, cutoff, zeros: float = 8):
        self._lowpasses = LowPassFilters([cutoff], zeros=zeros)

    def __call__(self, x):
        return self._lowpasses(x)[0]


class HighPassFilter:
    """julius.HighPassFilter = x - lowpass(x) -- used at ref:audiotools/core/dsp.py:210."""

    def __init__(self, cutoff, zeros: float = 8):
        self._lowpasses = LowPassFilters([cutoff], zeros=zeros)

    def __call__(self, x):
        return x - self._lowpasses(x)[0]


def hz_to_mel_htk(f: float) -> float:
    return 2595 * math.log10(1 + f / 700)


def mel_to_hz_htk(m):
    return 700 * (10 ** (m / 2595) - 1)


def mel_frequencies(n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """julius.utils.mel_frequencies (HTK formula)."""
    low = hz_to_mel_htk(fmin)
    high = hz_to_mel_htk(fmax)
    mels = np.linspace(low, high, n_mels)
    return mel_to_hz_htk(mels)


class SplitBands:
    """julius.SplitBands(sample_rate, n_bands) -- used at
    ref:audiotools/core/effects.py:399-403.  Returns ``[n_bands, *x.shape]``."""

    def __init__(self, sample_rate: float, n_bands: int, zeros: float = 8):
        if not n_bands >= 1:
            raise ValueError("n_bands must be greater than one")
        if n_bands == 1:
            self.cutoffs = []
            self.lowpass = None
        else:
            self.cutoffs = list(mel_frequencies(n_bands + 1, 0, sample_rate / 2)[1:-1])
            self.lowpass = LowPassFilters([c / sample_rate for c in self.cutoffs], zeros=zeros)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.lowpass is None:
            return x[None]
        lows = self.lowpass(x)
        low = lows[0]
        bands = [low]
        for low_and_band in lows[1:]:
            bands.append(low_and_band - low)
            low = low_and_band
        bands.append(x - low)
        return torch.stack(bands)


# ----------------------------------------------------------------------------
# pyloudnorm
# ----------------------------------------------------------------------------
class IIRfilter:
    """pyloudnorm.iirfilter.IIRfilter coefficient design (RBJ-style biquads),
    reached from ref:audiotools/core/loudness.py:253-260 through
    ``pyloudnorm.Meter(rate)._filters``.  All float64."""

    def __init__(self, G, Q, fc, rate, filter_type, passband_gain=1.0):
        self.G, self.Q, self.fc, self.rate = G, Q, fc, rate
        self.filter_type = filter_type
        self.passband_gain = passband_gain
        self.b, self.a = self.generate_coefficients()

    def generate_coefficients(self):
        A = 10 ** (self.G / 40.0)
        w0 = 2.0 * np.pi * (self.fc / self.rate)
        alpha = np.sin(w0) / (2.0 * self.Q)
        if self.filter_type == "high_shelf":
            b0 = A * ((A + 1) + (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha)
            b1 = -2 * A * ((A - 1) + (A + 1) * np.cos(w0))
            b2 = A * ((A + 1) + (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha)
            a0 = (A + 1) - (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha
            a1 = 2 * ((A - 1) - (A + 1) * np.cos(w0))
            a2 = (A + 1) - (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha
        elif self.filter_type == "high_pass":
            b0 = (1 + np.cos(w0)) / 2
            b1 = -(1 + np.cos(w0))
            b2 = (1 + np.cos(w0)) / 2
            a0 = 1 + alpha
            a1 = -2 * np.cos(w0)
            a2 = 1 - alpha
        else:
            raise NotImplementedError(
                f"filter type {self.filter_type!r}: only the K-weighting stages are restated"
            )
        return np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0


def k_weighting_filters(rate: float, filter_class: str = "K-weighting") -> "OrderedDict[str, IIRfilter]":
    """pyloudnorm.Meter(rate).filter_class = 'K-weighting' -> ``_filters``.

    Insertion order matters: the shelf is applied first
    (ref:audiotools/core/loudness.py:115 iterates the dict)."""
    if filter_class != "K-weighting":
        raise NotImplementedError(
            f"filter_class {filter_class!r}: only 'K-weighting' is restated (SURVEY A.1)"
        )
    f = OrderedDict()
    f["high_shelf"] = IIRfilter(4.0, 1 / np.sqrt(2), 1500.0, rate, "high_shelf")
    f["high_pass"] = IIRfilter(0.0, 0.5, 38.0, rate, "high_pass")
    return f


class PyloudnormMeterShim:
    """Just enough of ``pyloudnorm.Meter`` for ref:audiotools/core/loudness.py:253-260."""

    def __init__(self, rate, filter_class="K-weighting", block_size=0.400):
        self.rate = rate
        self.block_size = block_size
        self.filter_class = filter_class

    @property
    def filter_class(self):
        return self._filter_class

    @filter_class.setter
    def filter_class(self, value):
        self._filters = k_weighting_filters(self.rate, value)
        self._filter_class = value


# ----------------------------------------------------------------------------
# librosa.filters.mel
# ----------------------------------------------------------------------------
def _hz_to_mel_slaney(frequencies):
    frequencies = np.asanyarray(frequencies, dtype=np.float64)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def _mel_to_hz_slaney(mels):
    mels = np.asanyarray(mels, dtype=np.float64)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def librosa_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32) -- used at
    ref:audiotools/core/audio_signal.py:1323-1331.  Returns ``[n_mels, 1+n_fft//2]``."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    min_mel = _hz_to_mel_slaney(fmin)
    max_mel = _hz_to_mel_slaney(fmax)
    mels = np.linspace(min_mel, max_mel, n_mels + 2)
    mel_f = _mel_to_hz_slaney(mels)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights

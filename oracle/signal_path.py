"""CPU restatement of the reference's first-party hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Functional, tensor-in / tensor-out restatement of
``ref:audiotools/core/{audio_signal,loudness,effects,dsp}.py`` for float32
``[B, C, T]`` CPU tensors.  Every function cites the reference lines it follows
and calls the same installed primitives the reference calls (``torch.stft``,
``torchaudio.functional.lfilter``, ``scipy.signal.get_window``); the absent
third-party arithmetic comes from ``oracle/third_party.py``.
"""
import math

import numpy as np
import scipy.signal
import torch
import torch.nn.functional as F

from . import third_party as tp

GAIN_FACTOR = np.log(10) / 20  # ref:audiotools/core/effects.py:12
MIN_LOUDNESS = -70  # ref:audiotools/core/loudness.py:265


# ----------------------------------------------------------------------------
# util
# ----------------------------------------------------------------------------
def ensure_tensor(x, ndim=None, batch_size=None):
    """ref:audiotools/core/util.py:56-89."""
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    if ndim is not None:
        assert x.ndim <= ndim
        while x.ndim < ndim:
            x = x.unsqueeze(-1)
    if batch_size is not None:
        if x.shape[0] != batch_size:
            shape = list(x.shape)
            shape[0] = batch_size
            x = x.expand(*shape)
    return x


# ----------------------------------------------------------------------------
# STFT / mel  (ref:audiotools/core/audio_signal.py)
# ----------------------------------------------------------------------------
def get_window(window_type: str, window_length: int) -> torch.Tensor:
    """ref:audiotools/core/audio_signal.py:1009-1039 (scipy, periodic, f64->f32)."""
    if window_type == "average":
        window = np.ones(window_length) / window_length
    elif window_type == "sqrt_hann":
        window = np.sqrt(scipy.signal.get_window("hann", window_length))
    else:
        window = scipy.signal.get_window(window_type, window_length)
    return torch.from_numpy(window).float()


def default_stft_params(sample_rate: int):
    """ref:audiotools/core/audio_signal.py:1064-1087."""
    win = int(2 ** (np.ceil(np.log2(0.032 * sample_rate))))
    return dict(window_length=win, hop_length=win // 4, window_type="hann",
                match_stride=False, padding_type="reflect")


def _resolve_stft(sample_rate, window_length, hop_length, window_type, match_stride, padding_type):
    d = default_stft_params(sample_rate)
    window_length = d["window_length"] if window_length is None else int(window_length)
    # NB the reference resolves hop from the *signal's* stft_params, whose default is
    # derived from the default window (ref :1167-1175), not from the passed window.
    hop_length = d["hop_length"] if hop_length is None else int(hop_length)
    window_type = d["window_type"] if window_type is None else window_type
    match_stride = d["match_stride"] if match_stride is None else match_stride
    padding_type = d["padding_type"] if padding_type is None else padding_type
    return window_length, hop_length, window_type, match_stride, padding_type


def compute_stft_padding(length: int, window_length: int, hop_length: int, match_stride: bool):
    """ref:audiotools/core/audio_signal.py:1089-1121 -> (right_pad, pad)."""
    if match_stride:
        assert hop_length == window_length // 4, "For match_stride, hop must equal n_fft // 4"
        right_pad = math.ceil(length / hop_length) * hop_length - length
        pad = (window_length - hop_length) // 2
    else:
        right_pad = 0
        pad = 0
    return right_pad, pad


def stft(audio: torch.Tensor, sample_rate: int, window_length=None, hop_length=None,
         window_type=None, match_stride=None, padding_type=None) -> torch.Tensor:
    """ref:audiotools/core/audio_signal.py:1123-1212.  [B,C,T] f32 -> [B,C,F,N] c64."""
    window_length, hop_length, window_type, match_stride, padding_type = _resolve_stft(
        sample_rate, window_length, hop_length, window_type, match_stride, padding_type)
    B, C, T = audio.shape
    window = get_window(window_type, window_length)
    right_pad, pad = compute_stft_padding(T, window_length, hop_length, match_stride)
    x = F.pad(audio, (pad, pad + right_pad), padding_type)
    s = torch.stft(x.reshape(-1, x.shape[-1]), n_fft=window_length, hop_length=hop_length,
                   window=window, return_complex=True, center=True)
    _, nf, nt = s.shape
    s = s.reshape(B, C, nf, nt)
    if match_stride:
        s = s[..., 2:-2]
    return s


def istft(stft_data: torch.Tensor, sample_rate: int, original_length: int, window_length=None,
          hop_length=None, window_type=None, match_stride=None, length=None) -> torch.Tensor:
    """ref:audiotools/core/audio_signal.py:1214-1296."""
    window_length, hop_length, window_type, match_stride, _ = _resolve_stft(
        sample_rate, window_length, hop_length, window_type, match_stride, None)
    window = get_window(window_type, window_length)
    nb, nch, nf, nt = stft_data.shape
    s = stft_data.reshape(nb * nch, nf, nt)
    right_pad, pad = compute_stft_padding(original_length, window_length, hop_length, match_stride)
    if length is None:
        length = original_length + 2 * pad + right_pad
    if match_stride:
        s = F.pad(s, (2, 2))
    audio = torch.istft(s, n_fft=window_length, hop_length=hop_length, window=window,
                        length=length, center=True)
    audio = audio.reshape(nb, nch, -1)
    if match_stride:
        audio = audio[..., pad: -(pad + right_pad)]
    return audio


def mel_filters(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """ref:audiotools/core/audio_signal.py:1298-1331 (librosa.filters.mel)."""
    return tp.librosa_mel(sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax)


def mel_spectrogram(audio, sample_rate, n_mels=80, mel_fmin=0.0, mel_fmax=None, **kwargs):
    """ref:audiotools/core/audio_signal.py:1333-1369.  NB magnitude, not power."""
    s = stft(audio, sample_rate, **kwargs)
    magnitude = torch.abs(s)
    nf = magnitude.shape[2]
    mel_basis = torch.from_numpy(mel_filters(sample_rate, 2 * (nf - 1), n_mels, mel_fmin, mel_fmax))
    mel = magnitude.transpose(2, -1) @ mel_basis.T
    return mel.transpose(-1, 2)


def log_mel(mel: torch.Tensor, clamp_eps: float = 1e-5, pow: float = 2.0) -> torch.Tensor:
    """ref:audiotools/metrics/spectral.py:187-190 -- the reference's log-mel definition."""
    return mel.clamp(clamp_eps).pow(pow).log10()


def mfcc(audio, sample_rate, n_mfcc=40, n_mels=80, log_offset=1e-6, **kwargs):
    """ref:audiotools/core/audio_signal.py:1398-1426."""
    from torchaudio.functional import create_dct

    mel = mel_spectrogram(audio, sample_rate, n_mels, **kwargs)
    mel = torch.log(mel + log_offset)
    dct_mat = create_dct(n_mfcc, n_mels, "ortho")
    return (mel.transpose(-1, -2) @ dct_mat).transpose(-1, -2)


def log_magnitude(stft_data, ref_value=1.0, amin=1e-5, top_db=80.0):
    """ref:audiotools/core/audio_signal.py:1457-1487."""
    magnitude = torch.abs(stft_data)
    amin = amin ** 2
    log_spec = 10.0 * torch.log10(magnitude.pow(2).clamp(min=amin))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = torch.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


# ----------------------------------------------------------------------------
# Spectral masks  (ref:audiotools/core/dsp.py:217-370) -- all on the complex STFT [B, C, F, N]
# ----------------------------------------------------------------------------
def _band_mask(stft_data, bins, lo, hi, val):
    """mag/phase masked_fill + recombination, ref:audiotools/core/dsp.py:241-264 (frequency) / :290-306 (time)."""
    mag, phase = torch.abs(stft_data), torch.angle(stft_data)
    lo = ensure_tensor(lo, ndim=mag.ndim)
    hi = ensure_tensor(hi, ndim=mag.ndim)
    assert torch.all(lo < hi)
    mask = (lo <= bins) & (bins < hi)
    mag = mag.masked_fill(mask, val)
    phase = phase.masked_fill(mask, val)
    return mag * torch.exp(1j * phase)


def mask_frequencies(stft_data, sample_rate, fmin_hz, fmax_hz, val=0.0):
    """ref:audiotools/core/dsp.py:217-264."""
    B, _, F, N = stft_data.shape
    bins_hz = torch.linspace(0, sample_rate / 2, F)[None, None, :, None].repeat(B, 1, 1, N)
    return _band_mask(stft_data, bins_hz, fmin_hz, fmax_hz, val)


def mask_timesteps(stft_data, signal_duration, tmin_s, tmax_s, val=0.0):
    """ref:audiotools/core/dsp.py:266-306."""
    B, _, F, N = stft_data.shape
    bins_t = torch.linspace(0, signal_duration, N)[None, None, None, :].repeat(B, 1, F, 1)
    return _band_mask(stft_data, bins_t, tmin_s, tmax_s, val)


def mask_low_magnitudes(stft_data, db_cutoff, val=0.0):
    """ref:audiotools/core/dsp.py:308-333 (+ the magnitude setter, ref:audiotools/core/audio_signal.py:1452-1455)."""
    mag = torch.abs(stft_data)
    log_mag = log_magnitude(stft_data)
    db_cutoff = ensure_tensor(db_cutoff, ndim=mag.ndim)
    mag = mag.masked_fill(log_mag < db_cutoff, val)
    return mag * torch.exp(1j * torch.angle(stft_data))


def shift_phase(stft_data, shift):
    """ref:audiotools/core/dsp.py:335-351 (+ the phase setter, ref:audiotools/core/audio_signal.py:1512-1516)."""
    phase = torch.angle(stft_data)
    shift = ensure_tensor(shift, ndim=phase.ndim)
    return torch.abs(stft_data) * torch.exp(1j * (phase + shift))


# ----------------------------------------------------------------------------
# Loudness  (ref:audiotools/core/loudness.py)
# ----------------------------------------------------------------------------
class Meter:
    """ref:audiotools/core/loudness.py:11-260 (``Meter``), CPU."""

    def __init__(self, rate: int, filter_class: str = "K-weighting", block_size: float = 0.400,
                 zeros: int = 512, use_fir: bool = False):
        self.rate = rate
        self.block_size = block_size
        self.use_fir = use_fir
        self._filters = tp.k_weighting_filters(rate, filter_class)  # :253-260
        self.G = torch.from_numpy(np.array([1.0, 1.0, 1.0, 1.41, 1.41]))  # :49-50 (float64)
        impulse = np.zeros((zeros,))
        impulse[..., 0] = 1.0
        firs = np.zeros((len(self._filters), 1, zeros))
        passband_gain = torch.zeros(len(self._filters))
        for i, (_, st) in enumerate(self._filters.items()):
            firs[i] = scipy.signal.lfilter(st.b, st.a, impulse)
            passband_gain[i] = st.passband_gain
        self.firs = torch.from_numpy(firs[..., ::-1].copy()).float()
        self.passband_gain = passband_gain

    def apply_filter_gpu(self, data):
        """:69-100 -- 512-tap FIR approximation (NOT the parity target)."""
        nb, nt, nch = data.shape
        data = data.permute(0, 2, 1).reshape(nb * nch, 1, nt)
        pad_length = self.firs.shape[-1]
        for i in range(self.firs.shape[0]):
            data = F.pad(data, (pad_length, pad_length))
            data = tp.fft_conv1d(data, self.firs[i, None, ...])
            data = self.passband_gain[i] * data
            data = data[..., 1: nt + 1]
        # Literal restatement: the reference permutes the *folded* [nb*nch, 1, nt] tensor, so for
        # nch > 1 channels stay folded into the batch dim ([nb*nch, nt, 1]) -- a reference quirk of
        # the FIR/GPU path only.  The IIR path (``apply_filter_cpu``) is the parity target.
        data = data.permute(0, 2, 1)
        return data[:, :nt, :]

    def apply_filter_cpu(self, data):
        """:102-126 -- exact IIR (float32 sequential recursion in torchaudio)."""
        import torchaudio

        for _, st in self._filters.items():
            a = torch.from_numpy(st.a).float()
            b = torch.from_numpy(st.b).float()
            _data = data.permute(0, 2, 1)
            filtered = torchaudio.functional.lfilter(_data, a, b, clamp=False)
            data = st.passband_gain * filtered.permute(0, 2, 1)
        return data

    def apply_filter(self, data):
        """:128-147 (is_cuda is never true for the oracle)."""
        return self.apply_filter_gpu(data) if self.use_fir else self.apply_filter_cpu(data)

    def _unfold(self, input_data):
        """:164-174."""
        T_g = self.block_size
        step = 1.0 - 0.75
        kernel_size = int(T_g * self.rate)
        stride = int(T_g * self.rate * step)
        unfolded = tp.unfold(input_data.permute(0, 2, 1), kernel_size, stride)
        return unfolded.transpose(-1, -2)

    def block_energies(self, data):
        """z of :214 (before gating): [nb, nch, nblk] float32."""
        input_data = self.apply_filter(data.float())
        unfolded = self._unfold(input_data)
        return (1.0 / (self.block_size * self.rate)) * unfolded.square().sum(2)

    def integrated_loudness(self, data):
        """:176-247.  data: [nb, nt, nch]."""
        data = data.float()
        input_data = data
        if input_data.ndim < 2:
            input_data = input_data.unsqueeze(-1)
        if input_data.ndim < 3:
            input_data = input_data.unsqueeze(0)
        nb, nt, nch = input_data.shape
        z = self.block_energies(input_data)
        return gate_blocks(z, self.G)


def gate_blocks(z: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    """ref:audiotools/core/loudness.py:208-247 given block energies ``z`` [nb, nch, nblk]
    (float32; modified in place exactly like the reference's aliasing does)."""
    nb, nch, _ = z.shape
    Gamma_a = -70.0
    l = -0.691 + 10.0 * torch.log10((G[None, :nch, None] * z).sum(1, keepdim=True))
    l = l.expand_as(z)
    z_avg_gated = z
    z_avg_gated[l <= Gamma_a] = 0
    masked = l > Gamma_a
    z_avg_gated = z_avg_gated.sum(2) / masked.sum(2)
    Gamma_r = -0.691 + 10.0 * torch.log10((z_avg_gated * G[None, :nch]).sum(-1)) - 10.0
    Gamma_r = Gamma_r[:, None, None].expand(nb, nch, l.shape[-1])
    z_avg_gated = z
    z_avg_gated[l <= Gamma_a] = 0
    z_avg_gated[l <= Gamma_r] = 0
    masked = (l > Gamma_a) * (l > Gamma_r)
    z_avg_gated = z_avg_gated.sum(2) / masked.sum(2)
    z_avg_gated = torch.where(z_avg_gated.isnan(), torch.zeros_like(z_avg_gated), z_avg_gated)
    z_avg_gated[z_avg_gated == float("inf")] = float(np.finfo(np.float32).max)
    z_avg_gated[z_avg_gated == -float("inf")] = float(np.finfo(np.float32).min)
    LUFS = -0.691 + 10.0 * torch.log10((G[None, :nch] * z_avg_gated).sum(1))
    return LUFS.float()


def loudness(audio: torch.Tensor, sample_rate: int, filter_class="K-weighting",
             block_size=0.400, **kwargs) -> torch.Tensor:
    """ref:audiotools/core/loudness.py:268-320.  [B,C,T] -> [B] f32, clamped to >= -70."""
    T = audio.shape[-1]
    if T / sample_rate < 0.5:
        pad_len = int((0.5 - T / sample_rate) * sample_rate)
        audio = F.pad(audio, (0, pad_len))
    meter = Meter(sample_rate, filter_class=filter_class, block_size=block_size, **kwargs)
    l = meter.integrated_loudness(audio.permute(0, 2, 1))
    return torch.maximum(l, torch.ones_like(l) * MIN_LOUDNESS)


def normalize(audio: torch.Tensor, sample_rate: int, db=-24.0, **loudness_kwargs):
    """ref:audiotools/core/effects.py:200-220.  Returns (audio*gain, measured LUFS)."""
    db = ensure_tensor(db)
    ref_db = loudness(audio, sample_rate, **loudness_kwargs)
    gain = db - ref_db
    gain = torch.exp(gain * GAIN_FACTOR)
    return audio * gain[:, None, None], ref_db


def volume_change(audio: torch.Tensor, db):
    """ref:audiotools/core/effects.py:222-238."""
    db = ensure_tensor(db, ndim=1)
    gain = torch.exp(db * GAIN_FACTOR)
    return audio * gain[:, None, None]


# ----------------------------------------------------------------------------
# Resample / FIR filters
# ----------------------------------------------------------------------------
def resample(audio: torch.Tensor, old_sr: int, new_sr: int) -> torch.Tensor:
    """ref:audiotools/core/audio_signal.py:716-736."""
    if new_sr == old_sr:
        return audio
    return tp.resample_frac(audio, old_sr, new_sr)


def low_pass(audio: torch.Tensor, sample_rate: int, cutoffs, zeros: int = 51):
    """ref:audiotools/core/dsp.py:153-183 (per-item filter, python loop)."""
    B = audio.shape[0]
    cutoffs = ensure_tensor(cutoffs, 2, B)
    cutoffs = cutoffs / sample_rate
    filtered = torch.empty_like(audio)
    for i, cutoff in enumerate(cutoffs):
        filtered[i] = tp.LowPassFilter(cutoff.cpu(), zeros=zeros)(audio[i])
    return filtered


def high_pass(audio: torch.Tensor, sample_rate: int, cutoffs, zeros: int = 51):
    """ref:audiotools/core/dsp.py:185-215."""
    B = audio.shape[0]
    cutoffs = ensure_tensor(cutoffs, 2, B)
    cutoffs = cutoffs / sample_rate
    filtered = torch.empty_like(audio)
    for i, cutoff in enumerate(cutoffs):
        filtered[i] = tp.HighPassFilter(cutoff.cpu(), zeros=zeros)(audio[i])
    return filtered


def mel_filterbank(audio: torch.Tensor, sample_rate: int, n_bands: int):
    """ref:audiotools/core/effects.py:386-403 -> [B,C,T,n_bands]."""
    return tp.SplitBands(sample_rate, n_bands)(audio).permute(1, 2, 3, 0)


def equalizer(audio: torch.Tensor, sample_rate: int, db):
    """ref:audiotools/core/effects.py:405-433 (weights are 10**db, sic)."""
    db = ensure_tensor(db)
    n_bands = db.shape[-1]
    fbank = mel_filterbank(audio, sample_rate, n_bands)
    if db.ndim == 2:
        if db.shape[0] != 1:
            assert db.shape[0] == fbank.shape[0]
    else:
        db = db.unsqueeze(0)
    weights = (10 ** db).float()
    fbank = fbank * weights[:, None, None, :]
    return fbank.sum(-1)


# ----------------------------------------------------------------------------
# IR convolution  (ref:audiotools/core/effects.py)
# ----------------------------------------------------------------------------
def ensure_max_of_audio(audio: torch.Tensor, max: float = 1.0):
    """ref:audiotools/core/effects.py:181-198."""
    peak = audio.abs().max(dim=-1, keepdims=True)[0]
    peak_gain = torch.ones_like(peak)
    peak_gain[peak > max] = max / peak[peak > max]
    return audio * peak_gain


def convolve(audio: torch.Tensor, ir: torch.Tensor, start_at_max: bool = True):
    """ref:audiotools/core/effects.py:66-123 -- CIRCULAR convolution, n = T."""
    T = audio.shape[-1]
    pad_len = T - ir.shape[-1]
    if pad_len > 0:
        ir = F.pad(ir, (0, pad_len))
    else:
        ir = ir[..., :T]
    if start_at_max:
        idx = ir.abs().argmax(axis=-1)
        irs = torch.zeros_like(ir)
        for i in range(ir.shape[0]):
            irs[i] = torch.roll(ir[i], -idx[i].item(), -1)
        ir = irs
    delta = torch.zeros_like(ir)
    delta[..., 0] = 1
    length = T
    delta_fft = torch.fft.rfft(delta, length)
    other_fft = torch.fft.rfft(ir, length)
    self_fft = torch.fft.rfft(audio, length)
    convolved = torch.fft.irfft(other_fft * self_fft, length)
    delta_audio = torch.fft.irfft(other_fft * delta_fft, length)
    delta_max = delta_audio.abs().max(dim=-1, keepdims=True)[0]
    scale = 1 / delta_max.clamp(1e-5)
    return convolved * scale


def decompose_ir(ir: torch.Tensor, sample_rate: int):
    """ref:audiotools/core/effects.py:540-577."""
    B = ir.shape[0]
    td = torch.argmax(ir, dim=-1, keepdim=True)
    t0 = int(sample_rate * 0.0025)
    idx = torch.arange(ir.shape[-1])[None, None, :].expand(B, -1, -1)
    early_idx = (idx >= td - t0) * (idx <= td + t0)
    early = torch.zeros_like(ir)
    early[early_idx] = ir[early_idx]
    late_idx = ~early_idx
    late = torch.zeros_like(ir)
    late[late_idx] = ir[late_idx]
    window = torch.zeros_like(ir)
    for b in range(B):
        window_idx = early_idx[b, 0].nonzero()
        window[b, ..., window_idx] = get_window("hann", window_idx.shape[-1])
    return early, late, window


def measure_drr(ir, sample_rate):
    """ref:audiotools/core/effects.py:579-592."""
    early, late, _ = decompose_ir(ir, sample_rate)
    return 10 * torch.log10((early ** 2).sum(dim=-1) / (late ** 2).sum(dim=-1))


def solve_alpha(early, late, wd, target_drr):
    """ref:audiotools/core/effects.py:594-617."""
    wd_sq = wd ** 2
    wd_sq_1 = (1 - wd) ** 2
    e_sq = early ** 2
    l_sq = late ** 2
    a = (wd_sq * e_sq).sum(dim=-1)
    b = (2 * (1 - wd) * wd * e_sq).sum(dim=-1)
    c = (wd_sq_1 * e_sq).sum(dim=-1) - torch.pow(10, target_drr / 10) * l_sq.sum(dim=-1)
    expr = ((b ** 2) - 4 * a * c).sqrt()
    return torch.maximum((-b - expr) / (2 * a), (-b + expr) / (2 * a))


def alter_drr(ir, sample_rate, drr):
    """ref:audiotools/core/effects.py:619-647."""
    drr = ensure_tensor(drr, 2, ir.shape[0])
    early, late, window = decompose_ir(ir, sample_rate)
    alpha = solve_alpha(early, late, window, drr)
    min_alpha = late.abs().max(dim=-1)[0] / early.abs().max(dim=-1)[0]
    alpha = torch.maximum(alpha, min_alpha)[..., None]
    aug = alpha * window * early + ((1 - window) * early) + late
    return ensure_max_of_audio(aug)


def apply_ir(audio, ir, sample_rate, drr=None, ir_eq=None):
    """ref:audiotools/core/effects.py:125-179 with use_original_phase=False."""
    if ir_eq is not None:
        ir = equalizer(ir, sample_rate, ir_eq)
    if drr is not None:
        ir = alter_drr(ir, sample_rate, drr)
    max_spk = audio.abs().max(dim=-1, keepdims=True).values
    out = convolve(audio, ir)
    max_transformed = out.abs().max(dim=-1, keepdims=True).values
    scale_factor = max_spk.clamp(1e-8) / max_transformed.clamp(1e-8)
    return out * scale_factor

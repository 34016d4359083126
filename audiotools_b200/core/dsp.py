"""``DSPMixin``: per-item windowed-sinc low-/high-pass (ref:audiotools/core/dsp.py:153-215).

The reference loops over the batch in Python, building one ``julius.LowPassFilter`` per item;
here the per-item tap design and the filtering are one grouped launch each (``csrc/fir.cu``).
The spectral masks (ref:audiotools/core/dsp.py:217-370) work on ``stft_data``: the two band
masks run as one store-only kernel (``csrc/specmask.cu``), the phase operations are container arithmetic on the
complex tensor.  Windowing / overlap-add helpers of the reference's DSPMixin (``collect_windows``) are not mirrored."""
import torch

from . import util


def _engine():
    from ..engine import get_engine

    return get_engine()


class DSPMixin:
    def low_pass(self, cutoffs, zeros: int = 51):
        """Low-pass each item at its own cutoff (Hz)."""
        cutoffs = util.ensure_tensor(cutoffs, 2, self.batch_size)
        self.audio_data = _engine().sinc_filter(self._materialized(), cutoffs[:, 0], self.sample_rate, zeros,
                                                highpass=False)
        self.stft_data = None
        return self

    def high_pass(self, cutoffs, zeros: int = 51):
        """High-pass each item at its own cutoff (Hz): ``x - low_pass(x)``."""
        cutoffs = util.ensure_tensor(cutoffs, 2, self.batch_size)
        self.audio_data = _engine().sinc_filter(self._materialized(), cutoffs[:, 0], self.sample_rate, zeros,
                                                highpass=True)
        self.stft_data = None
        return self

    # ------------------------------------------------------------------ spectral masks (ref :217-370)
    def _band_mask(self, lo, hi, axis_vals, axis: int, val: float):
        if self.stft_data is None:
            self.stft()
        # ref :249 / :300 ``assert torch.all(lo < hi)``: checked on the host mirrors when there are some (no sync)
        h_lo, h_hi = torch.as_tensor(util.host_view(lo)).float(), torch.as_tensor(util.host_view(hi)).float()
        assert bool(torch.all(h_lo < h_hi)), "mask band must satisfy min < max"
        lo = util.ensure_tensor(lo, ndim=1).float().reshape(-1)
        hi = util.ensure_tensor(hi, ndim=1).float().reshape(-1)
        spec = self.stft_data
        if spec.dtype != torch.complex64 or not spec.is_contiguous():
            spec = spec.to(torch.complex64).contiguous()
        self.stft_data = _engine().spec_band_mask(spec, axis_vals, lo, hi, axis, val)
        return self

    def mask_frequencies(self, fmin_hz, fmax_hz, val: float = 0.0):
        """Fill the band ``fmin_hz <= f < fmax_hz`` (per item) of ``stft_data`` with ``val`` (magnitude and phase):
        SpecAugment (ref :217-264).  Cells outside the band keep their value; the reference rebuilds them as
        ``|X| exp(1j angle X)``, which differs from X by float32 rounding only."""
        if self.stft_data is None:
            self.stft()
        nbins = self.stft_data.shape[-2]
        bins_hz = torch.linspace(0, self.sample_rate / 2, nbins, device=self.device)  # the reference's own grid
        return self._band_mask(fmin_hz, fmax_hz, bins_hz, 0, val)

    def mask_timesteps(self, tmin_s, tmax_s, val: float = 0.0):
        """Fill the frames ``tmin_s <= t < tmax_s`` (per item) of ``stft_data`` with ``val`` (ref :266-306)."""
        if self.stft_data is None:
            self.stft()
        nt = self.stft_data.shape[-1]
        bins_t = torch.linspace(0, self.signal_duration, nt, device=self.device)
        return self._band_mask(tmin_s, tmax_s, bins_t, 1, val)

    def mask_low_magnitudes(self, db_cutoff, val: float = 0.0):
        """Fill magnitudes whose ``log_magnitude()`` is below ``db_cutoff`` (per item) with ``val`` (ref :308-333)."""
        mag = self.magnitude
        log_mag = self.log_magnitude()
        db_cutoff = util.ensure_tensor(db_cutoff, ndim=mag.ndim).to(mag.device)
        self.magnitude = mag.masked_fill(log_mag < db_cutoff, val)
        return self

    def shift_phase(self, shift):
        """``phase += shift`` (scalar, per item, or a full [B, C, F, N] tensor) (ref :335-351)."""
        shift = util.ensure_tensor(shift, ndim=self.phase.ndim).to(self.device)
        self.phase = self.phase + shift
        return self

    def corrupt_phase(self, scale):
        """``phase += scale * N(0, 1)`` drawn on the signal's device (ref :353-369)."""
        scale = util.ensure_tensor(scale, ndim=self.phase.ndim).to(self.device)
        self.phase = self.phase + scale * torch.randn_like(self.phase)
        return self

"""Spectral noise gate (ref:audiotools/ml/layers/spectral_gate.py:10-127; after noisereduce / Audacity's noise
reduction): a per-bin threshold from a noise excerpt's STFT statistics, a smoothed binary mask, applied to the
signal's STFT.  Both STFTs and the inverse run on the engine (``csrc/spectral.cu``, ``csrc/istft.cu``); the mask
algebra in between is tensor arithmetic on the device."""
import torch
from torch import nn

from ...core import AudioSignal
from ...core import STFTParams
from ...core import util


def _ramp(n: int) -> torch.Tensor:
    """0 < ... < 1 > ... > 0 triangle with n points on each flank (the reference's concatenated linspaces)."""
    up = torch.linspace(0, 1, n + 2)[:-1]
    down = torch.linspace(1, 0, n + 2)
    return torch.cat([up, down])[1:-1]


class SpectralGate(nn.Module):
    def __init__(self, n_freq: int = 3, n_time: int = 5):
        super().__init__()
        kernel = torch.outer(_ramp(n_freq), _ramp(n_time))
        self.register_buffer("smoothing_filter", (kernel / kernel.sum())[None, None])

    def forward(self, audio_signal: AudioSignal, nz_signal: AudioSignal, denoise_amount: float = 1.0,
                n_std: float = 3.0, win_length: int = 2048, hop_length: int = 512):
        stft_params = STFTParams(win_length, hop_length, "sqrt_hann")
        audio_signal = audio_signal.clone()
        audio_signal.stft_data = None
        audio_signal.stft_params = stft_params
        nz_signal = nz_signal.clone()
        nz_signal.stft_params = stft_params

        nz_db = 20 * nz_signal.magnitude.clamp(1e-4).log10()
        thresh = nz_db.mean(keepdim=True, dim=-1) + nz_db.std(keepdim=True, dim=-1) * n_std  # per bin
        sig_db = 20 * audio_signal.magnitude.clamp(1e-4).log10()
        nb, nac, nf, nt = sig_db.shape
        mask = (sig_db < thresh.expand(nb, nac, -1, nt)).float()
        kf, kt = self.smoothing_filter.shape[-2:]
        mask = torch.nn.functional.conv2d(mask.reshape(nb * nac, 1, nf, nt), self.smoothing_filter.to(mask.device),
                                          padding=(kf // 2, kt // 2)).reshape(nb, nac, nf, nt)
        mask = 1 - mask * util.ensure_tensor(denoise_amount, ndim=mask.ndim).to(mask.device)
        audio_signal.stft_data = audio_signal.stft_data * mask
        audio_signal.istft()
        return audio_signal

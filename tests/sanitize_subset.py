"""Small driver for compute-sanitizer runs (memcheck / racecheck / synccheck): exercises every kernel of
libb2a on small shapes through the public API.  `compute-sanitizer --tool racecheck python tests/sanitize_subset.py`"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

graft.build()
from audiotools_b200 import AudioSignal  # noqa: E402

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = 0.1 * torch.randn(3, 2, 30000, generator=g)
sig = AudioSignal(x.clone(), 44100).to(dev)
sig.normalize(-24.0)
lm = sig.mel_spectrogram(n_mels=128, window_length=2048, hop_length=512, log=True)
_ = sig.stft(window_length=512, hop_length=128)
_ = AudioSignal(x.clone(), 44100).to(dev).stft(window_length=4096, hop_length=1024)
_ = AudioSignal(x[..., :9000].clone(), 16000).to(dev).mel_spectrogram(n_mels=80)
y = AudioSignal(x.clone(), 44100).to(dev).low_pass(4000).high_pass(100).equalizer(-np.random.RandomState(0).rand(6))
ir = torch.randn(3, 1, 4000, generator=g) * torch.exp(-torch.arange(4000) / 800.0)
ir[..., 10] = 2.0
y = y.convolve(AudioSignal(ir, 44100).to(dev))
y = y.resample(16000).pitch_shift(2)          # 44.1k -> 16k: general polyphase kernel; WSOLA search / OLA / rate kernels
z = AudioSignal(x.clone(), 48000).to(dev).resample(16000).low_pass(7000)  # fir.cu: single-phase decimator + 103-tap FIR
for wl, hop in ((2048, 512), (512, 100), (64, 16)):  # istft.cu: inverse FFT + gather overlap-add, three plans
    s2 = AudioSignal(x.clone(), 44100).to(dev)
    s2.stft(window_length=wl, hop_length=hop)
    s2.istft(window_length=wl, hop_length=hop)
s3 = AudioSignal(x.clone(), 44100, stft_params=None).to(dev)
s3.stft(window_length=256, hop_length=64, window_type="sqrt_hann", match_stride=True)
s3.istft(window_length=256, hop_length=64, window_type="sqrt_hann", match_stride=True)
torch.cuda.synchronize()
print("ok", float(lm.mean()), float(y.audio_data.abs().mean()))

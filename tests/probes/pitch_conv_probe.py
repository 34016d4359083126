"""Probe for ncu launch lists of the secondary kernels (pitch shift, apply_ir) at the cfg4 per-GPU share."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as graft
graft.build()
from audiotools_b200 import AudioSignal
g = torch.Generator().manual_seed(0)
B, T, sr = 128, 441000, 44100
x = (0.1 * torch.randn(B, 1, T, generator=g)).cuda()
t = torch.arange(sr) / sr
ir = (torch.randn(B, 1, sr, generator=g) * torch.exp(-t / 0.3) * 0.1)
ir[..., 40] = 1.0
ir = ir.cuda()
for _ in range(2):
    y = AudioSignal(x, sr).pitch_shift(2)
    z = AudioSignal(x, sr).convolve(AudioSignal(ir.clone(), sr))
    w = AudioSignal(x, sr).equalizer(-torch.rand(B, 6))
    v = AudioSignal(x[:, :, :].reshape(B, 1, T), 48000).resample(16000).low_pass(8000)
torch.cuda.synchronize()
print("ok")

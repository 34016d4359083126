"""Section timing of the SpecAug chain at cfg2's shape (CUDA events)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as graft
graft.build()
from audiotools_b200 import AudioSignal

def timed(fn, warm=2, steps=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

g = torch.Generator().manual_seed(0)
B = 64
x = (0.1 * torch.randn(B, 2, 441000, generator=g)).cuda()
fmin = (torch.rand(B, generator=g) * 8000).cuda(); fmax = fmin + 2000
tmin = (torch.rand(B, generator=g) * 9).cuda(); tmax = tmin + 0.25
s = AudioSignal(x, 44100)
print("stft (materialised)  %.3f ms" % timed(lambda: s.stft(window_length=2048, hop_length=512)))
print("mask_frequencies     %.3f ms" % timed(lambda: s.mask_frequencies(fmin, fmax)))
print("mask_timesteps       %.3f ms" % timed(lambda: s.mask_timesteps(tmin, tmax)))
print("shift_phase(scalar)  %.3f ms" % timed(lambda: s.shift_phase(0.3)))
print("mask_low_magnitudes  %.3f ms" % timed(lambda: s.mask_low_magnitudes(-20.0)))
print("istft                %.3f ms" % timed(lambda: s.istft(window_length=2048, hop_length=512)))
print("mel(log) no stft out %.3f ms" % timed(lambda: AudioSignal(x, 44100).mel_spectrogram(n_mels=128, window_length=2048, hop_length=512, log=True)))

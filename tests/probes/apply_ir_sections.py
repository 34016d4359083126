"""Section timing of apply_ir / equalizer / pitch_shift at the cfg4 per-GPU share (CUDA events)."""
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
B, T, sr = 128, 441000, 44100
x = (0.1 * torch.randn(B, 1, T, generator=g)).cuda()
t = torch.arange(sr) / sr
ir = (torch.randn(B, 1, sr, generator=g) * torch.exp(-t / 0.3) * 0.1); ir[..., 40] = 1.0
ir = ir.cuda()
eq = -torch.rand(B, 6).cuda()
drr = (torch.rand(B) * 30).cuda()
sig = AudioSignal(x, sr); irs = AudioSignal(ir, sr)
print("ir.equalizer      %.3f ms" % timed(lambda: irs.clone().equalizer(eq)))
print("ir.alter_drr      %.3f ms" % timed(lambda: irs.clone().alter_drr(drr)))
print("abs().max         %.3f ms" % timed(lambda: sig.audio_data.abs().max(dim=-1, keepdims=True).values))
print("convolve          %.3f ms" % timed(lambda: sig.clone().convolve(irs)))
print("clone             %.3f ms" % timed(lambda: sig.clone()))
print("apply_ir(full)    %.3f ms" % timed(lambda: sig.clone().apply_ir(irs.clone(), drr, eq)))
print("sig.equalizer     %.3f ms" % timed(lambda: sig.clone().equalizer(eq)))
print("pitch_shift(+2)   %.3f ms" % timed(lambda: sig.clone().pitch_shift(2)))
print("pitch_shift(-2)   %.3f ms" % timed(lambda: sig.clone().pitch_shift(-2)))

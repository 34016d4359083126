"""istft at cfg2's shape -- target of `ncu --set full -k regex:istft_kernel`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as graft
graft.build()
from audiotools_b200 import AudioSignal
g = torch.Generator().manual_seed(0)
x = (0.1 * torch.randn(64, 2, 441000, generator=g)).cuda()
sig = AudioSignal(x, 44100)
sig.stft(window_length=2048, hop_length=512)
for _ in range(3):
    sig.istft(window_length=2048, hop_length=512)
torch.cuda.synchronize()
print("ok")

"""pitch_shift only (cfg4 per-GPU share) -- target of `ncu --set full -k regex:wsola_search|rate_kernel`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as graft
graft.build()
from audiotools_b200 import AudioSignal
g = torch.Generator().manual_seed(0)
B, T, sr = 128, 441000, 44100
x = (0.1 * torch.randn(B, 1, T, generator=g)).cuda()
for _ in range(2):
    y = AudioSignal(x, sr).pitch_shift(2)
torch.cuda.synchronize()
print("ok")

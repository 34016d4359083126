"""Seeded inputs shared by ``make_golden.py`` (which runs the real reference on
them here) and the tests (which re-generate them on any box).  torch's CPU
generator is deterministic for a given torch build; the GPU box runs the same
image, and ``reference_golden.npz`` stores a checksum of every input so a
mismatch is detected instead of silently comparing different signals."""
import numpy as np
import torch

# name -> (seed, B, C, T, scale, per_item_gains, sample_rate)
CASES = {
    "cfg1": (0, 4, 1, 16000, 1.0, False, 16000),     # BASELINE.json configs[0]
    "cfg2": (1, 3, 2, 66150, 0.1, True, 44100),      # configs[1] shape, 1.5 s
    "lufs16k": (0, 16, 2, 16000, 1.0, True, 16000),  # ref:tests/core/test_loudness.py:31-52
    "short": (2, 2, 1, 4000, 0.3, True, 16000),      # 0.25 s -> padded to 0.5 s
    "lufs48k": (3, 2, 2, 30000, 0.2, True, 48000),
    "lufs11k": (4, 2, 1, 22050, 0.2, True, 11025),   # K=4410, stride=1102: K != 4*stride
    "rs": (5, 2, 1, 24000, 0.1, True, 48000),
    "fir": (6, 3, 2, 12000, 0.1, True, 44100),
    "tfm": (8, 4, 1, 12000, 0.1, True, 44100),
}


def make_input(name: str) -> torch.Tensor:
    seed, B, C, T, scale, gains, _ = CASES[name]
    g = torch.Generator().manual_seed(seed)
    x = scale * torch.randn(B, C, T, generator=g)
    if name != "cfg1":
        x = x.clamp(-1, 1)
    if gains:
        x = x * (0.05 + 0.95 * torch.rand(B, 1, 1, generator=g))
    x = x.float()
    if name == "cfg2":
        x[1] = 0.0  # an all-silent item (-> NaN scrub, -70 clamp)
    return x


def sample_rate(name: str) -> int:
    return CASES[name][6]


def make_ir() -> torch.Tensor:
    """Synthetic room IR (stands in for the LFS audio): decaying noise + direct-path spike
    with 37 samples of pre-delay.  Mono: the reference's ``convolve`` only supports 1-channel
    IRs when ``start_at_max=True`` (``idx[i].item()`` at ref:audiotools/core/effects.py:99)."""
    g = torch.Generator().manual_seed(7)
    t = torch.arange(4000) / 44100
    ir = torch.randn(3, 1, 4000, generator=g) * torch.exp(-t / 0.02) * 0.3
    ir[..., 37] = 1.0
    return ir.float()


def make_eq() -> np.ndarray:
    """ref:audiotools/data/transforms.py:594-597 with eq_amount=1."""
    return -1.0 * np.stack([np.random.RandomState(i).rand(6) for i in range(3)])


def checksum(x: torch.Tensor) -> float:
    return float(x.double().abs().sum())

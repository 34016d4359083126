"""Golden vectors for pitch_shift / time_stretch: outputs of the INDEPENDENT specification oracle
(oracle/pitch_spec.py -- numpy, float64, written from the prose specification) on seeded inputs.

The reference's own implementation is libsox (not under /root/reference, output never pinned by the reference's tests),
so these vectors pin the written specification, not SoX; they exist so that (a) the oracle cannot drift unnoticed
(tests/test_oracle_golden.py re-derives them) and (b) the GPU tests compare the CUDA path with stored numbers.

    python tests/golden/make_golden_pitch.py     ->  tests/golden/pitch_golden.npz
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from oracle import pitch_spec as ps  # noqa: E402

SR = 16000
T = 24000
SHIFTS = (2.0, -2.0, 7.0, -0.5)
FACTORS = (1.25, 0.8)


def make_input():
    """[3, T] float32: tone + noise, pure noise, speech-like AM harmonic stack."""
    rng = np.random.RandomState(20260923)
    t = np.arange(T) / SR
    a = 0.3 * np.sin(2 * np.pi * 220 * t) + 0.05 * rng.randn(T)
    b = 0.1 * rng.randn(T)
    c = sum(0.2 / k * np.sin(2 * np.pi * 140 * k * t + k) for k in range(1, 9)) * (0.6 + 0.4 * np.sin(2 * np.pi * 3 * t))
    return np.stack([a, b, c]).astype(np.float32)


def main():
    x = make_input()
    out = {"x": x, "sr": np.int64(SR)}
    for st in SHIFTS:
        ys, ps_, ms = zip(*(ps.pitch_shift_row(r, SR, st) for r in x))
        out[f"pitch_{st:g}_y"] = np.stack(ys).astype(np.float32)
        out[f"pitch_{st:g}_pos"] = np.stack(ps_).astype(np.int32)
        out[f"pitch_{st:g}_margin"] = np.stack(ms)
    for fac in FACTORS:
        ys, ps_, ms = zip(*(ps.time_stretch_row(r, SR, fac) for r in x))
        out[f"stretch_{fac:g}_y"] = np.stack(ys).astype(np.float32)
        out[f"stretch_{fac:g}_pos"] = np.stack(ps_).astype(np.int32)
        out[f"stretch_{fac:g}_margin"] = np.stack(ms)
    path = os.path.join(REPO, "tests", "golden", "pitch_golden.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()

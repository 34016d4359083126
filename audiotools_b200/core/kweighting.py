"""K-weighting biquad design (float64, host).

The reference obtains these coefficients from ``pyloudnorm.Meter(rate)._filters``
(ref:audiotools/core/loudness.py:253-260): an RBJ-style high-shelf (+4 dB, Q 1/sqrt2,
1500 Hz) followed by a high-pass (Q 0.5, 38 Hz), designed at the signal's rate, each
with passband gain 1.0.  At 48 kHz they agree with the ITU-R BS.1770-4 table to 4e-5.
Design is host-side scalar math (6 numbers per stage); the filtering itself is
``libb2a`` (``csrc/lufs.cu``).
"""
import math
from collections import OrderedDict

import numpy as np

# (G dB, Q, fc Hz, type) in application order -- the shelf first.
_CLASSES = {
    "K-weighting": OrderedDict(
        high_shelf=(4.0, 1.0 / math.sqrt(2.0), 1500.0, "high_shelf"),
        high_pass=(0.0, 0.5, 38.0, "high_pass"),
    ),
}

CHANNEL_GAINS = np.array([1.0, 1.0, 1.0, 1.41, 1.41])  # ref:audiotools/core/loudness.py:49-50


def biquad(G: float, Q: float, fc: float, rate: float, kind: str):
    """Returns (b[3], a[3]) normalised by a0, float64."""
    A = 10.0 ** (G / 40.0)
    w0 = 2.0 * math.pi * (fc / rate)
    alpha = math.sin(w0) / (2.0 * Q)
    cw = math.cos(w0)
    if kind == "high_shelf":
        sA = math.sqrt(A)
        b0 = A * ((A + 1) + (A - 1) * cw + 2 * sA * alpha)
        b1 = -2 * A * ((A - 1) + (A + 1) * cw)
        b2 = A * ((A + 1) + (A - 1) * cw - 2 * sA * alpha)
        a0 = (A + 1) - (A - 1) * cw + 2 * sA * alpha
        a1 = 2 * ((A - 1) - (A + 1) * cw)
        a2 = (A + 1) - (A - 1) * cw - 2 * sA * alpha
    elif kind == "high_pass":
        b0 = (1 + cw) / 2
        b1 = -(1 + cw)
        b2 = (1 + cw) / 2
        a0 = 1 + alpha
        a1 = -2 * cw
        a2 = 1 - alpha
    else:
        raise NotImplementedError(f"biquad type {kind!r}")
    return np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0


def design(rate: float, filter_class: str = "K-weighting"):
    """-> (sos[n_stage, 6] = b0 b1 b2 a0 a1 a2, passband_gain[n_stage]) float64."""
    if filter_class not in _CLASSES:
        raise NotImplementedError(
            f"filter_class {filter_class!r}: only 'K-weighting' is implemented "
            "(the other pyloudnorm classes named in ref:audiotools/core/loudness.py:18-22 are not)")
    rows, gains = [], []
    for G, Q, fc, kind in _CLASSES[filter_class].values():
        b, a = biquad(G, Q, fc, rate, kind)
        rows.append(np.concatenate([b, a]))
        gains.append(1.0)
    return np.ascontiguousarray(np.stack(rows), dtype=np.float64), np.asarray(gains, dtype=np.float64)

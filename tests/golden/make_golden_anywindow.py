"""Golden vectors of the REAL reference for window lengths that are not powers of two (the dense-DFT path of
csrc/dft.cu), produced exactly like ``make_golden.py`` (same shims; run here only):
``python tests/golden/make_golden_anywindow.py`` -> ``reference_golden_anywindow.npz``
(ref:audiotools/core/audio_signal.py:1123-1212 stft, :1214-1296 istft, :1333-1369 mel_spectrogram, :1398-1426 mfcc)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from tests.golden.make_golden import import_reference  # noqa: E402
from tests.golden.cases import make_input  # noqa: E402

# (key, window_length, hop_length, window_type, match_stride, padding_type)
STFT_CASES = [
    ("w400", 400, 160, "hann", False, "reflect"),        # 25 ms / 10 ms speech front-end at 16 kHz
    ("w480", 480, 120, "sqrt_hann", False, "reflect"),
    ("w400_ms", 400, 100, "hann", True, "reflect"),      # match_stride: explicit reflect pad + dropped edge frames
    ("w400_ms_const", 400, 100, "hann", True, "constant"),
    ("w201", 201, 50, "hamming", False, "replicate"),    # odd length: torch loses one sample in the frame count
    ("w1200", 1200, 300, "hann", False, "reflect"),
]


def main():
    at = import_reference()
    AudioSignal = at.AudioSignal
    x = make_input("cfg1")  # [4, 1, 16000] @ 16 kHz
    out = {}
    for key, wl, hop, wt, ms, pt in STFT_CASES:
        s = AudioSignal(x.clone(), 16000)
        X = s.stft(window_length=wl, hop_length=hop, window_type=wt, match_stride=ms, padding_type=pt)
        out[key + "_stft"] = X[:2].numpy()  # (first two items: fixture size)
        y = s.istft(window_length=wl, hop_length=hop, window_type=wt, match_stride=ms)
        out[key + "_istft"] = y.audio_data.numpy()
    s = AudioSignal(x.clone(), 16000)
    out["w400_mel80"] = s.mel_spectrogram(n_mels=80, window_length=400, hop_length=160, window_type="hann").numpy()
    s = AudioSignal(x.clone(), 16000)
    out["w400_mfcc"] = s.mfcc(n_mfcc=20, n_mels=40, window_length=400, hop_length=160, window_type="hann").numpy()
    path = os.path.join(HERE, "reference_golden_anywindow.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

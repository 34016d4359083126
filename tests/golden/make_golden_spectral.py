"""Golden vectors of the SpectralTransform family, produced by the REAL reference exactly like
``make_golden.py`` (same shims; run here only):  ``python tests/golden/make_golden_spectral.py``
-> ``reference_golden_spectral.npz`` (ref:audiotools/core/dsp.py:217-370, ref:audiotools/data/transforms.py:1200-1453)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from tests.golden.make_golden import _flatten, import_reference  # noqa: E402
from tests.golden.cases import make_input  # noqa: E402

# per-item parameters shared with the tests
FMIN = torch.tensor([500.0, 0.0, 3000.0, 7000.0])
FMAX = torch.tensor([1500.0, 250.0, 3100.0, 8000.0])
TMIN = torch.tensor([0.10, 0.0, 0.50, 0.90])
TMAX = torch.tensor([0.20, 0.05, 0.51, 1.00])
DBCUT = torch.tensor([-10.0, 0.0, 5.0, -40.0])
SHIFT = torch.tensor([0.5, -1.0, float(np.pi), 0.0])
SEEDS = [20, 21, 22, 23]


def main():
    at = import_reference()
    AudioSignal = at.AudioSignal
    from audiotools.data import transforms as tfm

    out = {}
    x = make_input("cfg1")  # [4, 1, 16000] @ 16 kHz, default STFT 512/128 hann

    def fresh():
        s = AudioSignal(x.clone(), 16000)
        s.stft()
        return s

    s = fresh().mask_frequencies(FMIN, FMAX)
    out["maskfreq_stft"] = s.stft_data.numpy()
    out["maskfreq_audio"] = s.istft().audio_data.numpy()
    s = fresh().mask_frequencies(FMIN, FMAX, val=0.25)
    out["maskfreq_val_stft"] = s.stft_data[:1].numpy()
    s = fresh().mask_timesteps(TMIN, TMAX)
    out["masktime_stft"] = s.stft_data.numpy()
    out["masktime_audio"] = s.istft().audio_data.numpy()
    s = fresh().mask_low_magnitudes(DBCUT)
    out["masklow_stft"] = s.stft_data[:2].numpy()
    out["masklow_audio"] = s.istft().audio_data.numpy()
    s = fresh().shift_phase(SHIFT)
    out["shift_stft"] = s.stft_data[2:].numpy()
    out["shift_audio"] = s.istft().audio_data.numpy()
    g = torch.Generator().manual_seed(77)
    corr = 0.3 * torch.randn(4, 1, 257, 126, generator=g)
    out["corrupt_in"] = corr.numpy()
    out["corrupt_audio"] = fresh().shift_phase(corr).istft().audio_data.numpy()

    # transforms (seeded instantiate; TimeNoise / FrequencyNoise draw device noise: properties only, not pinned here)
    transform = tfm.Compose([tfm.FrequencyMask(), tfm.TimeMask(prob=0.7), tfm.ShiftPhase(), tfm.MaskLowMagnitudes(prob=0.6),
                             tfm.CorruptPhase(prob=0.5), tfm.InvertPhase(prob=0.5)])
    sig = AudioSignal(x.clone(), 16000)
    kwargs = transform.batch_instantiate(SEEDS, sig)
    for k, v in _flatten(kwargs).items():
        out["kw/" + "/".join(k)] = v.numpy()
    out["compose_audio"] = transform(sig.clone(), **kwargs).audio_data.numpy()
    sm = tfm.Smoothing()
    kw = sm.batch_instantiate(SEEDS, sig)
    win = kw["Smoothing"]["window"]
    out["smooth_window"] = win.audio_data.numpy()
    out["smooth_audio"] = sm(sig.clone(), **kw).audio_data.numpy()

    # chunking helpers (ref:audiotools/core/dsp.py:15-151): collect_windows -> (a stand-in for a model) -> overlap_and_add
    x2 = torch.cat([x[:1], 0.5 * x[:1].flip(-1)], 1)  # one stereo item
    sw = AudioSignal(x2.clone(), 16000).collect_windows(0.1, 0.05)
    out["win_collect"] = sw.audio_data.numpy()
    sw.audio_data = sw.audio_data * 0.5 + 0.1
    out["win_ola"] = sw.overlap_and_add(0.05).audio_data.numpy()
    out["win_iter_count"] = np.int64(sum(1 for _ in AudioSignal(x2.clone(), 16000).windows(0.064, 0.016)))

    # elementwise effects and their transforms (ref:audiotools/core/effects.py:435-523, transforms.py:531-666,1173-1197);
    # clip_distortion only works on mono input in the reference (its quantile indexing)
    xs = x * 0.3
    xs2 = torch.cat([xs, 0.5 * xs.flip(-1)], 1)
    out["fx_clip"] = AudioSignal(xs.clone(), 16000).clip_distortion(torch.tensor([0.05, 0.2, 0.0, 0.5])).audio_data.numpy()
    out["fx_quant"] = AudioSignal(xs2.clone(), 16000).quantization(torch.tensor([8, 16, 256, 3])).audio_data.numpy()
    out["fx_mulaw"] = AudioSignal(xs2.clone(), 16000).mulaw_quantization(torch.tensor([8, 16, 256, 3])).audio_data.numpy()
    out["fx_maxaudio"] = AudioSignal(xs2.clone() * 5, 16000).ensure_max_of_audio(0.7).audio_data.numpy()
    for cls, kw, inp in (("ClippingDistortion", {}, xs), ("Quantization", {}, xs2), ("MuLawQuantization", {}, xs2),
                         ("RescaleAudio", {"val": 0.5}, xs2)):
        t = getattr(tfm, cls)(**kw)
        s_in = AudioSignal(inp.clone() * 3, 16000)
        k = t.batch_instantiate([3, 4, 5, 6], s_in)
        for kk, v in _flatten(k).items():
            out[f"fxkw/{cls}/" + "/".join(kk)] = v.numpy()
        out[f"fxout/{cls}"] = t(s_in.clone(), **k).audio_data.numpy()
    sg = AudioSignal(xs2.clone(), 16000)
    sg.metadata["loudness"] = -30.0
    out["fx_globalvolnorm_db"] = tfm.GlobalVolumeNorm(db=("uniform", -20, -10)).instantiate(7, sg)["GlobalVolumeNorm"]["db"].numpy()

    # SpectralGate / SpectralDenoising (ref:audiotools/ml/layers/spectral_gate.py, transforms.py:1539-1592)
    from audiotools.ml.layers import SpectralGate

    xg = make_input("cfg2")[:2, :, :30000]
    nz = 0.01 * torch.randn(2, 1, 22050, generator=torch.Generator().manual_seed(0))
    out["gate_nz"] = nz.numpy()
    out["gate_out"] = SpectralGate()(AudioSignal(xg.clone(), 44100), AudioSignal(nz.clone(), 44100),
                                     torch.tensor([0.9, 0.8])).audio_data.numpy()
    sd = tfm.SpectralDenoising()
    sg_in = AudioSignal(xg.clone(), 44100)
    k = sd.batch_instantiate([3, 4], sg_in)
    for kk, v in _flatten(k).items():
        out["sdkw/" + "/".join(kk)] = (v.audio_data if hasattr(v, "audio_data") else v).numpy()
    out["sd_out"] = sd(sg_in.clone(), **k).audio_data.numpy()

    path = os.path.join(HERE, "reference_golden_spectral.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)/1e6:.1f} MB on disk")


if __name__ == "__main__":
    main()

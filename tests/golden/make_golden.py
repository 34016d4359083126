"""Generate golden vectors by running the REAL reference (``/root/reference``).

Run here (build container) only:  ``python tests/golden/make_golden.py``.
``/root/reference`` does not exist on the GPU box, so the outputs are committed
as small ``.npz`` fixtures next to this script and the tests read only those.

The reference cannot be imported as-is: ``flatten_dict``,
``julius``, ``pyloudnorm``, ``librosa``, ``soundfile``, ``matplotlib``, ... are
not installed and there is no network.  This script therefore installs import
shims:

* ``julius`` / ``pyloudnorm`` / ``librosa.filters``: backed by the restated
  arithmetic in ``oracle/third_party.py``  (third-party code; its source is not
  in /root/reference -- what the golden vectors pin is the reference's OWN code
  in ``audiotools/core/*.py`` and ``audiotools/data/transforms.py`` running on
  top of that arithmetic);
* ``flatten_dict``: 12-line nested-dict flatten/unflatten with tuple keys;
* everything else that is absent and irrelevant to the hot path (plotting,
  file IO, UI): inert dummy modules.

Everything executed between the shims is the unmodified reference.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import third_party as tp  # noqa: E402


# ---------------------------------------------------------------------------
# shims
# ---------------------------------------------------------------------------
class _Dummy(types.ModuleType):
    """Inert module: any attribute is another dummy / a no-op callable class."""

    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None,
                               "__call__": lambda self, *a, **k: None})


class _DummyFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("soundfile", "matplotlib", "IPython", "randomname", "markdown2", "ffmpy",
                "gradio", "argbind", "pystoi", "torch_stoi",
                "tensorboard", "torch.utils.tensorboard", "pesq", "visqol", "librosa.display",
                "whisper", "rich")

    def find_spec(self, fullname, path, target=None):
        if any(fullname == p or fullname.startswith(p + ".") for p in self.PREFIXES):
            try:  # prefer the real module when it is installed
                sys.meta_path.remove(self)
                try:
                    spec = importlib.util.find_spec(fullname)
                finally:
                    sys.meta_path.insert(0, self)
                if spec is not None:
                    return None
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Dummy(spec.name)

    def exec_module(self, module):
        pass


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _flatten(d, parent=()):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict) and v:
            out.update(_flatten(v, parent + (k,)))
        else:
            out[parent + (k,)] = v
    return out


def _unflatten(d):
    out = {}
    for path, v in d.items():
        cur = out
        for k in path[:-1]:
            cur = cur.setdefault(k, {})
        cur[path[-1]] = v
    return out


def install_shims():
    import importlib.util  # noqa: F401

    import importlib.resources

    sys.meta_path.insert(0, _DummyFinder())
    sys.modules.setdefault("importlib_resources", importlib.resources)  # stdlib twin of the backport
    _module("flatten_dict", flatten=_flatten, unflatten=_unflatten)

    class _AsModule(torch.nn.Module):
        """julius filters are nn.Modules in the original (``.float().to(device)`` is called)."""

        def __init__(self, impl):
            super().__init__()
            self._impl = impl

        def forward(self, x):
            return self._impl(x)

    julius = _module(
        "julius",
        resample_frac=tp.resample_frac,
        LowPassFilter=lambda cutoff, zeros=8: _AsModule(tp.LowPassFilter(cutoff, zeros)),
        HighPassFilter=lambda cutoff, zeros=8: _AsModule(tp.HighPassFilter(cutoff, zeros)),
        SplitBands=lambda sr, n_bands=None: _AsModule(tp.SplitBands(sr, n_bands)),
    )
    julius.fftconv = _module("julius.fftconv", fft_conv1d=tp.fft_conv1d)
    julius.core = _module("julius.core", unfold=tp.unfold)
    _module("pyloudnorm", Meter=tp.PyloudnormMeterShim)
    librosa = _module("librosa")
    librosa.__path__ = []
    librosa.filters = _module("librosa.filters", mel=tp.librosa_mel)
    librosa.load = None


def import_reference():
    install_shims()
    sys.path.insert(0, "/root/reference")
    import audiotools  # noqa: F401

    assert audiotools.__file__.startswith("/root/reference"), audiotools.__file__
    return audiotools


# ---------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------
from tests.golden.cases import CASES, checksum, make_eq, make_input, make_ir  # noqa: E402


def main():
    at = import_reference()
    AudioSignal, STFTParams = at.AudioSignal, at.STFTParams
    from audiotools.data import transforms as tfm
    from audiotools.core.loudness import Meter

    out = {}
    for name in CASES:
        out[f"checksum/{name}"] = np.float64(checksum(make_input(name)))

    # -- cfg1: batch=4 mono 1s@16k stft(n_fft=512, hop=128)  (BASELINE.json configs[0])
    x = make_input("cfg1")
    sig = AudioSignal(x.clone(), 16000)
    out["cfg1_stft"] = sig.stft(window_length=512, hop_length=128).numpy()
    # the signal's defaults @16k are (512, 128, "hann", False, "reflect"): must be the same thing
    assert np.array_equal(AudioSignal(x.clone(), 16000).stft().numpy(), out["cfg1_stft"])
    x2b = x[:2]
    sig = AudioSignal(x2b.clone(), 16000, stft_params=STFTParams(256, 64, "sqrt_hann", True, "reflect"))
    out["cfg1_stft_match_stride"] = sig.stft().numpy()
    out["cfg1_istft_match_stride"] = sig.istft().audio_data.numpy()
    sig = AudioSignal(x2b[..., :15999].clone(), 16000, stft_params=STFTParams(256, 64, "hann", True, "reflect"))
    out["cfg1_stft_match_stride_odd"] = sig.stft().numpy()
    sig = AudioSignal(x2b.clone(), 16000)
    out["cfg1_stft_average_hop100"] = sig.stft(window_length=256, hop_length=100,
                                               window_type="average").numpy()
    out["cfg1_mel80"] = AudioSignal(x.clone(), 16000).mel_spectrogram(n_mels=80).numpy()
    out["cfg1_mel40_fmin_fmax"] = AudioSignal(x2b.clone(), 16000).mel_spectrogram(
        n_mels=40, mel_fmin=100.0, mel_fmax=6000.0, window_length=1024, hop_length=256).numpy()
    out["cfg1_mfcc"] = AudioSignal(x.clone(), 16000).mfcc().numpy()
    sig = AudioSignal(x2b.clone(), 16000)
    sig.stft()
    out["cfg1_logmag"] = sig.log_magnitude().numpy()
    out["cfg1_istft"] = sig.istft().audio_data.numpy()

    # -- cfg2-shaped (small): stereo 44.1k, 1.5 s, log-mel 2048/512/128 + LUFS normalize
    x2 = make_input("cfg2")
    out["cfg2_lufs"] = AudioSignal(x2.clone(), 44100).loudness().numpy()
    sig = AudioSignal(x2.clone(), 44100)
    sig.normalize(-24.0)
    out["cfg2_norm"] = sig.audio_data.numpy()
    mel = sig.mel_spectrogram(n_mels=128, window_length=2048, hop_length=512, window_type="hann")
    out["cfg2_mel"] = mel.numpy()
    out["cfg2_logmel"] = mel.clamp(1e-5).pow(2).log10().numpy()  # ref:audiotools/metrics/spectral.py:187-190
    # block energies straight from the reference Meter (pre-gating), and the FIR-path LUFS
    m = Meter(44100)
    filt = m.apply_filter(x2.clone().permute(0, 2, 1))
    out["cfg2_z"] = ((1.0 / (0.4 * 44100)) * m._unfold(filt).square().sum(2)).numpy()
    out["cfg2_lufs_fir_mono"] = AudioSignal(x2[:, :1].clone(), 44100).loudness(use_fir=True).numpy()

    # -- loudness edge cases (mirrors ref:tests/core/test_loudness.py:31-52)
    x3 = make_input("lufs16k")
    out["lufs16k"] = AudioSignal(x3.clone(), 16000).loudness().numpy()
    for i in range(x3.shape[0]):  # batch == per-item, the reference's own assertion
        li = AudioSignal(x3[i:i + 1].clone(), 16000).loudness().numpy()
        assert np.allclose(li, out["lufs16k"][i], atol=1e-4)
    out["lufs_short"] = AudioSignal(make_input("short"), 16000).loudness().numpy()
    out["lufs48k"] = AudioSignal(make_input("lufs48k"), 48000).loudness().numpy()
    out["lufs11k"] = AudioSignal(make_input("lufs11k"), 11025).loudness().numpy()
    db = torch.linspace(-40, -10, 4)
    out["norm16k_db"] = db.numpy()
    out["norm16k"] = AudioSignal(x3[:4].clone(), 16000).normalize(db).audio_data.numpy()
    out["volchange16k"] = AudioSignal(x3[:4].clone(), 16000).volume_change(db).audio_data.numpy()

    # -- resample (cfg3-shaped, small) + low/high-pass
    x6 = make_input("rs")
    out["rs_48k_16k"] = AudioSignal(x6.clone(), 48000).resample(16000).audio_data.numpy()
    out["rs_44k_16k"] = AudioSignal(x6[..., :22050].clone(), 44100).resample(16000).audio_data.numpy()
    out["rs_16k_44k"] = AudioSignal(x6[..., :8000].clone(), 16000).resample(44100).audio_data.numpy()
    out["rs_16k_48k"] = AudioSignal(x6[..., :8001].clone(), 16000).resample(48000).audio_data.numpy()
    out["rs_44k_48k"] = AudioSignal(x6[..., :4410].clone(), 44100).resample(48000).audio_data.numpy()
    out["rs_48k_16k_lp8k"] = AudioSignal(x6.clone(), 48000).resample(16000).low_pass(8000).audio_data.numpy()
    x7 = make_input("fir")
    cut = torch.tensor([4000.0, 8000.0, 1000.0])
    out["fir_cut"] = cut.numpy()
    out["lp_peritem"] = AudioSignal(x7.clone(), 44100).low_pass(cut).audio_data.numpy()
    out["hp_peritem"] = AudioSignal(x7.clone(), 44100).high_pass(cut / 8).audio_data.numpy()
    out["lp_scalar"] = AudioSignal(x7.clone(), 44100).low_pass(4000).audio_data.numpy()

    # -- equalizer / mel_filterbank
    eq = make_eq()
    out["eq_db"] = eq
    out["eq_out"] = AudioSignal(x7.clone(), 44100).equalizer(eq).audio_data.numpy()
    out["eq_out_1d"] = AudioSignal(x7.clone(), 44100).equalizer(eq[0]).audio_data.numpy()
    out["fbank4"] = AudioSignal(x7[:1, :1].clone(), 44100).mel_filterbank(4).numpy()

    # -- convolve / apply_ir
    ir = make_ir()
    out["conv_out"] = AudioSignal(x7.clone(), 44100).convolve(AudioSignal(ir.clone(), 44100)).audio_data.numpy()
    out["conv_out_nomax"] = AudioSignal(x7.clone(), 44100).convolve(
        AudioSignal(ir.clone(), 44100), start_at_max=False).audio_data.numpy()
    out["applyir_plain"] = AudioSignal(x7.clone(), 44100).apply_ir(AudioSignal(ir.clone(), 44100)).audio_data.numpy()
    drr = torch.tensor([5.0, 15.0, 25.0])
    out["drr"] = drr.numpy()
    out["applyir_full"] = AudioSignal(x7.clone(), 44100).apply_ir(
        AudioSignal(ir.clone(), 44100), drr=drr, ir_eq=eq).audio_data.numpy()
    out["alter_drr"] = AudioSignal(ir.clone(), 44100).alter_drr(drr).audio_data.numpy()
    out["measure_drr"] = AudioSignal(ir.clone(), 44100).measure_drr().numpy()

    # -- transforms: Compose semantics with masks (seeded instantiate)
    x8 = make_input("tfm")
    transform = tfm.Compose(
        [tfm.VolumeNorm(db=("uniform", -30, -16)), tfm.Equalizer(prob=0.5), tfm.LowPass(prob=0.7),
         tfm.HighPass(prob=0.6), tfm.VolumeChange()],
    )
    sig = AudioSignal(x8.clone(), 44100)
    kwargs = transform.batch_instantiate([10, 11, 12, 13], sig)
    for k, v in _flatten(kwargs).items():
        out["tfm_kw/" + "/".join(k)] = v.numpy()
    out["tfm_out"] = transform(sig.clone(), **kwargs).audio_data.numpy()

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    total = sum(np.asarray(v).nbytes for v in out.values())
    print(f"wrote {path}: {len(out)} arrays, {total/1e6:.1f} MB raw, "
          f"{os.path.getsize(path)/1e6:.1f} MB on disk")


if __name__ == "__main__":
    main()

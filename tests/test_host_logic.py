"""Host-side logic of the drop-in layer on CPU tensors (no kernels involved): the container
semantics of AudioSignal and the Compose / Choose / Repeat / mask algebra of the transforms,
modelled on ref:tests/data/test_transforms.py:114-330 and ref:tests/core/test_audio_signal.py
(synthetic signals instead of the LFS audio)."""
import numpy as np
import pytest
import torch

from audiotools_b200 import AudioSignal, STFTParams, util
from audiotools_b200.data import transforms as tfm


def noise(B=1, C=1, T=2000, seed=0, sr=44100):
    return AudioSignal(torch.randn(B, C, T, generator=torch.Generator().manual_seed(seed)), sr)


class MulTransform(tfm.BaseTransform):
    def __init__(self, num, name=None):
        self.num = num
        super().__init__(name=name, keys=["num"])

    def _transform(self, signal, num):
        signal.audio_data = signal.audio_data * num[:, None, None]
        return signal

    def _instantiate(self, state):
        return {"num": self.num}


MULS = [0.5, 0.25, 0.125]


@pytest.mark.parametrize("build", [
    lambda: tfm.Compose([MulTransform(x) for x in MULS]),
    lambda: tfm.Compose([MulTransform(MULS[0]),
                         tfm.Compose([MulTransform(MULS[1]), tfm.Compose([MulTransform(MULS[2])])])]),
    lambda: tfm.Compose([tfm.Compose([MulTransform(MULS[0])]),
                         tfm.Compose([MulTransform(MULS[1]), MulTransform(MULS[2])])]),
])
def test_compose_products(build):
    transform = build()
    signal = noise()
    out = transform(signal.clone(), **transform.instantiate(0))
    assert torch.allclose(out.audio_data, signal.audio_data * np.prod(MULS))


def test_compose_naming_indexing_and_filter():
    transform = tfm.Compose([MulTransform(x, name=str(x)) for x in MULS])
    assert [t.name for t in transform] == ["0.0.5", "1.0.25", "2.0.125"]
    assert len(transform) == 3 and isinstance(transform[1], MulTransform)
    kwargs = transform.instantiate(0)
    assert set(kwargs["Compose"]) == {"0.0.5", "1.0.25", "2.0.125", "mask"}
    signal = noise()
    rng = np.random.RandomState(0)
    for size in range(len(MULS)):
        for _ in range(5):
            chosen = rng.choice(MULS, size=size, replace=False).tolist()
            with transform.filter(*[str(x) for x in chosen]):
                out = transform(signal.clone(), **kwargs)
            assert torch.allclose(out.audio_data, signal.audio_data * np.prod(chosen))
    assert transform.transforms_to_apply == ["0.0.5", "1.0.25", "2.0.125"]  # restored


def test_choose_one_hot_and_batch():
    signal = noise()
    transform = tfm.Choose([MulTransform(0.0), MulTransform(2.0)])
    targets = [signal.clone() * 0.0, signal.clone() * 2.0]
    for seed in range(10):
        kwargs = transform.instantiate(seed, signal)
        assert sum(int(m.item()) for m in kwargs["Choose"]["one_hot"]) == 1
        out = transform(signal.clone(), **kwargs)
        assert any(out == t for t in targets)
    batch = AudioSignal.batch([signal.clone() for _ in range(4)])
    kwargs = transform.batch_instantiate([0, 1, 2, 3], batch)
    out = transform(batch, **kwargs)
    for nb in range(4):
        assert any(out[nb] == t for t in targets)
    weighted = tfm.Choose([MulTransform(0.0), MulTransform(2.0)], weights=[0.0, 1.0])
    out = weighted(AudioSignal.batch([signal.clone() for _ in range(4)]), **weighted.batch_instantiate([0, 1, 2, 3], batch))
    for nb in range(4):
        assert out[nb] == targets[1]
    nested = tfm.Choose([tfm.Compose([MulTransform(0.0)]), tfm.Compose([MulTransform(2.0)])])
    for seed in range(6):
        assert any(nested(signal.clone(), **nested.instantiate(seed, signal)) == t for t in targets)


def test_repeat_and_repeat_up_to():
    signal = AudioSignal(torch.randn(1, 1, 100).clamp(1e-5), 44100)
    transform = tfm.Repeat(MulTransform(0.5), n_repeat=3)
    out = transform(signal.clone(), **transform.instantiate(0, signal))
    assert (out.audio_data / signal.audio_data).mean() == 0.5 ** 3
    up = tfm.RepeatUpTo(MulTransform(0.5), max_repeat=4)
    out = up(signal.clone(), **up.instantiate(3, signal))
    ratio = (out.audio_data / signal.audio_data).mean().item()
    assert any(abs(ratio - 0.5 ** n) < 1e-7 for n in (1, 2, 3))


def test_masks_gate_items_and_instantiate_is_seed_reproducible():
    transform = tfm.Compose([MulTransform(3.0), tfm.VolumeChange(prob=0.5)], prob=0.7)
    a, b = transform.instantiate(5), transform.instantiate(5)
    fa, fb = util.flatten(a), util.flatten(b)
    assert fa.keys() == fb.keys() and all(torch.equal(fa[k], fb[k]) for k in fa)
    batch = noise(B=6)
    kwargs = MulTransform(3.0).batch_instantiate(list(range(6)), batch)
    kwargs["MulTransform"]["mask"] = torch.tensor([True, False, True, False, False, True])
    out = MulTransform(3.0)(batch.clone(), **kwargs)
    for i, m in enumerate(kwargs["MulTransform"]["mask"].tolist()):
        assert torch.allclose(out.audio_data[i], batch.audio_data[i] * (3.0 if m else 1.0))
    kwargs["MulTransform"]["mask"][:] = False
    assert MulTransform(3.0)(batch.clone(), **kwargs) == batch


def test_keys_come_from_transform_signature():
    assert tfm.LowPass().keys == ["cutoff", "mask"]
    assert tfm.RoomImpulseResponse(sources=[noise()]).keys == ["ir_signal", "drr", "eq", "mask"]
    p = tfm.Equalizer(n_bands=6).instantiate(0)["Equalizer"]
    assert p["eq"].shape == (6,) and (p["eq"] <= 0).all() and p["mask"].item() is True
    expected = -1.0 * np.random.RandomState(0).rand(6)
    assert np.allclose(p["eq"].numpy(), expected)  # same draw order as the reference (ref :594-597)


def test_pool_transforms_instantiate_inside_compose():
    """Every concrete transform must draw through BaseTransform._draw when it is a child of Compose (a pool transform
    that shadowed `_draw` broke cfg4's Compose[Equalizer + RoomImpulseResponse + PitchShift])."""
    irs = [noise(B=1, T=500) for _ in range(3)]
    pool = [noise(B=1, T=4000) for _ in range(2)]
    transform = tfm.Compose([tfm.Equalizer(), tfm.RoomImpulseResponse(sources=irs),
                             tfm.PitchShift(("choice", [-2, -1, 1, 2])), tfm.BackgroundNoise(sources=pool),
                             tfm.CrossTalk(sources=pool)])
    sig = noise(B=4, T=4000)
    kwargs = transform.batch_instantiate(list(range(4)), sig)
    flat = util.flatten(kwargs)
    assert flat[("Compose", "1.RoomImpulseResponse", "ir_signal")].shape == (4, 1, sig.sample_rate)
    assert flat[("Compose", "2.PitchShift", "n_semitones")].shape == (4,)
    assert all(v.shape[0] == 4 for v in flat.values())
    for cls in (c for c in vars(tfm).values() if isinstance(c, type) and issubclass(c, tfm.BaseTransform)):
        assert cls._draw is tfm.BaseTransform._draw, f"{cls.__name__} shadows BaseTransform._draw"


def test_sample_from_dist_and_ensure_tensor():
    assert util.sample_from_dist(("const", 3)) == 3
    v = util.sample_from_dist(("uniform", 1.0, 2.0), 0)
    assert v == np.random.RandomState(0).uniform(1.0, 2.0)
    assert util.sample_from_dist(("choice", [4, 8]), 1) in (4, 8)
    t = util.ensure_tensor(2.0, 2, 3)
    assert t.shape == (3, 1)
    assert util.ensure_tensor(np.arange(3), 2, 3).shape == (3, 1)
    with pytest.raises(ValueError):
        util.random_state("nope")


def test_audio_signal_container_semantics():
    with pytest.raises(ValueError):
        AudioSignal(object(), 16000)
    with pytest.raises(AssertionError):
        AudioSignal(torch.zeros(10))
    s = AudioSignal(np.zeros(100), 16000)  # float64 numpy -> float32 [1,1,T]
    assert s.audio_data.dtype == torch.float32 and s.shape == (1, 1, 100)
    assert s.stft_params == STFTParams(512, 128, "hann", False, "reflect")  # 2**ceil(log2(0.032*sr))
    assert AudioSignal(torch.zeros(44100), 44100).stft_params.window_length == 2048
    with pytest.raises(AssertionError):
        s.audio_data = torch.zeros(5)
    s._loudness = torch.tensor([-20.0])
    s.audio_data = s.audio_data * 2  # the setter drops the loudness cache ...
    assert s._loudness is None
    b = noise(B=3)
    b._loudness = torch.tensor([-1.0, -2.0, -3.0])
    sub = b[torch.tensor([True, False, True])]
    assert sub.batch_size == 2 and torch.equal(sub._loudness, torch.tensor([-1.0, -3.0]))
    sub.audio_data = sub.audio_data * 0  # (clears sub's cache)
    b[torch.tensor([True, False, True])] = sub  # ... but __setitem__ keeps the parent's (Silence relies on it)
    assert torch.equal(b._loudness, torch.tensor([-1.0, -2.0, -3.0]))
    assert b.audio_data[0].abs().sum() == 0 and b.audio_data[1].abs().sum() > 0
    c = b.clone()
    assert c == b and c.audio_data.data_ptr() != b.audio_data.data_ptr()
    c += 1.0
    assert c != b and torch.allclose((c - 1.0).audio_data, b.audio_data, atol=1e-6)
    assert torch.equal((2 * b).audio_data, (b * 2).audio_data)
    s = noise(T=100)
    assert s.zero_pad(5, 7).signal_length == 112 and s.trim(5, 7).signal_length == 100
    assert s.zero_pad_to(150).signal_length == 150 and s.truncate_samples(60).signal_length == 60
    assert noise(C=2).to_mono().num_channels == 1
    assert s.compute_stft_padding(512, 128, False) == (0, 0)
    assert noise(T=1000).compute_stft_padding(256, 64, True) == (24, 96)
    with pytest.raises(AssertionError):
        s.compute_stft_padding(512, 100, True)


def test_batching_rules():
    a, b = noise(T=100), noise(T=80, seed=1)
    with pytest.raises(RuntimeError, match="same length"):
        AudioSignal.batch([a.clone(), b.clone()])
    assert AudioSignal.batch([a.clone(), b.clone()], pad_signals=True).shape == (2, 1, 100)
    assert AudioSignal.batch([a.clone(), b.clone()], truncate_signals=True).shape == (2, 1, 80)
    with pytest.raises(RuntimeError, match="same sample rate"):
        AudioSignal.batch([a.clone(), noise(T=100, sr=16000)])
    d = util.collate([{"signal": a.clone(), "x": {"y": torch.tensor(1.0)}}, {"signal": b.clone(), "x": {"y": torch.tensor(2.0)}}])
    assert d["signal"].shape == (2, 1, 100) and torch.equal(d["x"]["y"], torch.tensor([1.0, 2.0]))
    assert util.unflatten(util.flatten({"a": {"b": 1, "c": {"d": 2}}, "e": 3})) == {"a": {"b": 1, "c": {"d": 2}}, "e": 3}


def test_deferred_gain_on_cpu_is_plain_arithmetic():
    s = noise(B=2)
    x = s.audio_data.clone()
    s.volume_change(torch.tensor([-6.0, 0.0]))
    assert s._pending_gain is None
    assert torch.allclose(s.audio_data[0], x[0] * 10 ** (-6 / 20), rtol=1e-6) and torch.equal(s.audio_data[1], x[1])


def test_drr_algebra_matches_reference_on_cpu(golden):
    """decompose_ir / measure_drr / alter_drr are host-side tensor algebra (no kernel): the vectorised early/late
    split must reproduce the real reference's outputs (tests/golden/make_golden.py) without its per-item loop."""
    from tests.conftest import rel_err
    from tests.golden import cases

    ir = cases.make_ir()
    drr = torch.from_numpy(golden["drr"])
    assert rel_err(AudioSignal(ir.clone(), 44100).alter_drr(drr).audio_data, torch.from_numpy(golden["alter_drr"])) < 1e-5
    assert torch.allclose(AudioSignal(ir.clone(), 44100).measure_drr(), torch.from_numpy(golden["measure_drr"]), atol=1e-3)
    stereo = AudioSignal(torch.cat([ir, 0.5 * ir.roll(7, -1)], dim=1), 44100)
    early, late, window = stereo.decompose_ir()
    assert torch.equal(window[:, 0], window[:, 1])  # channel 0's early region serves every channel (ref :569-573)
    assert torch.equal(early + late, stereo.audio_data)


def test_all_true_mask_fast_path_matches_masked_round_trip():
    """BaseTransform.transform skips the gather / scatter copies when every item is selected; the samples and the
    cache state must equal what the masked round trip of the reference leaves (ref transforms.py:133-166,
    audio_signal.py:1658-1679: a cache is only overwritten when both sides hold one)."""
    t = MulTransform(0.5)
    sig = noise(B=3)
    kw = t.batch_instantiate([0, 1, 2])
    assert bool(kw["MulTransform"]["mask"].all())
    # generic path, forced by a mask with one item off; then only compare the selected items
    ref = sig.clone()
    ref._loudness = torch.tensor([-20.0, -21.0, -22.0])
    kw_part = {"MulTransform": {"num": kw["MulTransform"]["num"], "mask": torch.tensor([True, True, False])}}
    ref = t(ref, **kw_part)
    out = sig.clone()
    out._loudness = torch.tensor([-20.0, -21.0, -22.0])
    out = t(out, **kw)
    assert torch.equal(out.audio_data[:2], ref.audio_data[:2]) and torch.equal(out.audio_data[2], sig.audio_data[2] * 0.5)
    assert torch.equal(out._loudness, ref._loudness)  # the stale cache survives in both paths, as in the reference
    assert util.host_view(kw["MulTransform"]["mask"]) is kw["MulTransform"]["mask"]  # CPU tensors are their own view
    moved = util.prepare_batch(kw, "cpu")
    assert not hasattr(moved["MulTransform"]["mask"], "_b2a_host")


def test_chunking_helpers_match_reference(golden_spec):
    """windows / collect_windows / overlap_and_add are container reshapes (ref:audiotools/core/dsp.py:15-151): same
    windows and the same reassembled signal as the real reference (tests/golden/make_golden_spectral.py)."""
    from tests.golden import cases

    x = cases.make_input("cfg1")
    x2 = torch.cat([x[:1], 0.5 * x[:1].flip(-1)], 1)
    s = AudioSignal(x2.clone(), 16000).collect_windows(0.1, 0.05)
    assert torch.equal(s.audio_data, torch.from_numpy(golden_spec["win_collect"]))
    s.audio_data = s.audio_data * 0.5 + 0.1
    out = s.overlap_and_add(0.05).audio_data
    assert out.shape == (1, 2, 16000) and torch.allclose(out, torch.from_numpy(golden_spec["win_ola"]), atol=1e-6)
    wins = list(AudioSignal(x2.clone(), 16000).windows(0.064, 0.016))
    assert len(wins) == int(golden_spec["win_iter_count"]) and wins[0].audio_data.shape == (1, 1, 1024)
    # identity round trip: windows of an untouched signal add back to the signal
    r = AudioSignal(x2.clone(), 16000).collect_windows(0.1, 0.05).overlap_and_add(0.05)
    assert torch.allclose(r.audio_data, x2, atol=1e-6)


def test_elementwise_effects_match_reference(golden_spec):
    """clip_distortion / quantization / mulaw_quantization / ensure_max_of_audio and their transforms are container
    arithmetic (ref:audiotools/core/effects.py:181-198,435-523; transforms.py:531-666,1006-1050,1173-1197): same drawn
    parameters and outputs as the real reference (tests/golden/make_golden_spectral.py)."""
    from tests.golden import cases

    g = golden_spec
    xs = cases.make_input("cfg1") * 0.3
    xs2 = torch.cat([xs, 0.5 * xs.flip(-1)], 1)
    T = lambda k: torch.from_numpy(g[k])
    assert torch.allclose(AudioSignal(xs.clone(), 16000).clip_distortion(torch.tensor([0.05, 0.2, 0.0, 0.5])).audio_data,
                          T("fx_clip"), atol=1e-7)
    q = torch.tensor([8, 16, 256, 3])
    assert torch.allclose(AudioSignal(xs2.clone(), 16000).quantization(q).audio_data, T("fx_quant"), atol=1e-6)
    assert torch.allclose(AudioSignal(xs2.clone(), 16000).mulaw_quantization(q).audio_data, T("fx_mulaw"), atol=1e-6)
    assert torch.allclose(AudioSignal(xs2.clone() * 5, 16000).ensure_max_of_audio(0.7).audio_data, T("fx_maxaudio"), atol=1e-7)
    for cls, kw, inp in (("ClippingDistortion", {}, xs), ("Quantization", {}, xs2), ("MuLawQuantization", {}, xs2),
                         ("RescaleAudio", {"val": 0.5}, xs2)):
        t = getattr(tfm, cls)(**kw)
        s_in = AudioSignal(inp.clone() * 3, 16000)
        k = t.batch_instantiate([3, 4, 5, 6], s_in)
        for kk, v in util.flatten(k).items():
            assert np.allclose(v.numpy(), g[f"fxkw/{cls}/" + "/".join(kk)]), (cls, kk)
        assert torch.allclose(t(s_in.clone(), **k).audio_data, T(f"fxout/{cls}"), atol=1e-6), cls
    sg = AudioSignal(xs2.clone(), 16000)
    sg.metadata["loudness"] = -30.0
    db = tfm.GlobalVolumeNorm(db=("uniform", -20, -10)).instantiate(7, sg)["GlobalVolumeNorm"]["db"]
    assert np.allclose(db.numpy(), g["fx_globalvolnorm_db"])


def test_small_util_helpers():
    """hz_to_bin / choose_from_list_of_lists / chdir (ref:audiotools/core/util.py:100-126,302-343); the expected bins
    below are the real reference's."""
    import os

    hz = torch.tensor([[0.0, 100.0, 440.0], [7999.0, 12000.0, 30000.0]])
    assert util.hz_to_bin(hz, 2048, 44100).tolist() == [[0, 5, 20], [372, 558, 1025]]
    item, si, ii = util.choose_from_list_of_lists(np.random.RandomState(3), [[1, 2, 3], [4, 5], [6]], p=[0.5, 0.3, 0.2])
    assert item == [[1, 2, 3], [4, 5], [6]][si][ii]
    here = os.getcwd()
    with util.chdir("/tmp"):
        assert os.getcwd() == "/tmp"
    assert os.getcwd() == here


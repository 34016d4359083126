"""BASELINE.json configs[2..4] AT THEIR STATED SIZES on one GPU's share (``-m gpu``), each checked against the oracle on
a strided subset of items -- the subset is what keeps the CPU side to seconds; the GPU side runs the whole batch:

  cfg3  batch=256 mono 30 s @ 48 kHz: resample -> 16 kHz + low_pass(8 kHz)
  cfg4  per-GPU share (128 of 512) mono 10 s @ 44.1 kHz: Compose[Equalizer + RoomImpulseResponse + PitchShift(+-2)]
  cfg5  per-GPU share (256 of 2048) 2ch 10 s @ 44.1 kHz: the same Compose + LUFS normalize(-24) + log-mel

Criteria: ``rel_err`` (max|a-b| / max|b|) < 1e-4 AND the element-wise criterion of tests/conftest.py; shapes exactly.
The pitch stage is compared with the specification oracle (oracle/pitch_spec.py) the way the golden test does it:
splice positions exactly wherever the correlation margin decides them, waveforms on rows whose every splice agrees.
"""
import numpy as np
import pytest
import torch

from tests.conftest import elementwise_ok, rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
TOL = 1e-4


@pytest.fixture(scope="module")
def at():
    import __graft_entry__ as graft

    graft.build()
    import audiotools_b200

    return audiotools_b200


@pytest.fixture(scope="module")
def sp():
    from oracle import signal_path

    return signal_path


def _close(a, b, tol=TOL):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert rel_err(a, b) < tol, rel_err(a, b)
    # waveforms: per-cell |a - b| <= 1e-4 |b| + 1e-5 max|row| -- a quiet row cannot hide behind a loud one
    assert elementwise_ok(a, b, rtol=tol, atol_frame=1e-5, frame_dim=-1), "element-wise criterion"


def test_cfg3_full_size_strided_oracle(at, sp):
    B, T = 256, 1_440_000
    g = torch.Generator().manual_seed(3)
    x = torch.empty(B, 1, T)
    for i in range(0, B, 32):  # chunked generation keeps the host peak low
        x[i:i + 32] = 0.1 * torch.randn(32, 1, T, generator=g)
    x *= 0.05 + 0.95 * torch.rand(B, 1, 1, generator=g)
    sig = at.AudioSignal(x, 48000).to(DEV).resample(16000)
    assert sig.sample_rate == 16000 and sig.audio_data.shape == (B, 1, 480000)
    mid = sig.audio_data.clone()
    y = sig.low_pass(8000).audio_data
    assert y.shape == (B, 1, 480000) and torch.isfinite(y).all()
    sub = [0, 37, 128, 255]
    ref_mid = sp.resample(x[sub], 48000, 16000)
    _close(mid[sub], ref_mid)
    _close(y[sub], sp.low_pass(ref_mid, 16000, 8000))
    # batch == per-item (the reference's own property test, ref:tests/core/test_audio_signal.py resample / low_pass)
    one = at.AudioSignal(x[200:201].clone(), 48000).to(DEV).resample(16000).low_pass(8000).audio_data
    assert torch.equal(one, y[200:201])


def _augment(at, B, C, seed):
    from audiotools_b200.data import transforms as tfm

    sr, T = 44100, 441000
    g = torch.Generator().manual_seed(seed)
    x = (0.1 * torch.randn(B, C, T, generator=g)).clamp(-1, 1) * (0.05 + 0.95 * torch.rand(B, 1, 1, generator=g))
    t = torch.arange(sr) / sr
    irs = []
    for i in range(4):
        h = torch.randn(1, 1, sr, generator=g) * torch.exp(-t / (0.15 + 0.1 * i)) * 0.1
        h[..., 40 + 7 * i] = 1.0
        irs.append(at.AudioSignal(h, sr))
    transform = tfm.Compose([tfm.Equalizer(), tfm.RoomImpulseResponse(sources=irs),
                             tfm.PitchShift(("choice", [-2, 2]))])
    sig = at.AudioSignal(x.clone(), sr)
    kwargs = transform.batch_instantiate(list(range(B)), sig)
    return x, sr, transform, kwargs


def _check_augment_stages(at, sp, x, sr, transform, kwargs, sub):
    """Run the Compose on the whole batch; check every stage on the items ``sub`` against the oracle.  Returns the
    augmented signal (device)."""
    from audiotools_b200.engine import get_engine
    from oracle import pitch_spec as ps

    eng = get_engine()
    dk = at.util.prepare_batch(kwargs, DEV)
    sig = at.AudioSignal(x.clone(), sr).to(DEV)
    with transform.filter("Equalizer"):
        s1 = transform(sig.clone(), **dk).audio_data
    with transform.filter("Equalizer", "RoomImpulseResponse"):
        s2 = transform(sig.clone(), **dk).audio_data
    out = transform(sig.clone(), **dk)
    B, C, T = x.shape
    assert out.audio_data.shape == (B, C, T) and torch.isfinite(out.audio_data).all()
    flat = at.util.flatten(kwargs)
    eq = flat[("Compose", "0.Equalizer", "eq")]
    ir = flat[("Compose", "1.RoomImpulseResponse", "ir_signal")]
    ir = (ir.audio_data if hasattr(ir, "audio_data") else ir).cpu()  # (prepare_batch moved the AudioSignal in place)
    ir_eq = flat[("Compose", "1.RoomImpulseResponse", "eq")]
    drr = flat[("Compose", "1.RoomImpulseResponse", "drr")]
    st = flat[("Compose", "2.PitchShift", "n_semitones")]
    assert all(bool(flat[k].all()) for k in flat if k[-1] == "mask")  # prob = 1: every stage applies to every item
    r1 = sp.equalizer(x[sub], sr, eq[sub])
    _close(s1[sub], r1)
    r2 = sp.apply_ir(r1, ir[sub], sr, drr[sub], ir_eq[sub])
    _close(s2[sub], r2)
    # pitch stage: oracle on OUR stage-2 rows (so that a near-tie flip cannot be caused by the 1e-5 differences above)
    agree = 0
    rows = 0
    for i in sub:
        semis = float(st[i])
        y_i, pos_i = eng.pitch_shift(s2[i:i + 1], sr, semis, return_positions=True)
        assert torch.equal(y_i, out.audio_data[i:i + 1])  # batched multi-shift launch == single launch
        for c in range(C):
            ref_y, ref_pos, margin = ps.pitch_shift_row(s2[i, c].double().cpu().numpy(), sr, semis)
            pos = pos_i.reshape(C, -1)[c].cpu().numpy()[: len(ref_pos)]
            decided = margin > 1e-5
            assert np.array_equal(pos[decided], ref_pos[decided])
            rows += 1
            if np.array_equal(pos, ref_pos):
                agree += 1
                yc = y_i[0, c].cpu().numpy()
                assert np.abs(yc - ref_y).max() / np.abs(ref_y).max() < TOL
    assert agree * 2 >= rows
    return out


def test_cfg4_full_size_strided_oracle(at, sp):
    x, sr, transform, kwargs = _augment(at, B=128, C=1, seed=4)
    _check_augment_stages(at, sp, x, sr, transform, kwargs, sub=[0, 41, 86, 127])


def test_cfg5_full_size_strided_oracle(at, sp):
    x, sr, transform, kwargs = _augment(at, B=256, C=2, seed=5)
    out = _check_augment_stages(at, sp, x, sr, transform, kwargs, sub=[3, 130, 255])
    aug = out.audio_data.clone()
    lufs = out.loudness().clone()
    out.normalize(-24.0)
    logmel = out.mel_spectrogram(n_mels=128, window_length=2048, hop_length=512, window_type="hann", log=True)
    assert logmel.shape == (256, 2, 128, 862)
    sub = [0, 64, 131, 255]
    y_ref, l_ref = sp.normalize(aug[sub].cpu(), sr, -24.0)
    assert torch.allclose(lufs[sub].cpu(), l_ref, atol=2e-3)
    _close(out.audio_data[sub], y_ref)
    lm_ref = sp.log_mel(sp.mel_spectrogram(y_ref, sr, 128, window_length=2048, hop_length=512, window_type="hann"))
    assert rel_err(logmel[sub].cpu(), lm_ref) < TOL
    assert (logmel[sub].cpu() - lm_ref).abs().max().item() < 2e-4
    assert torch.allclose(out.loudness().cpu(), torch.full((256,), -24.0), atol=1e-2)

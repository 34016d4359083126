"""Pin the oracle: ``oracle/signal_path.py`` vs outputs of the REAL reference
(``tests/golden/reference_golden.npz``), plus the known-answer constants and
property tests the reference's own test-suite holds for the third-party pieces
(SURVEY.md §8c).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import signal_path as sp
from oracle import third_party as tp
from tests.golden import cases

T = torch.from_numpy


def close(a, b, atol=1e-6, rtol=1e-5):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, atol=atol, rtol=rtol, equal_nan=True), (a - b).abs().max()


# ---------------------------------------------------------------------------
# first-party restatement vs the real reference
# ---------------------------------------------------------------------------
def test_stft_family(golden):
    x = cases.make_input("cfg1")
    s = sp.stft(x, 16000, window_length=512, hop_length=128)
    assert s.shape == (4, 1, 257, 126)  # bit-exact frame count: 1 + T//hop
    close(s, golden["cfg1_stft"])
    close(sp.stft(x, 16000), golden["cfg1_stft"])
    x2 = x[:2]
    close(sp.stft(x2, 16000, 256, 64, "sqrt_hann", True, "reflect"), golden["cfg1_stft_match_stride"])
    s_odd = sp.stft(x2[..., :15999], 16000, 256, 64, "hann", True, "reflect")
    assert s_odd.shape[-1] == 250  # == ceil(15999 / 64)
    close(s_odd, golden["cfg1_stft_match_stride_odd"])
    close(sp.stft(x2, 16000, 256, 100, "average"), golden["cfg1_stft_average_hop100"])
    close(sp.mel_spectrogram(x, 16000, n_mels=80), golden["cfg1_mel80"], atol=1e-5)
    close(sp.mel_spectrogram(x2, 16000, 40, 100.0, 6000.0, window_length=1024, hop_length=256),
          golden["cfg1_mel40_fmin_fmax"], atol=1e-5)
    close(sp.mfcc(x, 16000), golden["cfg1_mfcc"], atol=1e-4)
    close(sp.log_magnitude(sp.stft(x2, 16000)), golden["cfg1_logmag"], atol=1e-4)
    close(sp.istft(sp.stft(x2, 16000), 16000, 16000), golden["cfg1_istft"], atol=1e-5)
    close(sp.istft(sp.stft(x2, 16000, 256, 64, "sqrt_hann", True), 16000, 16000, 256, 64, "sqrt_hann", True),
          golden["cfg1_istft_match_stride"], atol=1e-5)


def test_lufs_normalize_logmel_cfg2(golden):
    x = cases.make_input("cfg2")
    m = sp.Meter(44100)
    z = m.block_energies(x.permute(0, 2, 1))
    assert z.shape == golden["cfg2_z"].shape  # block count bit-exact
    close(z, golden["cfg2_z"], atol=0, rtol=1e-6)
    l = sp.loudness(x, 44100)
    assert l[1].item() == -70.0  # silent item: NaN scrub -> -inf -> clamp
    close(l, golden["cfg2_lufs"], atol=1e-5)
    y, _ = sp.normalize(x, 44100, -24.0)
    close(y, golden["cfg2_norm"], atol=1e-7, rtol=1e-5)
    mel = sp.mel_spectrogram(y, 44100, 128, window_length=2048, hop_length=512, window_type="hann")
    close(mel, golden["cfg2_mel"], atol=1e-5, rtol=1e-5)
    close(sp.log_mel(mel), golden["cfg2_logmel"], atol=1e-4)
    close(sp.loudness(x[:, :1], 44100, use_fir=True), golden["cfg2_lufs_fir_mono"], atol=1e-4)


def test_lufs_edge_cases(golden):
    close(sp.loudness(cases.make_input("lufs16k"), 16000), golden["lufs16k"], atol=1e-5)
    close(sp.loudness(cases.make_input("short"), 16000), golden["lufs_short"], atol=1e-5)
    close(sp.loudness(cases.make_input("lufs48k"), 48000), golden["lufs48k"], atol=1e-5)
    close(sp.loudness(cases.make_input("lufs11k"), 11025), golden["lufs11k"], atol=1e-5)
    x = cases.make_input("lufs16k")[:4]
    db = T(golden["norm16k_db"])
    close(sp.normalize(x, 16000, db)[0], golden["norm16k"], rtol=1e-5)
    close(sp.volume_change(x, db), golden["volchange16k"], rtol=1e-6)


def test_resample_and_fir(golden):
    x = cases.make_input("rs")
    close(sp.resample(x, 48000, 16000), golden["rs_48k_16k"])
    close(sp.resample(x[..., :22050], 44100, 16000), golden["rs_44k_16k"])
    close(sp.resample(x[..., :8000], 16000, 44100), golden["rs_16k_44k"])
    close(sp.resample(x[..., :8001], 16000, 48000), golden["rs_16k_48k"])
    close(sp.resample(x[..., :4410], 44100, 48000), golden["rs_44k_48k"])
    close(sp.low_pass(sp.resample(x, 48000, 16000), 16000, 8000), golden["rs_48k_16k_lp8k"])
    x7 = cases.make_input("fir")
    cut = T(golden["fir_cut"])
    close(sp.low_pass(x7, 44100, cut), golden["lp_peritem"])
    close(sp.high_pass(x7, 44100, cut / 8), golden["hp_peritem"])
    close(sp.low_pass(x7, 44100, 4000), golden["lp_scalar"])


def test_equalizer_and_ir(golden):
    x7 = cases.make_input("fir")
    eq = golden["eq_db"]
    close(sp.equalizer(x7, 44100, eq), golden["eq_out"])
    close(sp.equalizer(x7, 44100, eq[0]), golden["eq_out_1d"])
    close(sp.mel_filterbank(x7[:1, :1], 44100, 4), golden["fbank4"])
    ir = cases.make_ir()
    close(sp.convolve(x7, ir), golden["conv_out"], atol=1e-6)
    close(sp.convolve(x7, ir, start_at_max=False), golden["conv_out_nomax"], atol=1e-6)
    close(sp.apply_ir(x7, ir, 44100), golden["applyir_plain"], atol=1e-6)
    drr = T(golden["drr"])
    close(sp.alter_drr(ir, 44100, drr), golden["alter_drr"], atol=1e-6)
    close(sp.measure_drr(ir, 44100), golden["measure_drr"], atol=1e-4)
    close(sp.apply_ir(x7, ir, 44100, drr=drr, ir_eq=eq), golden["applyir_full"], atol=1e-6)


# ---------------------------------------------------------------------------
# third-party restatements: known-answer constants + the reference's property tests
# ---------------------------------------------------------------------------
def test_k_weighting_coefficients_vs_itu_table():
    """ITU-R BS.1770-4 table 1/2 (48 kHz).  pyloudnorm's RBJ design lands within 4e-5."""
    f = tp.k_weighting_filters(48000)
    assert list(f) == ["high_shelf", "high_pass"]  # shelf first (ref loudness.py:115)
    np.testing.assert_allclose(f["high_shelf"].b, [1.53512485958697, -2.69169618940638, 1.19839281085285], rtol=5e-5)
    np.testing.assert_allclose(f["high_shelf"].a, [1.0, -1.69065929318241, 0.73248077421585], rtol=5e-5)
    np.testing.assert_allclose(f["high_pass"].a, [1.0, -1.99004745483398, 0.99007225036621], rtol=5e-5)
    np.testing.assert_allclose(f["high_pass"].b / f["high_pass"].b[0], [1.0, -2.0, 1.0], atol=1e-12)
    with pytest.raises(NotImplementedError):
        tp.k_weighting_filters(48000, "Fenton/Lee 1")


@pytest.mark.parametrize("sr,n_fft,n_mels", [(44100, 2048, 128), (16000, 512, 80), (44100, 512, 40)])
def test_librosa_mel_vs_torchaudio(sr, n_fft, n_mels):
    import torchaudio

    ours = tp.librosa_mel(sr, n_fft, n_mels)
    ta = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, sr / 2, n_mels, sr, norm="slaney",
                                               mel_scale="slaney").T.numpy()
    assert ours.dtype == np.float32 and ours.shape == (n_mels, n_fft // 2 + 1)
    np.testing.assert_allclose(ours, ta, atol=5e-7)


def _sine_1khz(sr, seconds, peak_dbfs, channels):
    t = torch.arange(int(sr * seconds)) / sr
    x = 10 ** (peak_dbfs / 20) * torch.sin(2 * np.pi * 1000 * t)
    return x[None, None].repeat(1, channels, 1).float()


@pytest.mark.parametrize("sr", [44100, 48000])
def test_lufs_known_answers(sr):
    """ref:tests/core/test_loudness.py:110-116,164-170 (1 kHz stereo at -24/-23 dBFS -> -24/-23 LKFS,
    atol 0.1) and :56-62 (1 kHz full-scale mono -> -3.05, atol 0.1); re-synthesised tones."""
    assert abs(sp.loudness(_sine_1khz(sr, 20, -24, 2), sr).item() + 24.0) < 0.1
    assert abs(sp.loudness(_sine_1khz(sr, 20, -23, 2), sr).item() + 23.0) < 0.1
    assert abs(sp.loudness(_sine_1khz(sr, 1.06, 0, 1), sr).item() + 3.0523438444331137) < 0.1


def test_normalize_then_measure():
    """ref:tests/core/test_effects.py:15-33."""
    x = torch.randn(1, 2, 32000, generator=torch.Generator().manual_seed(0))
    for db in (-70.0, -50.0, -30.0, -10.0):
        y, _ = sp.normalize(x, 16000, db)
        assert abs(sp.loudness(y, 16000).item() - db) < 0.1


def test_low_high_pass_sine_thresholds():
    """ref:tests/core/test_dsp.py:76-109."""
    sr, f = 44100, 440
    t = torch.arange(sr) / sr
    x = (torch.sin(2 * np.pi * f * t) * torch.hann_window(sr))[None, None]
    assert sp.low_pass(x, sr, 220).abs().max() < 1e-4
    assert (sp.low_pass(x, sr, 880) - x).abs().max() < 1e-3
    assert (sp.high_pass(x, sr, 220) - x).abs().max() < 1e-4
    both = sp.low_pass(x.repeat(2, 1, 1), sr, torch.tensor([220.0, 880.0]))
    assert both[0].abs().max() < 1e-4 and (both[1] - x[0]).abs().max() < 1e-3


@pytest.mark.parametrize("n_bands", [1, 2, 4, 8, 12, 16])
def test_filterbank_sums_to_input(n_bands):
    """ref:tests/core/test_effects.py:184-231."""
    x = cases.make_input("fir")[:1]
    fb = sp.mel_filterbank(x, 44100, n_bands)
    assert fb.shape[-1] == n_bands
    assert torch.allclose(fb.sum(-1), x, atol=1e-6)
    assert torch.allclose(sp.equalizer(x, 44100, np.zeros(n_bands)), x, atol=1e-6)


def test_delta_ir_is_identity():
    """ref:tests/core/test_effects.py:86-121."""
    x = cases.make_input("fir")
    for delay in (0, 1, 777):
        ir = torch.zeros(3, 1, 1000)
        ir[..., delay] = 1.0
        assert torch.allclose(sp.convolve(x, ir), x, atol=1e-6)


def test_resample_lengths_and_dc():
    """ref:tests/core/test_audio_signal.py:524-533 (length/sr only) + DC preservation."""
    for old, new, T_ in [(48000, 16000, 1000), (44100, 16000, 4410), (16000, 44100, 1600), (44100, 48000, 441)]:
        y = sp.resample(torch.ones(1, 1, T_), old, new)
        assert y.shape[-1] == int(np.floor(new * T_ / old))
        assert torch.allclose(y, torch.ones_like(y), atol=1e-5)
    k, width, old, new = tp.resample_kernel(48000, 16000)
    assert (old, new, width, k.shape) == (3, 1, 77, (1, 157))
    k, width, old, new = tp.resample_kernel(44100, 16000)
    assert (old, new, k.shape) == (441, 160, (160, 2 * width + 441))


def test_unfold_block_count():
    """julius.core.unfold frame count: cfg2 -> 97 blocks."""
    assert tp.unfold_num_frames(441000, 17640, 4410) == 97
    assert tp.unfold_num_frames(8000, 6400, 1600) == 2
    assert tp.unfold_num_frames(100, 6400, 1600) == 1
    u = tp.unfold(torch.arange(10.0)[None], 4, 3)
    assert u.shape == (1, 3, 4) and u[0, 2].tolist() == [6.0, 7.0, 8.0, 9.0]
    u = tp.unfold(torch.arange(11.0)[None], 4, 3)
    assert u.shape == (1, 4, 4) and u[0, 3].tolist() == [9.0, 10.0, 0.0, 0.0]


def test_spectral_masks_match_reference(golden_spec):
    """the oracle restatement of DSPMixin's spectral masks against the real reference's outputs."""
    from tests.golden import make_golden_spectral as mg

    g = golden_spec
    x = cases.make_input("cfg1")
    X = sp.stft(x, 16000)
    close = lambda a, b, atol: np.testing.assert_allclose(a.numpy(), b, atol=atol, rtol=0)
    Y = sp.mask_frequencies(X, 16000, mg.FMIN, mg.FMAX)
    close(Y, g["maskfreq_stft"], 1e-5)
    close(sp.istft(Y, 16000, 16000), g["maskfreq_audio"], 1e-5)
    close(sp.mask_frequencies(X, 16000, mg.FMIN, mg.FMAX, val=0.25)[:1], g["maskfreq_val_stft"], 1e-5)
    Y = sp.mask_timesteps(X, 1.0, mg.TMIN, mg.TMAX)
    close(Y, g["masktime_stft"], 1e-5)
    close(sp.istft(Y, 16000, 16000), g["masktime_audio"], 1e-5)
    Y = sp.mask_low_magnitudes(X, mg.DBCUT)
    close(Y[:2], g["masklow_stft"], 1e-5)
    close(sp.istft(Y, 16000, 16000), g["masklow_audio"], 1e-5)
    Y = sp.shift_phase(X, mg.SHIFT)
    close(Y[2:], g["shift_stft"], 2e-5)
    close(sp.istft(Y, 16000, 16000), g["shift_audio"], 2e-5)
    close(sp.istft(sp.shift_phase(X, torch.from_numpy(g["corrupt_in"])), 16000, 16000), g["corrupt_audio"], 2e-5)


# ------------------------------------------------------------------------------------------
# pitch_shift / time_stretch specification oracle (oracle/pitch_spec.py) vs its committed vectors
# ------------------------------------------------------------------------------------------
def test_pitch_oracle_reproduces_committed_vectors():
    import os

    from oracle import pitch_spec as ps
    from tests.golden import make_golden_pitch as mg

    g = np.load(os.path.join(os.path.dirname(mg.__file__), "pitch_golden.npz"))
    x = mg.make_input()
    assert np.array_equal(x, g["x"])  # the seeded input is reproducible
    y, pos, _ = ps.pitch_shift_row(x[0], mg.SR, 2.0)
    assert np.array_equal(pos, g["pitch_2_pos"][0])
    assert np.abs(y - g["pitch_2_y"][0]).max() < 1e-6
    y, pos, _ = ps.time_stretch_row(x[2], mg.SR, 0.8)
    assert np.array_equal(pos, g["stretch_0.8_pos"][2]) and len(y) == 30000
    assert np.abs(y - g["stretch_0.8_y"][2]).max() < 1e-6
    # properties of the specification itself: length kept, pitch ratio 2^(n/12), unit DC gain
    t = np.arange(mg.T) / mg.SR
    tone = (0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    y, _, _ = ps.pitch_shift_row(tone, mg.SR, 3.0)
    spec = np.abs(np.fft.rfft(y * np.hanning(mg.T)))
    assert abs(spec.argmax() * mg.SR / mg.T - 440 * 2 ** (3 / 12)) < 2.0
    dc, _, _ = ps.pitch_shift_row(np.full(9000, 0.25, np.float32), 8000, -5.0)
    assert np.abs(dc[600:-1200] - 0.25).max() < 1e-6  # (the last frames are clamped to the end of the row)


# ------------------------------------------------------------------------------------------
# window lengths that are not powers of two: the oracle against the REAL reference (make_golden_anywindow.py)
# ------------------------------------------------------------------------------------------
def test_oracle_matches_reference_for_arbitrary_window_lengths():
    import os

    from tests.conftest import rel_err
    from tests.golden import make_golden_anywindow as mg

    g = np.load(os.path.join(os.path.dirname(mg.__file__), "reference_golden_anywindow.npz"))
    x = cases.make_input("cfg1")
    for key, wl, hop, wt, ms, pt in mg.STFT_CASES:
        X = sp.stft(x, 16000, wl, hop, wt, ms, pt)
        assert X.shape[2:] == g[key + "_stft"].shape[2:]
        assert rel_err(torch.view_as_real(X[:2]), torch.view_as_real(torch.from_numpy(g[key + "_stft"]))) < 1e-6
        y = sp.istft(X, 16000, 16000, wl, hop, wt, ms)
        assert rel_err(y, torch.from_numpy(g[key + "_istft"])) < 1e-6
    mel = sp.mel_spectrogram(x, 16000, 80, window_length=400, hop_length=160, window_type="hann")
    assert rel_err(mel, torch.from_numpy(g["w400_mel80"])) < 1e-6
    mf = sp.mfcc(x, 16000, 20, 40, window_length=400, hop_length=160, window_type="hann")
    assert rel_err(mf, torch.from_numpy(g["w400_mfcc"])) < 1e-5

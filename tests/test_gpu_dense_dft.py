"""Arbitrary window lengths (``-m gpu``): the dense-DFT path of csrc/dft.cu behind AudioSignal.stft / istft /
mel_spectrogram, against the oracle (torch.stft / torch.istft semantics restated in oracle/signal_path.py, which the
REAL reference's goldens pin for the power-of-two sizes) -- 25 ms speech windows (400 @ 16 kHz, 480 @ 48 kHz/10 ms hop,
1200 @ 48 kHz), an odd length, match_stride, every padding mode, and the two power-of-two sizes whose inverse used to
be delegated to torch.istft (32, 4096).  Frame counts exactly; values to 1e-4 (global) + the element-wise criterion."""
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


def _x(B, C, T, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(B, C, T, generator=g) * (0.2 + torch.rand(B, 1, 1, generator=g))


@pytest.mark.parametrize("sr,n_fft,hop,wtype,T", [(16000, 400, 160, "hann", 48000), (48000, 480, 120, "hann", 30001),
                                                  (48000, 1200, 300, "sqrt_hann", 96000), (22050, 1001, 250, "hamming", 40000),
                                                  (16000, 400, 100, "hann", 401), (8000, 96, 31, "blackman", 5000)])
def test_stft_any_window_length_vs_oracle(at, sp, sr, n_fft, hop, wtype, T):
    x = _x(3, 2, T, n_fft + hop)
    sig = at.AudioSignal(x.clone(), sr).to(DEV)
    s = sig.stft(window_length=n_fft, hop_length=hop, window_type=wtype)
    ref = sp.stft(x, sr, n_fft, hop, wtype)
    assert s.shape == ref.shape and s.dtype == torch.complex64  # frame indexing bit-exact
    a, b = torch.view_as_real(s.cpu()), torch.view_as_real(ref)
    assert rel_err(a, b) < TOL
    assert elementwise_ok(s.cpu().abs(), ref.abs())
    # and back (the inverse runs on the dense path too)
    y = sig.istft(window_length=n_fft, hop_length=hop, window_type=wtype).audio_data
    assert y.shape == x.shape
    y_ref = sp.istft(ref, sr, T, n_fft, hop, wtype)
    assert rel_err(y.cpu(), y_ref) < TOL
    if wtype in ("hann", "sqrt_hann", "hamming") and T > 3 * n_fft:  # COLA windows: the round trip gives the signal back
        assert rel_err(y.cpu()[..., n_fft: -n_fft], x[..., n_fft: -n_fft]) < 1e-4


@pytest.mark.parametrize("padding_type", ["reflect", "constant", "replicate"])
def test_stft_any_window_length_match_stride_and_padding(at, sp, padding_type):
    x = _x(2, 1, 16000, 7)
    sig = at.AudioSignal(x.clone(), 16000).to(DEV)
    s = sig.stft(window_length=400, hop_length=100, match_stride=True, padding_type=padding_type)
    ref = sp.stft(x, 16000, 400, 100, "hann", match_stride=True, padding_type=padding_type)
    assert s.shape == ref.shape
    assert rel_err(torch.view_as_real(s.cpu()), torch.view_as_real(ref)) < TOL
    y = sig.istft(window_length=400, hop_length=100, match_stride=True).audio_data
    y_ref = sp.istft(ref, 16000, 16000, 400, 100, "hann", match_stride=True)
    assert y.shape == y_ref.shape and rel_err(y.cpu(), y_ref) < TOL


def test_mel_and_logmel_any_window_length_vs_oracle(at, sp):
    """The 25 ms / 10 ms / 80-mel speech front-end: window 400, hop 160 at 16 kHz."""
    x = _x(4, 1, 64000, 11)
    sig = at.AudioSignal(x.clone(), 16000).to(DEV)
    mel = sig.mel_spectrogram(n_mels=80, window_length=400, hop_length=160, window_type="hann")
    ref = sp.mel_spectrogram(x, 16000, 80, window_length=400, hop_length=160, window_type="hann")
    assert mel.shape == ref.shape == (4, 1, 80, 401)
    assert rel_err(mel.cpu(), ref) < TOL and elementwise_ok(mel.cpu(), ref)
    sig2 = at.AudioSignal(x.clone(), 16000).to(DEV).normalize(-20.0)  # a deferred gain must be applied first
    lm = sig2.mel_spectrogram(n_mels=80, window_length=400, hop_length=160, window_type="hann", log=True)
    y_ref, _ = sp.normalize(x, 16000, -20.0)
    lm_ref = sp.log_mel(sp.mel_spectrogram(y_ref, 16000, 80, window_length=400, hop_length=160, window_type="hann"))
    assert (lm.cpu() - lm_ref).abs().max().item() < 2e-4
    assert rel_err(sig2.audio_data.cpu(), y_ref) < TOL


@pytest.mark.parametrize("n_fft,hop", [(32, 8), (4096, 1024), (4096, 2048)])
def test_istft_formerly_delegated_sizes(at, sp, n_fft, hop):
    """window lengths 32 and 4096: no torch.istft on the path any more."""
    x = _x(2, 2, 30000, n_fft)
    sig = at.AudioSignal(x.clone(), 44100).to(DEV)
    s = sig.stft(window_length=n_fft, hop_length=hop).clone()
    real_istft = torch.istft

    def forbidden(*a, **k):
        raise AssertionError("torch.istft called")

    torch.istft = forbidden
    try:
        y = sig.istft(window_length=n_fft, hop_length=hop).audio_data
    finally:
        torch.istft = real_istft
    y_ref = sp.istft(s.cpu(), 44100, 30000, n_fft, hop, "hann")
    assert y.shape == y_ref.shape and rel_err(y.cpu(), y_ref) < TOL


def test_arbitrary_window_lengths_match_reference(at):
    """GPU twin of tests/test_sim_signal_api.py::test_arbitrary_window_lengths_match_reference: the REAL reference's outputs."""
    import os

    from tests.golden import cases
    from tests.golden import make_golden_anywindow as mg

    g = np.load(os.path.join(os.path.dirname(mg.__file__), "reference_golden_anywindow.npz"))
    x = cases.make_input("cfg1")
    for key, wl, hop, wt, ms, pt in mg.STFT_CASES:
        sig = at.AudioSignal(x.clone(), 16000).to(DEV)
        X = sig.stft(window_length=wl, hop_length=hop, window_type=wt, match_stride=ms, padding_type=pt)
        ref = torch.from_numpy(g[key + "_stft"])
        assert X.shape[1:] == ref.shape[1:] and X.shape[0] == 4, key
        assert rel_err(torch.view_as_real(X[:2].cpu()), torch.view_as_real(ref)) < TOL, key
        assert elementwise_ok(X[:2].cpu().abs(), ref.abs()), key
        y = sig.istft(window_length=wl, hop_length=hop, window_type=wt, match_stride=ms).audio_data
        assert rel_err(y.cpu(), torch.from_numpy(g[key + "_istft"])) < TOL, key
    mel = at.AudioSignal(x.clone(), 16000).to(DEV).mel_spectrogram(n_mels=80, window_length=400, hop_length=160,
                                                                     window_type="hann")
    assert rel_err(mel.cpu(), torch.from_numpy(g["w400_mel80"])) < TOL
    mf = at.AudioSignal(x.clone(), 16000).to(DEV).mfcc(n_mfcc=20, n_mels=40, window_length=400, hop_length=160,
                                                        window_type="hann")
    assert rel_err(mf.cpu(), torch.from_numpy(g["w400_mfcc"])) < TOL

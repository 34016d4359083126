"""The C-ABI library loads and exports every symbol include/b2a.h declares (no compute calls:
this runs without a GPU); argument validation returns error codes instead of crashing."""
import ctypes
import os
import re

import numpy as np

import pytest

import __graft_entry__ as graft
from audiotools_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    graft.build()
    return _lib.B2ALibrary(_lib.LIB_PATH)


def declared_symbols():
    src = open(os.path.join(REPO, "include", "b2a.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2a_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    syms = declared_symbols()
    assert "b2a_spectral_f32" in syms and "b2a_lufs_f32" in syms
    assert sorted(_lib.SIGNATURES) == syms, "audiotools_b200/_lib.py must bind exactly what b2a.h declares"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert getattr(raw, s) is not None


def test_version_and_pure_host_queries(lib):
    assert lib.b2a_version() == 100
    # frame / block counts are integer host arithmetic (bit-exact requirement of BASELINE.json)
    assert lib.b2a_stft_num_frames(16000, 512, 128, 0, 0, 0) == 126          # cfg1
    assert lib.b2a_stft_num_frames(441000, 2048, 512, 0, 0, 0) == 862        # cfg2
    assert lib.b2a_stft_num_frames(16000, 256, 64, 96, 0, 2) == 250          # match_stride: T/hop
    assert lib.b2a_stft_num_frames(15999, 256, 64, 96, 1, 2) == 250
    assert lib.b2a_lufs_num_blocks(441000, 44100.0, 0.4) == 97               # cfg2
    assert lib.b2a_lufs_num_blocks(8000, 16000.0, 0.4) == 2
    assert lib.b2a_lufs_num_blocks(22050, 11025.0, 0.4) == 18                # K=4410, stride=1102
    assert lib.b2a_lufs_workspace_bytes(64, 2, 441000, 44100.0, 0.4) > 0


def test_bad_arguments_return_codes_not_crashes(lib):
    # null pointers / unsupported sizes are rejected before any CUDA call is made
    rc = lib.b2a_spectral_f32(None, 1, 100, 512, 128, None, 0, 0, 0, 0, None, 1, None, None, None, None, 0, 0, 0, 0.0,
                              1.0, None, None, None)
    assert rc == -1 and b"null" in lib.b2a_last_error()
    buf = (ctypes.c_float * 1024)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.b2a_spectral_f32(p, 1, 1024, 500, 128, p, 0, 0, 0, 0, None, 1, None, None, None, None, 0, 0, 0, 0.0, 1.0,
                              None, p, None)
    assert rc == -2 and b"power of two" in lib.b2a_last_error()
    rc = lib.b2a_spectral_f32(p, 1, 100, 512, 128, p, 0, 0, 0, 0, None, 1, None, None, None, None, 0, 0, 0, 0.0, 1.0,
                              None, p, None)
    assert rc == -1 and b"n_fft/2" in lib.b2a_last_error()
    with pytest.raises(_lib.B2AError):
        lib.check(rc)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.B2ALibrary(str(tmp_path / "libb2a.so"))


def test_engine_rejects_cpu_tensors(lib):
    import torch

    from audiotools_b200 import AudioSignal
    from audiotools_b200.engine import Engine

    eng = Engine(lib)  # product configuration: require_cuda=True
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng.lufs(torch.zeros(1, 1, 16000), 16000)
    sig = AudioSignal(torch.zeros(1, 1, 16000), 16000)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sig.stft()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sig.loudness()


def test_host_num_frames_equals_the_abi_function():
    from audiotools_b200.engine import Engine

    lib = _lib.get_lib()
    rng = np.random.RandomState(0)
    for _ in range(300):
        T, n_fft, hop = int(rng.randint(-2, 100000)), int(2 ** rng.randint(0, 13)), int(rng.randint(-1, 5000))
        pad, rp, de = int(rng.randint(-1, 2048)), int(rng.randint(-1, 4096)), int(rng.choice([0, 0, 2]))
        assert Engine.num_frames(T, n_fft, hop, pad, rp, de) == lib.b2a_stft_num_frames(T, n_fft, hop, pad, rp, de)

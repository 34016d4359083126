"""Tensor-level entry points over the C ABI of ``libb2a`` (``include/b2a.h``).

One ``Engine`` wraps one loaded library.  Inputs are float32 CUDA tensors; every
call is enqueued on torch's current stream of the tensor's device and returns
torch tensors that torch allocated (the library owns no buffers).  There is no
CPU path: a non-CUDA tensor raises.  (``require_cuda=False`` exists only so that
tests can drive the *same marshalling code* against ``tests/cusim``'s CPU build of
the kernels; the module-level engine is always built with ``require_cuda=True``.)
"""
import ctypes
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from .core import kweighting


def _dptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, lib: _lib.B2ALibrary, require_cuda: bool = True):
        self.lib = lib
        self.require_cuda = require_cuda
        self.launches = 0  # kernels of libb2a launched through this engine (bench.py reports it)
        self._packed_cache = {}

    # ------------------------------------------------------------------ helpers
    def _prep(self, t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
        if not torch.is_tensor(t):
            raise TypeError(f"{name} must be a torch.Tensor")
        if self.require_cuda and not t.is_cuda:
            raise RuntimeError(
                f"{name} is on {t.device}: audiotools_b200 runs on CUDA (sm_100a) only and has no CPU fallback")
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    def _stream(self, t: torch.Tensor):
        if t.is_cuda:
            return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
        return None

    # ------------------------------------------------------------------ loudness
    def lufs(self, x: torch.Tensor, sample_rate: float, filter_class: str = "K-weighting",
             block_size: float = 0.400, padded_length: Optional[int] = None,
             target_db: Optional[torch.Tensor] = None, want_blocks: bool = False):
        """Integrated loudness of ``x`` [B, C, T] (ref:audiotools/core/loudness.py:176-247, IIR path).

        Returns a dict: ``lufs`` [B] (unclamped), ``loud`` [B] (= max(lufs, -70)), and, when
        ``target_db`` (1 or B values) is given, ``gain`` [B] = exp((target_db - loud) ln10/20);
        ``blocks`` [B, C, nblk] when ``want_blocks``.
        """
        x = self._prep(x, "x")
        assert x.ndim == 3, "x must be [B, C, T]"
        B, C, T = x.shape
        Tp = T if padded_length is None else int(padded_length)
        sos, sgain = kweighting.design(float(sample_rate), filter_class)
        G = np.ascontiguousarray(kweighting.CHANNEL_GAINS[:C], dtype=np.float64)
        if C > len(kweighting.CHANNEL_GAINS):
            raise ValueError(f"loudness supports at most 5 channels, got {C}")
        L = self.lib
        nblk = L.b2a_lufs_num_blocks(Tp, float(sample_rate), float(block_size))
        ws_bytes = L.b2a_lufs_workspace_bytes(B, C, Tp, float(sample_rate), float(block_size))
        if nblk < 1 or ws_bytes == 0:
            raise _lib.B2AError(f"lufs: unsupported geometry (T={Tp}, rate={sample_rate}, block={block_size})")
        dev = x.device
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        lufs = torch.empty(B, dtype=torch.float32, device=dev)
        loud = torch.empty(B, dtype=torch.float32, device=dev)
        blocks = torch.empty(B, C, nblk, dtype=torch.float32, device=dev) if want_blocks else None
        gain = None
        n_target = 0
        if target_db is not None:
            target_db = self._prep(torch.as_tensor(target_db, device=dev).reshape(-1), "target_db")
            n_target = target_db.numel()
            if n_target not in (1, B):
                raise ValueError(f"target_db must have 1 or {B} entries, got {n_target}")
            gain = torch.empty(B, dtype=torch.float32, device=dev)
        dp = ctypes.POINTER(ctypes.c_double)
        rc = L.b2a_lufs_f32(_dptr(x), B, C, T, Tp, float(sample_rate),
                            sos.ctypes.data_as(dp), sgain.ctypes.data_as(dp), sos.shape[0], float(block_size),
                            G.ctypes.data_as(dp), _dptr(blocks), _dptr(lufs), _dptr(loud),
                            _dptr(target_db), n_target, _dptr(gain), _dptr(ws), ws_bytes, self._stream(x))
        L.check(rc)
        self.launches += 3  # lufs_setup, kweight_energy, lufs_gate (+ one memset node)
        return {"lufs": lufs, "loud": loud, "gain": gain, "blocks": blocks}

    def gain(self, x: torch.Tensor, gain: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``x[b] * gain[b]`` (ref:audiotools/core/effects.py:219,237)."""
        x = self._prep(x, "x")
        B = x.shape[0]
        gain = self._prep(gain.reshape(-1), "gain")
        assert gain.numel() == B
        if out is None:
            out = torch.empty_like(x)
        per_item = x.numel() // B
        self.lib.check(self.lib.b2a_gain_f32(_dptr(x), _dptr(out), B, per_item, _dptr(gain), self._stream(x)))
        self.launches += 1
        return out

    # ------------------------------------------------------------------ STFT / mel
    def _packed_len(self, mel_lo: torch.Tensor, mel_hi: torch.Tensor) -> int:
        """Sum of the 4-aligned band widths (host int; cached per band table)."""
        key = (mel_lo.data_ptr(), mel_hi.data_ptr(), mel_lo.numel())
        if key not in self._packed_cache:
            lo, hi = mel_lo.cpu().numpy().astype("int64"), mel_hi.cpu().numpy().astype("int64")
            self._packed_cache[key] = int((((hi + 3) & ~3) - (lo & ~3)).clip(min=0).sum())
        return self._packed_cache[key]

    def num_frames(self, T: int, n_fft: int, hop: int, pad: int = 0, right_pad: int = 0, drop_edge: int = 0) -> int:
        return int(self.lib.b2a_stft_num_frames(T, n_fft, hop, pad, right_pad, drop_edge))

    def spectral(self, x: torch.Tensor, n_fft: int, hop: int, window: torch.Tensor, pad: int = 0,
                 right_pad: int = 0, pad_mode: str = "reflect", drop_edge: int = 0,
                 gain: Optional[torch.Tensor] = None, want_scaled: bool = False,
                 mel_fb: Optional[torch.Tensor] = None, mel_lo: Optional[torch.Tensor] = None,
                 mel_hi: Optional[torch.Tensor] = None, post: int = _lib.POST_NONE, post_eps: float = 0.0,
                 post_power: float = 1.0, want_stft: bool = True):
        """Fused framing -> window -> rFFT -> (|.| -> banded mel -> post) over ``x`` [B, C, T].

        Returns dict(stft=[B,C,F,N] complex64 | None, mel=[B,C,n_mels,N] | None, scaled=[B,C,T] | None).
        """
        x = self._prep(x, "x")
        assert x.ndim == 3
        B, C, T = x.shape
        rows = B * C
        if pad_mode not in _lib.PAD_MODES:
            raise NotImplementedError(f"padding_type {pad_mode!r} (supported: {sorted(_lib.PAD_MODES)})")
        window = self._prep(window, "window")
        assert window.numel() == n_fft
        N = self.num_frames(T, n_fft, hop, pad, right_pad, drop_edge)
        if N < 1:
            raise _lib.B2AError(f"stft: no frames (T={T}, n_fft={n_fft}, hop={hop})")
        F = n_fft // 2 + 1
        dev = x.device
        stft = torch.empty(B, C, F, N, dtype=torch.complex64, device=dev) if want_stft else None
        mel = None
        n_mels = 0
        packed_len = 0
        if mel_fb is not None:
            mel_fb = self._prep(mel_fb, "mel_fb")
            n_mels = mel_fb.shape[0]
            assert mel_fb.shape[1] == F, (mel_fb.shape, F)
            mel_lo = self._prep(mel_lo, "mel_lo", torch.int32)
            mel_hi = self._prep(mel_hi, "mel_hi", torch.int32)
            mel = torch.empty(B, C, n_mels, N, dtype=torch.float32, device=dev)
            packed_len = self._packed_len(mel_lo, mel_hi)
        scaled = None
        rows_per_gain = 1
        if gain is not None:
            gain = self._prep(gain.reshape(-1), "gain")
            assert gain.numel() == B
            rows_per_gain = C
            if want_scaled:
                scaled = torch.empty_like(x)
        rc = self.lib.b2a_spectral_f32(
            _dptr(x), rows, T, n_fft, hop, _dptr(window), pad, right_pad, _lib.PAD_MODES[pad_mode], drop_edge,
            _dptr(gain), rows_per_gain, _dptr(scaled),
            _dptr(mel_fb), _dptr(mel_lo), _dptr(mel_hi), n_mels, packed_len, post, float(post_eps),
            float(post_power),
            _dptr(mel), _dptr(torch.view_as_real(stft)) if stft is not None else None, self._stream(x))
        self.lib.check(rc)
        self.launches += 1
        return {"stft": stft, "mel": mel, "scaled": scaled}


_ENGINE = None


def get_engine() -> Engine:
    """The product engine: the in-tree CUDA library, CUDA tensors only."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(_lib.get_lib(), require_cuda=True)
    return _ENGINE

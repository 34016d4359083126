"""Batched, seed-reproducible transforms with the API of ref:audiotools/data/transforms.py.

A transform has two halves.  ``instantiate(state, signal)`` draws its parameters on the host from
distribution tuples with a ``numpy.random.RandomState`` (so a seed reproduces them) and returns a
nested dict keyed by the transform's name; ``transform(signal, **kwargs)`` applies it to a whole
batch on the device, gated per item by the boolean ``mask`` drawn with probability ``prob``.
``Compose`` names its children ``"{position}.{Name}"`` and threads the same nested dict through
them; ``Choose`` turns the children's masks into a one-hot choice.

Every concrete transform is one ``AudioSignal`` method, i.e. one or two launches of ``libb2a``.
Transforms of the reference that need file-backed loaders (``BackgroundNoise``, ``CrossTalk``,
file-based ``RoomImpulseResponse``) take in-memory ``AudioSignal`` pools here instead: decoding
audio files is outside the accelerated hot path.
"""
import copy
from contextlib import contextmanager
from inspect import signature
from typing import List

import numpy as np
import torch
from numpy.random import RandomState

from ..core import AudioSignal
from ..core import util
from ..core.audio_signal import _on_engine

tt = torch.tensor
"""Shorthand for converting things to torch.tensor."""


# ------------------------------------------------------------------------------------------
# Parameter tables.  A transform draws its parameters as PLAIN host values (python / numpy scalars, small arrays,
# AudioSignals) -- the RNG calls and their order are the reference's, so a seed reproduces them -- and they only become
# tensors once per batch: ``batch_instantiate`` stacks the B draws of every parameter into ONE tensor per key
# (the reference builds B x n_keys zero-dim tensors and collates them: 250 ms per 512 items for a seven-transform
# Compose on this host, 30 ms here), ``instantiate`` tensorises a single draw.  ``util.prepare_batch`` then uploads each
# table once and keeps its host mirror, so mask / cutoff / shift decisions never synchronise with the device.
# ------------------------------------------------------------------------------------------
def _tensorize(value):
    """One draw -> the reference's ``instantiate`` output: every leaf a tensor (``torch.tensor(v)``)."""
    if isinstance(value, dict):
        return {k: _tensorize(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_tensorize(v) for v in value]
    if isinstance(value, (AudioSignal, torch.Tensor)):
        return value
    return tt(value)


def _stack_leaf(vals: list):
    """B draws of one parameter -> one [B, ...] tensor with the dtype ``default_collate([torch.tensor(v) ...])`` gives."""
    v0 = vals[0]
    if isinstance(v0, AudioSignal):
        return AudioSignal.batch(list(vals), pad_signals=True)
    if isinstance(v0, torch.Tensor):
        return torch.stack(list(vals))
    if isinstance(v0, (list, tuple)):  # e.g. Choose's one_hot: a list of per-child flags -> a list of [B] tensors
        return [_stack_leaf([v[i] for v in vals]) for i in range(len(v0))]
    dtype = tt(v0).dtype  # python float -> float32, python int -> int64, bool -> bool, numpy scalars / arrays keep theirs
    arr = np.asarray(vals)
    return torch.as_tensor(arr).to(dtype) if arr.dtype != object else torch.stack([tt(v) for v in vals])


def _collate_draws(draws: list):
    flats = [util.flatten(d) for d in draws]
    return util.unflatten({k: _stack_leaf([f[k] for f in flats]) for k in flats[0]})


class BaseTransform:
    def __init__(self, keys: list = [], name: str = None, prob: float = 1.0):
        # parameter names come from the _transform signature (everything but signal / kwargs)
        params = signature(self._transform).parameters
        tfm_keys = [k for k in params.keys() if k not in ("signal", "kwargs", "_bypass")]
        self.keys = keys + tfm_keys + ["mask"]
        # mask-aware: ``_transform(signal, ..., _bypass=[B] bool)`` runs on the WHOLE batch and leaves the flagged items
        # untouched inside the kernels (csrc: bypass flags / unit gains / zero shifts): no gather, no scatter
        # ... where that pays: `_bypass_pays = False` marks the FFT-convolution transforms, whose forward block FFTs run
        # for every row of the launch -- measured on a prob-0.5 chain (128 x 10 s), flags on all five transforms 7.58 ms
        # against 7.44 ms for the gather path, so those keep the gather (half the rows, two cheap copies)
        self._mask_aware = "_bypass" in params and getattr(self, "_bypass_pays", True)
        self.prob = prob
        self.name = self.__class__.__name__ if name is None else name
        self._needs_signal = "signal" in signature(self._instantiate).parameters  # inspected once, not per item

    def _prepare(self, batch: dict):
        sub_batch = batch[self.name]
        for k in self.keys:
            assert k in sub_batch.keys(), f"{k} not in batch"
        return sub_batch

    def _transform(self, signal):
        return signal

    def _instantiate(self, state: RandomState, signal: AudioSignal = None):
        return {}

    @staticmethod
    def apply_mask(batch: dict, mask: torch.Tensor):
        return util.unflatten({k: v[mask] for k, v in util.flatten(batch).items()})

    def transform(self, signal: AudioSignal, **kwargs):
        """``signal[mask] = self._transform(signal[mask], **kwargs[mask])`` (ref :133-166).  When every item is
        selected the gather / scatter copies of the whole batch are skipped and the transform runs on ``signal``
        itself, leaving the same state the masked round trip leaves (samples replaced; a loudness / STFT cache is
        only overwritten when both sides hold one, :1658-1679).  The mask is read from its host mirror
        (``util.prepare_batch``) when there is one: no device synchronisation."""
        tfm_kwargs = self._prepare(kwargs)
        mask = tfm_kwargs["mask"]
        host_mask = util.host_view(mask)
        n_sel = int(host_mask.sum())
        if n_sel == 0:
            return signal
        if host_mask.ndim == 1 and n_sel == host_mask.numel() == signal.batch_size:
            tfm_kwargs = {k: v for k, v in tfm_kwargs.items() if k != "mask"}
            pre_loud, pre_stft = signal._loudness, signal.stft_data
            out = self._transform(signal, **tfm_kwargs)
            if out is signal:
                new_loud, new_stft = signal._loudness, signal.stft_data
                signal._loudness = new_loud if (pre_loud is not None and new_loud is not None) else pre_loud
                signal.stft_data = new_stft if (pre_stft is not None and new_stft is not None) else pre_stft
                return signal
            signal[mask] = out
            return signal
        if self._mask_aware and host_mask.ndim == 1 and host_mask.numel() == signal.batch_size and \
                _on_engine(signal._audio_data) and self._bypass_ok(signal, tfm_kwargs):
            # the reference gathers signal[mask], transforms the copy and scatters it back (:133-166): two extra passes
            # over the selected items.  Here the kernels take the complement of the mask as per-item bypass flags.
            args = {k: v for k, v in tfm_kwargs.items() if k != "mask"}
            pre_loud, pre_stft = signal._loudness, signal.stft_data
            bypass = ~mask.to(signal.device).bool()
            if bypass.is_cuda:  # host mirror (the mask's own): host-side decisions on the flags need no sync
                bypass._b2a_host = (~host_mask.bool(), bypass._version)
            out = self._transform(signal, **args, _bypass=bypass)
            if out is signal:
                # state of `signal[mask] = out` (:1658-1679): caches are overwritten only where both sides hold one,
                # and only for the selected items
                new_loud, new_stft = signal._loudness, signal.stft_data
                sel = mask.to(signal.device).bool()
                if pre_loud is not None and new_loud is not None:
                    signal._loudness = torch.where(sel, new_loud, pre_loud)
                else:
                    signal._loudness = pre_loud
                if pre_stft is not None and new_stft is not None and pre_stft.shape == new_stft.shape:
                    signal.stft_data = torch.where(sel.reshape(-1, 1, 1, 1), new_stft, pre_stft)
                else:
                    signal.stft_data = pre_stft
                return signal
            signal[mask] = out[mask]
            return signal
        tfm_kwargs = self.apply_mask(tfm_kwargs, mask)
        tfm_kwargs = {k: v for k, v in tfm_kwargs.items() if k != "mask"}
        signal[mask] = self._transform(signal[mask], **tfm_kwargs)
        return signal

    def _bypass_ok(self, signal, tfm_kwargs) -> bool:
        """Per-call veto of the bypass-flag path (e.g. a filter long enough to go through the FFT engine)."""
        return True

    def __call__(self, *args, **kwargs):
        return self.transform(*args, **kwargs)

    def _draw(self, state: RandomState, signal: AudioSignal = None) -> dict:
        """One item's parameters as plain host values, keyed by the transform's name (mask last, as in the reference:
        ``state.rand() <= prob`` is drawn AFTER the parameters, ref :228-240)."""
        params = self._instantiate(state, signal) if self._needs_signal else self._instantiate(state)
        params["mask"] = bool(state.rand() <= self.prob)
        return {self.name: params}

    def instantiate(self, state: RandomState = None, signal: AudioSignal = None):
        return _tensorize(self._draw(util.random_state(state), signal))

    def batch_instantiate(self, states: list = None, signal: AudioSignal = None):
        """Parameters for a batch, one seed / state per item (ref :242-265): B draws, ONE tensor per parameter."""
        return _collate_draws([self._draw(util.random_state(state), signal) for state in states])


class Identity(BaseTransform):
    pass


class SpectralTransform(BaseTransform):
    """stft -> transform -> istft."""

    def transform(self, signal, **kwargs):
        signal.stft()
        super().transform(signal, **kwargs)
        signal.istft()
        return signal


class Compose(BaseTransform):
    def __init__(self, *transforms: list, name: str = None, prob: float = 1.0):
        if isinstance(transforms[0], list):
            transforms = transforms[0]
        for i, tfm in enumerate(transforms):
            tfm.name = f"{i}.{tfm.name}"
        keys = [tfm.name for tfm in transforms]
        super().__init__(keys=keys, name=name, prob=prob)
        self.transforms = transforms
        self.transforms_to_apply = keys

    @contextmanager
    def filter(self, *names: list):
        old = self.transforms_to_apply
        self.transforms_to_apply = names
        yield
        self.transforms_to_apply = old

    def _transform(self, signal, **kwargs):
        for transform in self.transforms:
            if any([x in transform.name for x in self.transforms_to_apply]):
                signal = transform(signal, **kwargs)
        return signal

    def _instantiate(self, state: RandomState, signal: AudioSignal = None):
        parameters = {}
        for transform in self.transforms:
            parameters.update(transform._draw(state, signal))
        return parameters

    def __getitem__(self, idx):
        return self.transforms[idx]

    def __len__(self):
        return len(self.transforms)

    def __iter__(self):
        for transform in self.transforms:
            yield transform


class Choose(Compose):
    def __init__(self, *transforms: list, weights: list = None, name: str = None, prob: float = 1.0):
        super().__init__(*transforms, name=name, prob=prob)
        if weights is None:
            n = len(self.transforms)
            weights = [1 / n for _ in range(n)]
        self.weights = np.array(weights)

    def _instantiate(self, state: RandomState, signal: AudioSignal = None):
        kwargs = super()._instantiate(state, signal)
        tfm_idx = state.choice(list(range(len(self.transforms))), p=self.weights)
        one_hot = []
        for i, t in enumerate(self.transforms):
            if kwargs[t.name]["mask"]:
                kwargs[t.name]["mask"] = bool(i == tfm_idx)
            one_hot.append(kwargs[t.name]["mask"])
        kwargs["one_hot"] = one_hot
        return kwargs


class Repeat(Compose):
    def __init__(self, transform, n_repeat: int = 1, name: str = None, prob: float = 1.0):
        super().__init__([copy.copy(transform) for _ in range(n_repeat)], name=name, prob=prob)
        self.n_repeat = n_repeat


class RepeatUpTo(Choose):
    def __init__(self, transform, max_repeat: int = 5, weights: list = None, name: str = None, prob: float = 1.0):
        transforms = [Repeat(transform, n_repeat=n) for n in range(1, max_repeat)]
        super().__init__(transforms, name=name, prob=prob, weights=weights)
        self.max_repeat = max_repeat


# ------------------------------------------------------------------------------------------
# concrete transforms (one AudioSignal method each)
# ------------------------------------------------------------------------------------------
class VolumeChange(BaseTransform):
    """``signal.volume_change(db)`` (ref :941-970)."""

    def __init__(self, db: tuple = ("uniform", -12.0, 0.0), name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.db = db

    def _instantiate(self, state: RandomState):
        return {"db": util.sample_from_dist(self.db, state)}

    def _transform(self, signal, db, _bypass=None):
        return signal.volume_change(db, _bypass=_bypass)


class VolumeNorm(BaseTransform):
    """``signal.normalize(db)`` -- LUFS normalisation (ref :973-1003)."""

    def __init__(self, db: tuple = ("const", -24), name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.db = db

    def _instantiate(self, state: RandomState):
        return {"db": util.sample_from_dist(self.db, state)}

    def _transform(self, signal, db, _bypass=None):
        return signal.normalize(db, _bypass=_bypass)


class GlobalVolumeNorm(BaseTransform):
    """Normalise using the loudness of the whole source file, read from
    ``signal.metadata["loudness"]`` (ref :1006-1050)."""

    def __init__(self, db: tuple = ("const", -24), name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.db = db

    def _instantiate(self, state: RandomState, signal: AudioSignal):
        if "loudness" not in signal.metadata:
            db_change = 0.0
        elif float(signal.metadata["loudness"]) == float("-inf"):
            db_change = 0.0
        else:
            db_change = util.sample_from_dist(self.db, state) - float(signal.metadata["loudness"])
        return {"db": db_change}

    def _transform(self, signal, db):
        return signal.volume_change(db)


class Equalizer(BaseTransform):
    """``signal.equalizer(eq)`` with ``eq = -eq_amount * rand(n_bands)`` (ref :564-600)."""
    _bypass_pays = False  # 641 taps: FFT convolution

    def __init__(self, eq_amount: tuple = ("const", 1.0), n_bands: int = 6, name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.eq_amount = eq_amount
        self.n_bands = n_bands

    def _instantiate(self, state: RandomState):
        eq_amount = util.sample_from_dist(self.eq_amount, state)
        return {"eq": -eq_amount * state.rand(self.n_bands)}

    def _transform(self, signal, eq, _bypass=None):
        return signal.equalizer(eq, _bypass=_bypass)


class LowPass(BaseTransform):
    """``signal.low_pass(cutoff, zeros)`` (ref :1095-1131)."""

    def __init__(self, cutoff: tuple = ("choice", [4000, 8000, 16000]), zeros: int = 51, name: str = None,
                 prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.cutoff = cutoff
        self.zeros = zeros

    def _instantiate(self, state: RandomState):
        return {"cutoff": util.sample_from_dist(self.cutoff, state)}

    def _transform(self, signal, cutoff, _bypass=None):
        return signal.low_pass(cutoff, zeros=self.zeros, _bypass=_bypass)

    def _bypass_ok(self, signal, tfm_kwargs) -> bool:
        # flags pay when the time-domain kernel serves the call (a flagged row is a plain copy); a bank long enough for
        # the FFT engine (> 320 taps) would spend its forward FFTs on the unselected rows as well: gather instead
        cut = util.host_view(tfm_kwargs["cutoff"]).reshape(-1).float()
        sel = util.host_view(tfm_kwargs["mask"]).reshape(-1).bool()
        if not bool(sel.any()) or float(cut[sel].min()) <= 0:
            return False
        return 2 * int(self.zeros / (float(cut[sel].min()) / signal.sample_rate) / 2) + 1 <= 320


class HighPass(BaseTransform):
    """``signal.high_pass(cutoff, zeros)`` (ref :1134-1170)."""

    def __init__(self, cutoff: tuple = ("choice", [50, 100, 250, 500, 1000]), zeros: int = 51, name: str = None,
                 prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.cutoff = cutoff
        self.zeros = zeros

    def _instantiate(self, state: RandomState):
        return {"cutoff": util.sample_from_dist(self.cutoff, state)}

    def _transform(self, signal, cutoff, _bypass=None):
        return signal.high_pass(cutoff, zeros=self.zeros, _bypass=_bypass)

    def _bypass_ok(self, signal, tfm_kwargs) -> bool:
        # flags pay when the time-domain kernel serves the call (a flagged row is a plain copy); a bank long enough for
        # the FFT engine (> 320 taps) would spend its forward FFTs on the unselected rows as well: gather instead
        cut = util.host_view(tfm_kwargs["cutoff"]).reshape(-1).float()
        sel = util.host_view(tfm_kwargs["mask"]).reshape(-1).bool()
        if not bool(sel.any()) or float(cut[sel].min()) <= 0:
            return False
        return 2 * int(self.zeros / (float(cut[sel].min()) / signal.sample_rate) / 2) + 1 <= 320


class _PoolTransform(BaseTransform):
    """Shared by the transforms that draw another signal: ``sources`` is an in-memory pool -- a list of ``AudioSignal``
    (any lengths, resampled to the target rate on the device if needed) or a callable ``(state, signal) ->
    AudioSignal`` -- instead of the reference's csv lists of audio files (file decoding is outside the hot path)."""

    def _init_pool(self, sources, weights):
        self.sources = sources
        self.weights = weights

    def _draw_excerpt(self, state: RandomState, signal: AudioSignal):
        """One pool item, cut / zero-padded to the duration of ``signal`` at a random offset, with its channels."""
        if callable(self.sources):
            return self.sources(state, signal)
        if not self.sources:
            raise ValueError(f"{type(self).__name__} needs `sources`: a list of AudioSignal or a callable")
        src = self.sources[state.choice(len(self.sources), p=self.weights)].clone()
        if src.sample_rate != signal.sample_rate:
            raise ValueError(f"pool item at {src.sample_rate} Hz, signal at {signal.sample_rate} Hz: resample the pool")
        n = signal.signal_length
        if src.signal_length > n:
            off = int(state.randint(0, src.signal_length - n + 1))
            src.audio_data = src.audio_data[..., off:off + n]
        else:
            src.zero_pad_to(n)
        if src.num_channels != signal.num_channels:  # mono pool item -> every channel (ref loader: num_channels=...)
            src.audio_data = src.audio_data[:, :1].expand(-1, signal.num_channels, -1).contiguous()
        return src


class NoiseFloor(BaseTransform):
    """Adds Gaussian noise at ``db`` LUFS (ref :669-704).  The reference normalises the noise on the CPU while
    instantiating; here the (seeded, host-drawn) noise is normalised by the LUFS kernel when the transform runs."""

    def __init__(self, db: tuple = ("const", -50.0), name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.db = db

    def _instantiate(self, state: RandomState, signal: AudioSignal):
        db = util.sample_from_dist(self.db, state)
        audio_data = state.randn(signal.num_channels, signal.signal_length)
        return {"nz_signal": AudioSignal(torch.from_numpy(audio_data).float()[None], signal.sample_rate), "db": db}

    def _transform(self, signal, nz_signal, db):
        return signal + nz_signal.clone().normalize(db)


class BackgroundNoise(_PoolTransform):
    """``signal.mix(bg_signal, snr, eq)`` (ref :707-792) with an in-memory pool of noise signals."""

    def __init__(self, snr: tuple = ("uniform", 10.0, 30.0), sources: List[AudioSignal] = None,
                 weights: List[float] = None, eq_amount: tuple = ("const", 1.0), n_bands: int = 3, name: str = None,
                 prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.snr = snr
        self.eq_amount = eq_amount
        self.n_bands = n_bands
        self._init_pool(sources, weights)

    def _instantiate(self, state: RandomState, signal: AudioSignal):
        eq_amount = util.sample_from_dist(self.eq_amount, state)
        eq = -eq_amount * state.rand(self.n_bands)
        snr = util.sample_from_dist(self.snr, state)
        return {"eq": eq, "bg_signal": self._draw_excerpt(state, signal), "snr": snr}

    def _transform(self, signal, bg_signal, snr, eq):
        return signal.mix(bg_signal.clone(), snr, eq)


class CrossTalk(_PoolTransform):
    """Mixes another talker in at ``snr`` and restores the original loudness (ref :795-854)."""

    def __init__(self, snr: tuple = ("uniform", 0.0, 10.0), sources: List[AudioSignal] = None,
                 weights: List[float] = None, name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.snr = snr
        self._init_pool(sources, weights)

    def _instantiate(self, state: RandomState, signal: AudioSignal):
        snr = util.sample_from_dist(self.snr, state)
        return {"crosstalk_signal": self._draw_excerpt(state, signal), "snr": snr}

    def _transform(self, signal, crosstalk_signal, snr):
        loudness = signal.loudness()
        mix = signal.mix(crosstalk_signal.clone(), snr)
        mix.normalize(loudness)
        return mix


class RoomImpulseResponse(BaseTransform):
    """``signal.apply_ir(ir, drr, eq)`` (ref :857-938).  ``sources`` is an in-memory pool: a list of
    single-item ``AudioSignal`` impulse responses (or a callable ``(state, signal) -> AudioSignal``);
    one is drawn per item with ``state.choice`` and zero-padded to one second like the reference."""
    _bypass_pays = False  # FFT convolution

    def __init__(self, drr: tuple = ("uniform", 0.0, 30.0), sources: List[AudioSignal] = None,
                 weights: List[float] = None, eq_amount: tuple = ("const", 1.0), n_bands: int = 6,
                 name: str = None, prob: float = 1.0, use_original_phase: bool = False):
        super().__init__(name=name, prob=prob)
        self.drr = drr
        self.eq_amount = eq_amount
        self.n_bands = n_bands
        self.use_original_phase = use_original_phase
        self.sources = sources
        self.weights = weights

    def _draw_ir(self, state, signal):
        if callable(self.sources):
            return self.sources(state, signal)
        if not self.sources:
            raise ValueError("RoomImpulseResponse needs `sources`: a list of AudioSignal impulse responses")
        idx = state.choice(len(self.sources), p=self.weights)
        return self.sources[idx].clone()

    def _instantiate(self, state: RandomState, signal: AudioSignal = None):
        eq_amount = util.sample_from_dist(self.eq_amount, state)
        eq = -eq_amount * state.rand(self.n_bands)
        drr = util.sample_from_dist(self.drr, state)
        ir_signal = self._draw_ir(state, signal)
        ir_signal.zero_pad_to(signal.sample_rate)
        return {"eq": eq, "ir_signal": ir_signal, "drr": drr}

    def _transform(self, signal, ir_signal, drr, eq, _bypass=None):
        return signal.apply_ir(ir_signal.clone(), drr, eq, use_original_phase=self.use_original_phase, _bypass=_bypass)


class PitchShift(BaseTransform):
    """``signal.pitch_shift(n_semitones)``.  NEW: the reference has no PitchShift transform (only the
    ``AudioSignal.pitch_shift`` method, ref:audiotools/core/effects.py:247-277, which takes ONE shift
    for the whole batch); BASELINE.json config 4 needs it.  Items are grouped by their drawn shift inside the
    engine (one set of launches for the whole batch)."""

    def __init__(self, n_semitones: tuple = ("choice", [-2, -1, 0, 1, 2]), quick: bool = True, name: str = None,
                 prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.n_semitones = n_semitones
        self.quick = quick

    def _instantiate(self, state: RandomState):
        return {"n_semitones": util.sample_from_dist(self.n_semitones, state)}

    def _transform(self, signal, n_semitones, _bypass=None):
        # the grouping by shift is a host decision (host mirror: no sync); all groups share the kernel launches
        shifts = util.ensure_tensor(util.host_view(n_semitones), 1, signal.batch_size).reshape(-1)
        if _bypass is not None:  # shift 0 = the kernels copy the row through
            shifts = torch.where(util.host_view(_bypass).cpu().bool().reshape(-1), torch.zeros_like(shifts), shifts)
        return signal.pitch_shift(shifts, quick=self.quick)


class ClippingDistortion(BaseTransform):
    def __init__(self, perc: tuple = ("uniform", 0.0, 0.1), name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.perc = perc

    def _instantiate(self, state: RandomState):
        return {"perc": util.sample_from_dist(self.perc, state)}

    def _transform(self, signal, perc):
        return signal.clip_distortion(perc)


class Quantization(BaseTransform):
    def __init__(self, channels: tuple = ("choice", [8, 32, 128, 256, 1024]), name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.channels = channels

    def _instantiate(self, state: RandomState):
        return {"channels": util.sample_from_dist(self.channels, state)}

    def _transform(self, signal, channels):
        return signal.quantization(channels)


class MuLawQuantization(BaseTransform):
    def __init__(self, channels: tuple = ("choice", [8, 32, 128, 256, 1024]), name: str = None, prob: float = 1.0):
        super().__init__(name=name, prob=prob)
        self.channels = channels

    def _instantiate(self, state: RandomState):
        return {"channels": util.sample_from_dist(self.channels, state)}

    def _transform(self, signal, channels):
        return signal.mulaw_quantization(channels)


class RescaleAudio(BaseTransform):
    def __init__(self, val: float = 1.0, name: str = None, prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.val = val

    def _transform(self, signal):
        return signal.ensure_max_of_audio(self.val)


class ShiftPhase(SpectralTransform):
    """stft -> ``phase += shift`` -> istft (ref :1200-1229)."""

    def __init__(self, shift: tuple = ("uniform", -np.pi, np.pi), name: str = None, prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.shift = shift

    def _instantiate(self, state: RandomState):
        return {"shift": util.sample_from_dist(self.shift, state)}

    def _transform(self, signal, shift):
        return signal.shift_phase(shift)


class InvertPhase(ShiftPhase):
    """Phase shift by pi (ref :1232-1247)."""

    def __init__(self, name: str = None, prob: float = 1):
        super().__init__(shift=("const", np.pi), name=name, prob=prob)


class CorruptPhase(SpectralTransform):
    """Adds host-drawn (seeded) Gaussian noise to the phase (ref :1250-1278)."""

    def __init__(self, scale: tuple = ("uniform", 0, np.pi), name: str = None, prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.scale = scale

    def _instantiate(self, state: RandomState, signal: AudioSignal = None):
        scale = util.sample_from_dist(self.scale, state)
        # the shape of one item's phase, without touching the device: [C, F, N]
        wl, hop, _, match_stride, _ = signal._resolve_stft(None, None, None, None, None)
        right_pad, pad = signal.compute_stft_padding(wl, hop, match_stride)
        if signal.stft_data is not None:
            n_frames = signal.stft_data.shape[-1]
        else:
            from ..engine import get_engine

            n_frames = get_engine().num_frames(signal.signal_length, wl, hop, pad, right_pad, 2 if match_stride else 0)
        shape = (signal.num_channels, wl // 2 + 1, n_frames)
        corruption = state.normal(scale=scale, size=shape)
        return {"corruption": corruption.astype("float32")}

    def _transform(self, signal, corruption):
        return signal.shift_phase(shift=corruption)


class FrequencyMask(SpectralTransform):
    """Zeroes a frequency band around a drawn centre (SpecAugment; ref :1281-1324)."""

    def __init__(self, f_center: tuple = ("uniform", 0.0, 1.0), f_width: tuple = ("const", 0.1), name: str = None,
                 prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.f_center = f_center
        self.f_width = f_width

    def _instantiate(self, state: RandomState, signal: AudioSignal):
        f_center = util.sample_from_dist(self.f_center, state)
        f_width = util.sample_from_dist(self.f_width, state)
        fmin = max(f_center - (f_width / 2), 0.0)
        fmax = min(f_center + (f_width / 2), 1.0)
        return {"fmin_hz": (signal.sample_rate / 2) * fmin, "fmax_hz": (signal.sample_rate / 2) * fmax}

    def _transform(self, signal, fmin_hz: float, fmax_hz: float):
        return signal.mask_frequencies(fmin_hz=fmin_hz, fmax_hz=fmax_hz)


class TimeMask(SpectralTransform):
    """Zeroes a span of frames around a drawn centre (SpecAugment; ref :1327-1369)."""

    def __init__(self, t_center: tuple = ("uniform", 0.0, 1.0), t_width: tuple = ("const", 0.025), name: str = None,
                 prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.t_center = t_center
        self.t_width = t_width

    def _instantiate(self, state: RandomState, signal: AudioSignal):
        t_center = util.sample_from_dist(self.t_center, state)
        t_width = util.sample_from_dist(self.t_width, state)
        tmin = max(t_center - (t_width / 2), 0.0)
        tmax = min(t_center + (t_width / 2), 1.0)
        return {"tmin_s": signal.signal_duration * tmin, "tmax_s": signal.signal_duration * tmax}

    def _transform(self, signal, tmin_s: float, tmax_s: float):
        return signal.mask_timesteps(tmin_s=tmin_s, tmax_s=tmax_s)


class MaskLowMagnitudes(SpectralTransform):
    """Zeroes STFT cells below a drawn dB threshold (ref :1372-1402)."""

    def __init__(self, db_cutoff: tuple = ("uniform", -10, 10), name: str = None, prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.db_cutoff = db_cutoff

    def _instantiate(self, state: RandomState, signal: AudioSignal = None):
        return {"db_cutoff": util.sample_from_dist(self.db_cutoff, state)}

    def _transform(self, signal, db_cutoff: float):
        return signal.mask_low_magnitudes(db_cutoff)


class Smoothing(BaseTransform):
    """Convolves the signal with a drawn window and restores its peak (ref :1405-1453)."""

    def __init__(self, window_type: tuple = ("const", "average"),
                 window_length: tuple = ("choice", [8, 16, 32, 64, 128, 256, 512]), name: str = None, prob: float = 1):
        super().__init__(name=name, prob=prob)
        self.window_type = window_type
        self.window_length = window_length

    def _instantiate(self, state: RandomState, signal: AudioSignal = None):
        window_type = util.sample_from_dist(self.window_type, state)
        window_length = util.sample_from_dist(self.window_length, state)
        window = signal.get_window(window_type=window_type, window_length=window_length, device="cpu")
        return {"window": AudioSignal(window.clone(), signal.sample_rate)}

    def _transform(self, signal, window):
        sscale = signal.audio_data.abs().max(dim=-1, keepdim=True).values
        sscale = torch.where(sscale == 0.0, torch.ones_like(sscale), sscale)
        out = signal.convolve(window)
        oscale = out.audio_data.abs().max(dim=-1, keepdim=True).values
        oscale = torch.where(oscale == 0.0, torch.ones_like(oscale), oscale)
        return out * (sscale / oscale)


def _refill_masked_cells(signal):
    """Replace the cells a band mask zeroed (|X| == 0 and angle == 0) by N(0,1) magnitude / phase noise drawn on
    the device (ref :1485-1495 / :1526-1536)."""
    mag, phase = signal.magnitude, signal.phase
    mag_r, phase_r = torch.randn_like(mag), torch.randn_like(phase)
    mask = (mag == 0.0) & (phase == 0.0)
    # the reference assigns `signal.magnitude = mag` and then `signal.phase = phase`; the second setter re-reads
    # |stft_data|, so a refilled cell ends up as |mag_r| exp(1j phase_r)
    signal.stft_data = torch.where(mask, mag_r.abs(), mag) * torch.exp(1j * torch.where(mask, phase_r, phase))
    return signal


class TimeNoise(TimeMask):
    """TimeMask, then the masked frames are filled with noise (ref :1456-1495)."""

    def _transform(self, signal, tmin_s: float, tmax_s: float):
        return _refill_masked_cells(signal.mask_timesteps(tmin_s=tmin_s, tmax_s=tmax_s, val=0.0))


class FrequencyNoise(FrequencyMask):
    """FrequencyMask, then the masked band is filled with noise (ref :1498-1536)."""

    def _transform(self, signal, fmin_hz: float, fmax_hz: float):
        return _refill_masked_cells(signal.mask_frequencies(fmin_hz=fmin_hz, fmax_hz=fmax_hz))


class Silence(BaseTransform):
    """Replace the item by silence while KEEPING its cached loudness (ref :1053-1092)."""

    def __init__(self, name: str = None, prob: float = 0.1):
        super().__init__(name=name, prob=prob)

    def _transform(self, signal):
        _loudness = signal._loudness
        signal = AudioSignal(torch.zeros_like(signal.audio_data), sample_rate=signal.sample_rate,
                             stft_params=signal.stft_params)
        signal._loudness = _loudness  # so that the target still can be normalised relative to it
        return signal


class SpectralDenoising(Equalizer):
    """Spectral gating against a host-drawn (seeded) white-noise excerpt, normalised and equalised on the device
    (ref :1539-1592; ``ml.layers.SpectralGate``)."""

    def __init__(self, eq_amount: tuple = ("const", 1.0), denoise_amount: tuple = ("uniform", 0.8, 1.0),
                 nz_volume: float = -40, n_bands: int = 6, n_freq: int = 3, n_time: int = 5, name: str = None,
                 prob: float = 1):
        super().__init__(eq_amount=eq_amount, n_bands=n_bands, name=name, prob=prob)
        from ..ml.layers import SpectralGate

        self.nz_volume = nz_volume
        self.denoise_amount = denoise_amount
        self.spectral_gate = SpectralGate(n_freq, n_time)

    def _transform(self, signal, nz, eq, denoise_amount):
        nz = nz.normalize(self.nz_volume).equalizer(eq)
        self.spectral_gate = self.spectral_gate.to(signal.device)
        return self.spectral_gate(signal, nz, denoise_amount)

    def _instantiate(self, state: RandomState):
        kwargs = super()._instantiate(state)
        kwargs["denoise_amount"] = util.sample_from_dist(self.denoise_amount, state)
        kwargs["nz"] = AudioSignal(torch.from_numpy(state.randn(22050)).float(), 44100)
        return kwargs

"""Host-side glue of the hot path: argument normalisation, seeded parameter sampling,
nested-dict batching.  Mirrors the names and semantics of ``ref:audiotools/core/util.py``
(ensure_tensor :56-89, random_state :129-160, seed :163-188, prepare_batch :346-380,
sample_from_dist :383-423, collate :426-479).  File discovery / plotting helpers of the
reference are out of scope."""
import math
import numbers
import random
import contextlib
import typing

import numpy as np
import torch


def flatten(d: dict, _parent: tuple = ()) -> dict:
    """Nested dict -> {tuple_path: leaf}  (what the reference gets from ``flatten_dict``)."""
    out = {}
    for k, v in d.items():
        if isinstance(v, dict) and v:
            out.update(flatten(v, _parent + (k,)))
        else:
            out[_parent + (k,)] = v
    return out


def unflatten(d: dict) -> dict:
    out = {}
    for path, v in d.items():
        cur = out
        for k in path[:-1]:
            cur = cur.setdefault(k, {})
        cur[path[-1]] = v
    return out


def ensure_tensor(x, ndim: int = None, batch_size: int = None) -> torch.Tensor:
    """Scalar / array / tensor -> tensor with ``ndim`` dims (trailing singleton dims are
    appended) and a leading dim expanded to ``batch_size``."""
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    if ndim is not None:
        assert x.ndim <= ndim
        while x.ndim < ndim:
            x = x.unsqueeze(-1)
    if batch_size is not None and x.shape[0] != batch_size:
        shape = list(x.shape)
        shape[0] = batch_size
        x = x.expand(*shape)
    return x


def _get_value(other):
    from .audio_signal import AudioSignal

    return other.audio_data if isinstance(other, AudioSignal) else other


def random_state(seed: typing.Union[int, np.random.RandomState, None]):
    if isinstance(seed, np.random.RandomState):  # the hot case (every transform of a Compose passes its state on)
        return seed
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, (numbers.Integral, np.integer, int)):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError("%r cannot be used to seed a numpy.random.RandomState instance" % seed)


def seed(random_seed, set_cudnn: bool = False):
    torch.manual_seed(random_seed)
    np.random.seed(random_seed)
    random.seed(random_seed)
    if set_cudnn:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


def sample_from_dist(dist_tuple: tuple, state: np.random.RandomState = None):
    """("const", v) | ("uniform", lo, hi) | ("choice", [...]) | any RandomState method."""
    if dist_tuple[0] == "const":
        return dist_tuple[1]
    state = random_state(state)
    return getattr(state, dist_tuple[0])(*dist_tuple[1:])


_HOST_MIRROR_MAX = 1 << 16


def _to_device(v, device):
    """``v.to(device)``; a small CPU parameter tensor moved to an accelerator keeps a reference to its host
    original (``_b2a_host``, together with the device tensor's version counter), so that host-side decisions on it
    (mask all-true?, which pitch shifts?, cutoff range checks) need no device->host synchronisation later (see
    :func:`host_view`)."""
    out = v.to(device)
    if torch.is_tensor(v) and out is not v and not v.is_cuda and out.is_cuda and v.numel() <= _HOST_MIRROR_MAX:
        out._b2a_host = (v, out._version)
    return out


def host_view(t):
    """The values of tensor ``t`` on the host: ``t`` itself on CPU, the mirror recorded by :func:`prepare_batch`
    when there is one AND the device tensor has not been written in place since (its version counter is unchanged),
    else a (synchronising) copy."""
    if not torch.is_tensor(t) or not t.is_cuda:
        return t
    mirror = getattr(t, "_b2a_host", None)
    if mirror is not None and mirror[1] == t._version and mirror[0].shape == t.shape:
        return mirror[0]
    return t.cpu()


def prepare_batch(batch, device="cpu"):
    """Move every tensor / AudioSignal of a (nested) batch to ``device`` -- the
    host->device boundary of the training loop."""
    if isinstance(batch, dict):
        flat = flatten(batch)
        for k, v in flat.items():
            try:
                flat[k] = _to_device(v, device)
            except Exception:
                pass
        return unflatten(flat)
    if torch.is_tensor(batch):
        return _to_device(batch, device)
    if isinstance(batch, list):
        for i in range(len(batch)):
            try:
                batch[i] = _to_device(batch[i], device)
            except Exception:
                pass
    return batch


def collate(list_of_dicts: list, n_splits: int = None):
    """List of (nested) dicts -> dict of batched values; AudioSignals are batched with
    ``AudioSignal.batch(pad_signals=True)``, everything else with torch's default collate."""
    from .audio_signal import AudioSignal

    batches = []
    list_len = len(list_of_dicts)
    return_list = n_splits is not None
    n_splits = 1 if n_splits is None else n_splits
    n_items = int(math.ceil(list_len / n_splits))
    for i in range(0, list_len, n_items):
        flat = [flatten(d) for d in list_of_dicts[i: i + n_items]]
        batch = {}
        for k in flat[0]:
            v = [d[k] for d in flat]
            if all(isinstance(s, AudioSignal) for s in v):
                batch[k] = AudioSignal.batch(v, pad_signals=True)
            else:
                batch[k] = torch.utils.data._utils.collate.default_collate(v)
        batches.append(unflatten(batch))
    return batches if return_list else batches[0]


def hz_to_bin(hz: torch.Tensor, n_fft: int, sample_rate: int) -> torch.Tensor:
    """Closest bin of a ``2 + n_fft // 2``-point grid over ``[0, sample_rate / 2]`` for every frequency
    (ref:audiotools/core/util.py:100-126; frequencies above Nyquist are clamped, as there)."""
    shape = hz.shape
    flat = hz.flatten().clamp(max=sample_rate / 2)
    freqs = torch.linspace(0, sample_rate / 2, 2 + n_fft // 2)
    return (flat[None, :] - freqs[:, None]).abs().argmin(dim=0).reshape(*shape)


def choose_from_list_of_lists(state: np.random.RandomState, list_of_lists: list, p: float = None):
    """Pick a list (with probabilities ``p``), then one of its items uniformly: ``(item, list index, item index)``
    (ref:audiotools/core/util.py:302-324)."""
    source_idx = state.choice(list(range(len(list_of_lists))), p=p)
    item_idx = state.randint(len(list_of_lists[source_idx]))
    return list_of_lists[source_idx][item_idx], source_idx, item_idx


@contextlib.contextmanager
def chdir(newdir):
    """Temporarily change the working directory (ref:audiotools/core/util.py:327-343)."""
    import os

    curdir = os.getcwd()
    try:
        os.chdir(newdir)
        yield
    finally:
        os.chdir(curdir)

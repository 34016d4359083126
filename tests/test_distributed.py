"""The N>1 path on CPU: world_size-2 ``gloo`` processes exercise the batch sharding and the loudness
all-gather exactly as ``bench.py`` / a training loop would with NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiotools_b200 import AudioSignal
from audiotools_b200.parallel import LoudnessGather, shard_batch, shard_bounds


def test_shard_bounds_cover_the_batch():
    for B in (1, 2, 7, 64, 2048):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(2048, 3, 8) == (768, 1024)  # BASELINE configs[4]: 256 clips per GPU


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 6
        x = torch.arange(B * 2 * 10, dtype=torch.float32).reshape(B, 2, 10)
        sig = AudioSignal(x.clone(), 16000)
        mine = shard_batch(sig)  # AudioSignal slice
        lo, hi = shard_bounds(B, rank, world)
        assert mine.batch_size == hi - lo and torch.equal(mine.audio_data, x[lo:hi])
        assert torch.equal(shard_batch(x), x[lo:hi])
        # stand-in for the per-item LUFS of the local shard (the kernels need a GPU; the exchange does not)
        loud_local = x[lo:hi].mean(dim=(1, 2))
        gathered = LoudnessGather()(loud_local)
        assert torch.equal(gathered, x.mean(dim=(1, 2)))  # rank order == batch order
        uneven = LoudnessGather()(x[: 1 + rank].mean(dim=(1, 2)), counts=[1, 2])
        assert uneven.numel() == 3
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def _peer_worker(rank, world, port, q, fail_rank):
    """Set-up protocol of PeerLoudnessExchange with the CPU-simulated library: both ranks must come out the same way
    (mapped, or raising together) -- a one-sided failure must never leave the other rank waiting in a collective."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from audiotools_b200 import _lib
        from audiotools_b200.parallel import PeerLoudnessExchange
        from tests.cusim import build_sim

        lib = _lib.B2ALibrary(build_sim.build())
        if rank == fail_rank:  # this rank cannot create its buffer (stands for: no cudaIpc, other node, ...)
            lib.b2a_peer_buffer_create = lambda *a: 1
        try:
            ex = PeerLoudnessExchange(n_max=8, device="cpu", lib=lib)
            ex.close()
            q.put((rank, "mapped"))
        except RuntimeError as e:
            q.put((rank, "raised" if "peer exchange unavailable" in str(e) else repr(e)))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("fail_rank,expect", [(-1, "mapped"), (1, "raised"), (0, "raised")])
def test_peer_exchange_setup_is_all_or_nothing(fail_rank, expect):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q, fail_rank)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(0, expect), (1, expect)], results

"""Two-process, two-GPU test of the NVLink peer-memory loudness exchange (audiotools_b200/parallel.py
PeerLoudnessExchange over csrc/peer.cu) against an NCCL all-gather of the same vectors.  Needs >= 2 GPUs:
skipped on the single-GPU box (`gpurun --gpus 2 -- python -m pytest tests/test_peer_exchange_gpu.py -m gpu`)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import __graft_entry__ as graft

    graft.build()
    from audiotools_b200 import AudioSignal
    from audiotools_b200.parallel import PeerLoudnessExchange

    ok = True
    ex = PeerLoudnessExchange(n_max=64)
    g = torch.Generator().manual_seed(100 + rank)
    x = (0.1 * torch.randn(8, 2, 44100, generator=g)).cuda()
    pending = None
    for step in range(6):
        loud = AudioSignal(x * (1 + 0.1 * step), 44100).loudness().contiguous()
        ref = torch.empty(world * loud.numel(), device=loud.device)
        dist.all_gather_into_tensor(ref, loud)
        if pending is not None:  # consume one step late, as bench.py does
            seq_prev, ref_prev = pending
            ok &= torch.equal(ex.collect(seq_prev), ref_prev)
        pending = (ex.put(loud), ref)
    ok &= torch.equal(ex.collect(pending[0]), pending[1])
    refs = []
    for step in range(5):  # the fused one-launch form: put(seq) + collect(seq - 1)
        loud = AudioSignal(x * (1 + 0.05 * step), 44100).loudness().contiguous()
        ref = torch.empty(world * loud.numel(), device=loud.device)
        dist.all_gather_into_tensor(ref, loud)
        refs.append(ref)
        seq, prev = ex.put_collect(loud)
        ok &= (prev is None) == (step == 0)
        if prev is not None:
            ok &= torch.equal(prev, refs[step - 1])
    ok &= torch.equal(ex.collect(seq), refs[-1])
    torch.cuda.synchronize()
    ex.close()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


@pytest.mark.gpu
def test_peer_exchange_matches_nccl_all_gather():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}

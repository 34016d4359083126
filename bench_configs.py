#!/usr/bin/env python
"""bench_configs.py -- auxiliary measurements of the OTHER BASELINE.json configs (the contract bench is
bench.py, which measures configs[1]).  One JSON line per config: device-resident throughput (CUDA events,
>= 3 warm-ups, inputs larger than L2 or rotated), algorithmic bytes, and the CPU oracle on a bounded sample.

    python bench_configs.py [--only cfg1,cfg3,cfg4,cfg5,istft,specaug,dense,gate,masked] [--no-cpu]

Multi-GPU (BASELINE configs[3] = 512 items on 4 GPUs, configs[4] = 2048 items on 8 GPUs): one process per GPU,
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        bench_configs.py --gpus 4 --only cfg4
every rank owns its slice of the batch (128 / 256 items, seeded by rank; no data-path collective), the timed region
is bracketed by a barrier + synchronize, the step time is the MAX over ranks and rank 0 prints the whole-job line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def _barrier():
    if WORLD > 1:
        import torch.distributed as dist

        dist.barrier()


def timed(fn, warmup=3, steps=10):
    """ms per step: CUDA events on the launching stream, >= 3 warm-ups unless stated, barrier + synchronize on both
    sides and the MAX over ranks when launched under torchrun."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    _barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    _barrier()
    ms = e0.elapsed_time(e1) / steps
    if WORLD > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device=f"cuda:{LOCAL}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def emit(line):
    if RANK == 0:
        line["n_gpus"] = WORLD
        print(json.dumps(line), flush=True)


def cpu_time(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg1,cfg3,cfg4,cfg5,istft,specaug,dense,gate,masked")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--gpus", type=int, default=1, help="ranks (informational: the launcher sets WORLD_SIZE)")
    args = ap.parse_args()
    import __graft_entry__ as graft

    graft.build()
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{LOCAL}"))
        args.no_cpu = True
    from audiotools_b200 import AudioSignal
    from audiotools_b200.data import transforms as tfm
    from oracle import signal_path as sp

    dev = f"cuda:{LOCAL}"
    peak = 6576.1
    pp = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    torch.set_num_threads(os.cpu_count() or 1)
    only = set(args.only.split(","))

    if "cfg1" in only:  # batch=4 mono 1s@16kHz stft(512,128): launch-latency bound
        x = torch.randn(4, 1, 16000, generator=torch.Generator().manual_seed(0))
        sig = AudioSignal(x.clone(), 16000).to(dev)
        ms = timed(lambda: sig.stft(window_length=512, hop_length=128), steps=50)
        line = {"config": "cfg1 batch=4 mono 1s@16k stft(512,128)", "ms": ms, "clips_per_s": 4 / ms * 1e3,
                "alg_bytes": 64000 + 1036224}
        if not args.no_cpu:
            line["cpu_ms"] = 1e3 * cpu_time(lambda: sp.stft(x, 16000, 512, 128), reps=20)
        emit(line)

    if "cfg3" in only:  # batch=256 mono 30s@48k -> 16k polyphase resample + low_pass(8k)
        B = 256
        g = torch.Generator().manual_seed(0)
        x = (0.1 * torch.randn(B, 1, 1440000, generator=g)).to(dev)  # 1.47 GB > L2

        def run():
            s = AudioSignal(x, 48000)
            s.resample(16000)
            s.low_pass(8000)
            return s

        ms_rs = timed(lambda: AudioSignal(x, 48000).resample(16000), steps=5)
        ms = timed(run, steps=5)
        alg = B * (5760000 + 1920000) + B * 2 * 1920000
        line = {"config": "cfg3 batch=256 mono 30s@48k resample->16k + low_pass(8k)", "ms": ms, "ms_resample": ms_rs,
                "clips_per_s": B / ms * 1e3, "alg_bytes": alg, "achieved_GBps": alg / ms / 1e6,
                "frac_of_hbm_peak": alg / ms / 1e6 / peak}
        if not args.no_cpu:
            xc = x[:8].cpu()
            t = cpu_time(lambda: sp.low_pass(sp.resample(xc, 48000, 16000), 16000, 8000), reps=1)
            line["cpu_clips_per_s"] = 8 / t
        emit(line)
        del x

    if "cfg4" in only:  # batch=512 Compose[EQ + IR-convolve + pitch_shift +-2], mono 10s@44.1k (one GPU's share: 128)
        B, T, sr = 128, 441000, 44100
        g = torch.Generator().manual_seed(RANK)
        x = 0.1 * torch.randn(B, 1, T, generator=g)
        t = torch.arange(sr) / sr
        irs = []
        for i in range(8):
            h = torch.randn(1, 1, sr, generator=g) * torch.exp(-t / 0.3) * 0.1
            h[..., 40 + i] = 1.0
            irs.append(AudioSignal(h, sr))
        transform = tfm.Compose([tfm.Equalizer(), tfm.RoomImpulseResponse(sources=irs),
                                 tfm.PitchShift(("choice", [-2, 2]))])
        sig = AudioSignal(x, sr)
        kwargs = transform.batch_instantiate(list(range(RANK * B, (RANK + 1) * B)), sig)
        sig = sig.to(dev)
        from audiotools_b200 import util

        kwargs = util.prepare_batch(kwargs, dev)
        ms = timed(lambda: transform(sig.clone(), **kwargs), warmup=3, steps=5)
        per = {}
        for name, fn in [("equalizer", lambda: sig.clone().equalizer(kwargs["Compose"]["0.Equalizer"]["eq"])),
                         ("apply_ir", lambda: sig.clone().apply_ir(kwargs["Compose"]["1.RoomImpulseResponse"]["ir_signal"].clone(),
                                                                   kwargs["Compose"]["1.RoomImpulseResponse"]["drr"],
                                                                   kwargs["Compose"]["1.RoomImpulseResponse"]["eq"])),
                         ("pitch_shift", lambda: sig.clone().pitch_shift(2))]:
            per[name] = timed(fn, warmup=1, steps=3)
        line = {"config": f"cfg4 batch={B * WORLD} ({B} per GPU) mono 10s@44.1k Compose[EQ+RoomIR+PitchShift+-2]", "ms": ms,
                "clips_per_s": B * WORLD / ms * 1e3, "per_gpu_batch": B, "global_batch": B * WORLD, "ms_parts": per,
                "timing": "CUDA events, barrier + synchronize both sides, max over ranks; no data-path collective"}
        emit(line)

    if "cfg5" in only:  # batch=2048 2ch 10s@44.1k full augment + LUFS + log-mel on 8 GPUs: one GPU's share (256 items)
        B, T, sr = 256, 441000, 44100
        g = torch.Generator().manual_seed(100 + RANK)
        x = 0.1 * torch.randn(B, 2, T, generator=g)
        t = torch.arange(sr) / sr
        irs = []
        for i in range(8):
            h = torch.randn(1, 1, sr, generator=g) * torch.exp(-t / 0.3) * 0.1
            h[..., 40 + i] = 1.0
            irs.append(AudioSignal(h, sr))
        transform = tfm.Compose([tfm.Equalizer(), tfm.RoomImpulseResponse(sources=irs),
                                 tfm.PitchShift(("choice", [-2, 2]))])
        sig = AudioSignal(x, sr)
        kwargs = transform.batch_instantiate(list(range(RANK * B, (RANK + 1) * B)), sig)
        sig = sig.to(dev)
        from audiotools_b200 import util

        kwargs = util.prepare_batch(kwargs, dev)

        def full():
            s = transform(sig.clone(), **kwargs)
            s.normalize(-24.0)
            return s.mel_spectrogram(n_mels=128, window_length=2048, hop_length=512, log=True)

        ms = timed(full, warmup=3, steps=5)
        emit({"config": f"cfg5 batch={B * WORLD} ({B} per GPU) 2ch 10s@44.1k Compose[EQ+RoomIR+PitchShift+-2] + "
                        "LUFS normalize + log-mel", "ms": ms, "clips_per_s": B * WORLD / ms * 1e3, "per_gpu_batch": B,
              "global_batch": B * WORLD,
              "timing": "CUDA events, barrier + synchronize both sides, max over ranks; no data-path collective"})
        del x, sig

    if "istft" in only:  # inverse STFT at cfg2's shape (64 x 2ch x 10 s @ 44.1 kHz, 2048/512)
        g = torch.Generator().manual_seed(0)
        x = (0.1 * torch.randn(64, 2, 441000, generator=g)).to(dev)
        sig = AudioSignal(x, 44100)
        sig.stft(window_length=2048, hop_length=512)
        X = sig.stft_data  # 905 MB > L2
        ms = timed(lambda: sig.istft(window_length=2048, hop_length=512), steps=10)
        w = torch.hann_window(2048, periodic=True, device=dev)
        Xr = X.reshape(128, 1025, -1)
        ms_torch = timed(lambda: torch.istft(Xr, 2048, 512, window=w, center=True, length=441000), steps=5)
        alg = X.numel() * 8 + x.numel() * 4
        emit({"config": "istft 64x2ch 10s@44.1k n_fft=2048 hop=512", "ms": ms, "ms_torch_istft_cufft": ms_torch,
                          "clips_per_s": 64 / ms * 1e3, "alg_bytes": alg, "achieved_GBps": alg / ms / 1e6,
                          "frac_of_hbm_peak": alg / ms / 1e6 / peak})

    if "dense" in only:  # arbitrary window length (csrc/dft.cu): the 25 ms / 10 ms / 80-mel speech front-end at 16 kHz
        B, T, sr = 64, 160000, 16000
        x = (0.1 * torch.randn(B, 1, T, generator=torch.Generator().manual_seed(0))).to(dev)
        sig = AudioSignal(x, sr)
        ms_stft = timed(lambda: sig.stft(window_length=400, hop_length=160), steps=10)
        ms_mel = timed(lambda: sig.mel_spectrogram(n_mels=80, window_length=400, hop_length=160, log=True), steps=10)
        sig.stft(window_length=400, hop_length=160)
        ms_inv = timed(lambda: sig.istft(window_length=400, hop_length=160), steps=10)
        w = torch.hann_window(400, periodic=True, device=dev)
        ms_torch = timed(lambda: torch.stft(x.reshape(B, T), 400, 160, window=w, center=True, return_complex=True), steps=10)
        nfr = 1 + T // 160
        macs = B * nfr * 400 * 201  # complex-real multiply-accumulates = FFMA2 instructions x 32 lanes
        emit({"config": "dense DFT 64 x 1ch x 10s@16k window 400 hop 160 (+ 80-mel log-mel, inverse)", "ms_stft": ms_stft,
              "ms_logmel": ms_mel, "ms_istft": ms_inv, "ms_torch_stft_cufft": ms_torch, "clips_per_s": B / ms_mel * 1e3,
              "gflops_stft": 4 * macs / ms_stft / 1e6, "fp32_peak_gflops": 2 * 148 * 128 * 1.965,
              "frac_of_fp32_peak": 4 * macs / ms_stft / 1e6 / (2 * 148 * 128 * 1.965)})

    if "gate" in only:  # SpectralGate (csrc/specmask.cu) at 64 x 2ch x 10 s: stft x2 + gate + istft
        from audiotools_b200.ml.layers import SpectralGate

        g = torch.Generator().manual_seed(0)
        x = (0.1 * torch.randn(64, 2, 441000, generator=g)).to(dev)
        nz = (0.01 * torch.randn(1, 1, 88200, generator=g)).to(dev)
        gate = SpectralGate().to(dev)
        sig, nzs = AudioSignal(x, 44100), AudioSignal(nz, 44100)
        ms = timed(lambda: gate(sig, nzs, 0.9), warmup=3, steps=5)
        from audiotools_b200.engine import get_engine

        s2 = sig.clone(); s2.stft(2048, 512, "sqrt_hann"); n2 = nzs.clone(); n2.stft(2048, 512, "sqrt_hann")
        ms_k = timed(lambda: get_engine().spec_gate(s2.stft_data, n2.stft_data, 3.0, torch.tensor([0.9]), gate._rf.tolist(),
                                                    gate._rt.tolist()), steps=10)
        alg = 2 * s2.stft_data.numel() * 8
        emit({"config": "SpectralGate 64x2ch 10s@44.1k (2048/512): clone + 2 stft + gate kernels + istft", "ms": ms,
              "ms_gate_kernels": ms_k, "alg_bytes_gate": alg, "gate_GBps": alg / ms_k / 1e6, "gate_frac_of_hbm_peak": alg / ms_k / 1e6 / peak})

    if "masked" in only:  # a prob = 0.5 augmentation chain, kernel-side bypass flags vs gather / scatter
        B, T, sr = 128, 441000, 44100
        g = torch.Generator().manual_seed(0)
        x = 0.1 * torch.randn(B, 1, T, generator=g)
        transform = tfm.Compose([tfm.VolumeNorm(prob=0.5), tfm.Equalizer(prob=0.5), tfm.LowPass(prob=0.5),
                                 tfm.HighPass(prob=0.5), tfm.PitchShift(("choice", [-2, 2]), prob=0.5)])
        sig = AudioSignal(x, sr)
        kwargs = transform.batch_instantiate(list(range(B)), sig)
        sig = sig.to(dev)
        from audiotools_b200 import util

        kwargs = util.prepare_batch(kwargs, dev)
        ms_aware = timed(lambda: transform(sig.clone(), **kwargs), warmup=3, steps=5)
        for t in transform.transforms:
            t._mask_aware = False
        ms_gather = timed(lambda: transform(sig.clone(), **kwargs), warmup=3, steps=5)
        emit({"config": f"masked chain batch={B} mono 10s@44.1k Compose[VolumeNorm, Equalizer, LowPass, HighPass, PitchShift] "
                        "each prob 0.5", "ms_bypass_flags": ms_aware, "ms_gather_scatter": ms_gather,
              "clips_per_s": B / ms_aware * 1e3})

    if "specaug" in only:  # SpectralTransform chain stft -> FrequencyMask -> TimeMask -> istft at cfg2's shape
        g = torch.Generator().manual_seed(0)
        B = 64
        x = (0.1 * torch.randn(B, 2, 441000, generator=g)).to(dev)
        fmin, fmax = (torch.rand(B, generator=g) * 8000).to(dev), None
        fmax = fmin + 2000.0
        tmin = (torch.rand(B, generator=g) * 9.0).to(dev)
        tmax = tmin + 0.25

        def ours():
            s = AudioSignal(x, 44100)
            s.stft(window_length=2048, hop_length=512)
            s.mask_frequencies(fmin, fmax)
            s.mask_timesteps(tmin, tmax)
            return s.istft(window_length=2048, hop_length=512)

        w = torch.hann_window(2048, periodic=True, device=dev)

        def stock_ops():  # the reference's own tensor ops (torch.stft / polar masks / torch.istft), run on the GPU
            X = torch.stft(x.reshape(-1, 441000), 2048, 512, window=w, center=True, return_complex=True)
            X = X.reshape(B, 2, 1025, -1)
            mag, ph = torch.abs(X), torch.angle(X)
            bins = torch.linspace(0, 22050, 1025, device=dev)[None, None, :, None].repeat(B, 1, 1, X.shape[-1])
            m = (fmin[:, None, None, None] <= bins) & (bins < fmax[:, None, None, None])
            X = mag.masked_fill(m, 0.0) * torch.exp(1j * ph.masked_fill(m, 0.0))
            mag, ph = torch.abs(X), torch.angle(X)
            bt = torch.linspace(0, 10.0, X.shape[-1], device=dev)[None, None, None, :].repeat(B, 1, 1025, 1)
            m = (tmin[:, None, None, None] <= bt) & (bt < tmax[:, None, None, None])
            X = mag.masked_fill(m, 0.0) * torch.exp(1j * ph.masked_fill(m, 0.0))
            return torch.istft(X.reshape(-1, 1025, X.shape[-1]), 2048, 512, window=w, center=True, length=441000)

        ms = timed(ours, warmup=2, steps=5)
        ms_stock = timed(stock_ops, warmup=1, steps=3)
        emit({"config": "specaug 64x2ch 10s@44.1k stft->mask_frequencies->mask_timesteps->istft (2048/512)",
                          "ms": ms, "ms_reference_tensor_ops_on_gpu": ms_stock, "clips_per_s": B / ms * 1e3})


if __name__ == "__main__":
    main()
    if WORLD > 1:
        import torch.distributed as dist

        dist.destroy_process_group()

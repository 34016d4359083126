#!/usr/bin/env python
"""bench.py -- the BASELINE.json metric: clips/sec for 10 s @ 44.1 kHz stereo clips through
LUFS-normalise (-24) + log-mel (n_fft 2048, hop 512, 128 mels)  [BASELINE.json configs[1]].

  python bench.py [--gpus N --steps K --warmup W]                    our arm (CUDA, one rank per GPU)
  python bench.py --impl reference [--gpus N --steps K --warmup W]    the reference's CPU path (oracle port)

One "step" = one pass of the hot path over one batch of 64 clips per GPU (weak scaling): the
loudness kernels, then the fused gain + STFT + mel + log kernel.  Outputs per step: normalised
waveform [B,2,441000], log-mel [B,2,128,862], LUFS [B].

  value     whole-job clips/s with inputs resident in HBM (device-timed, max over ranks)
  e2e       the same metric through the public AudioSignal API with HOST (pinned) inputs: the H2D copy
            of every step's batch and a D2H read of the step's LUFS vector are inside the timed region
  roofline  the dominant kernel (fused spectral) vs the measured HBM copy bandwidth
  cpu_baseline  the oracle (CPU port of the reference path) on this box's host cores, bounded sample
Timing hygiene: >= 3 warm-ups, CUDA events on the launching stream, inputs rotate over 3 distinct
226 MB batches (each > the 126 MB L2), nvidia-smi clocks sampled during the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

SR, T, C = 44100, 441000, 2
N_FFT, HOP, N_MELS, TARGET_DB = 2048, 512, 128, -24.0
N_FRAMES = 1 + T // HOP
BYTES_X = C * T * 4
BYTES_MEL = C * N_MELS * N_FRAMES * 4
WORKLOAD = "batch=64/GPU 2ch 10s@44.1kHz LUFS-normalize(-24)+log-mel(n_fft=2048,hop=512,n_mels=128)"


def make_batch(B, seed, device="cpu"):
    """SURVEY.md §8d synthetic input: 0.1*randn clipped, per-item gain U(0.05, 1)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    x = (0.1 * torch.randn(B, C, T, generator=g)).clamp(-1, 1)
    x = x * (0.05 + 0.95 * torch.rand(B, 1, 1, generator=g))
    return x.float().to(device)


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path
# ----------------------------------------------------------------------------------------------
def cpu_pipeline(x):
    from oracle import signal_path as sp

    y, lufs = sp.normalize(x, SR, TARGET_DB)
    mel = sp.mel_spectrogram(y, SR, N_MELS, window_length=N_FFT, hop_length=HOP, window_type="hann")
    return y, sp.log_mel(mel), lufs


def time_cpu(n_clips, reps, warmup):
    """Time the CPU port on ``n_clips`` clips.  torch's default (one thread per core) oversubscribes torch.stft /
    lfilter on a many-core host, so the thread count is calibrated first on a 4-clip sample and the fastest setting
    is used: the reference arm gets its best configuration, not an accidental slow one."""
    import torch

    cores = os.cpu_count() or 1
    x = make_batch(n_clips, 1234)
    cand = sorted({c for c in (cores, 64, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    best, best_t = cores, float("inf")
    if len(cand) > 1:
        xs = x[: min(4, n_clips)]
        torch.set_num_threads(cand[0])
        cpu_pipeline(xs)  # page in, build windows / filterbanks
        for c in cand:
            torch.set_num_threads(c)
            t0 = time.perf_counter()
            cpu_pipeline(xs)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    torch.set_num_threads(best)
    for _ in range(warmup):
        cpu_pipeline(x)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_pipeline(x)
        ts.append(time.perf_counter() - t0)
    return ts, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_clips = 16  # bounded sample of the 64-clip batch
    ts, cores = time_cpu(n_clips, reps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    total = sum(ts)
    value = n_clips * len(ts) / total
    line = {
        "impl": "reference", "metric": "clips/sec (10s@44.1kHz) log-mel+LUFS pipeline", "value": value,
        "unit": "clips/s", "n_gpus": args.gpus, "steps": len(ts), "warmup": max(1, min(args.warmup, 2)),
        "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample_clips_per_step": n_clips},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port",
                         "sample": f"{n_clips} of the 64 clips per step, {len(ts)} steps; torch threads calibrated over "
                                   f"{{all cores, 64, 32, 16, 8}} on 4 clips, fastest used"},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path = the in-repo oracle port (torch.stft + torchaudio.lfilter + restated "
                "pyloudnorm/librosa); the reference package itself cannot be installed here (see DESIGN.md)",
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as graft

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a B200: there is no CPU fallback. Use --impl reference "
                         "for the CPU path.")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()
    from audiotools_b200 import AudioSignal, _lib
    from audiotools_b200.engine import get_engine

    eng = get_engine()
    B = args.batch
    NBUF = 3
    xs = [make_batch(B, 100 + 7 * rank + i, dev) for i in range(NBUF)]
    db = torch.tensor([TARGET_DB], device=dev)
    win = AudioSignal.get_window("hann", N_FFT, dev)
    fb, lo, hi = AudioSignal._mel_tables(SR, N_FFT, N_MELS, 0.0, None, dev)
    from audiotools_b200.parallel import LoudnessGather

    # Whole-batch loudness statistics (the path's only exchange: 256 B per rank and step).  Preferred: one-sided
    # stores into every peer's buffer over NVLink (csrc/peer.cu) -- no rendezvous, no NCCL kernel next to the
    # persistent spectral kernel.  Fallback if the peer mapping cannot be set up: NCCL all-gather on a side stream.
    exchange, gather, exchange_kind = None, None, "none"
    if world > 1 and not os.environ.get("B2A_BENCH_NO_GATHER"):
        try:
            if os.environ.get("B2A_BENCH_NCCL_GATHER"):
                raise RuntimeError("NCCL all-gather requested")
            from audiotools_b200.parallel import PeerLoudnessExchange

            exchange = PeerLoudnessExchange(n_max=B)
            exchange_kind = "peer-store (cudaIpc + NVLink P2P stores, csrc/peer.cu)"
        except Exception as e:  # noqa: BLE001
            gather = LoudnessGather(side_stream=torch.cuda.Stream(device=dev))
            exchange_kind = f"nccl all_gather on a side stream ({type(e).__name__}: {e})"
    spec_events = []
    pending = []  # statistics are logging data: they are consumed one step late, never inside the step that made them

    def drain():
        while pending:
            kind, h = pending.pop(0)
            if kind == "peer":
                exchange.collect(h)
            else:
                gather.wait()

    def step(i, timed=False):
        x = xs[i % NBUF]
        if exchange is None:
            drain()  # previous step's statistics (long complete)
        lu = eng.lufs(x, SR, target_db=db)
        if exchange is not None:  # ONE launch: publish this step's vector, read the previous step's statistics
            seq, lu["loud_all_prev"] = exchange.put_collect(lu["loud"])
            pending[:] = [("peer", seq)]
        elif gather is not None:
            lu["loud_all"] = gather(lu["loud"])
            pending.append(("nccl", None))
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        out = eng.spectral(x, N_FFT, HOP, win, gain=lu["gain"], want_scaled=True, mel_fb=fb, mel_lo=lo, mel_hi=hi,
                           post=_lib.POST_LOG10, post_eps=1e-5, post_power=2.0, want_stft=False)
        if timed:
            e1.record()
            spec_events.append((e0, e1))
        return out, lu

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing
    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = eng.launches
    xl0 = exchange.launches if exchange is not None else 0
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        t0.record()
        for i in range(args.steps):
            step(args.warmup + i, timed=True)
        drain()  # the last step's statistics are inside the timed region too
        t1.record()
        barrier()
    ms = t0.elapsed_time(t1)
    launches = eng.launches - launches0 + (exchange.launches - xl0 if exchange is not None else 0)
    spec_ms = sum(a.elapsed_time(b) for a, b in spec_events) / max(1, len(spec_events))
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the public API, host (pinned) inputs, double-buffered H2D
    hx = [make_batch(B, 500 + i).pin_memory() for i in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    dbuf = [torch.empty(B, C, T, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    h_metric = torch.empty(B, pin_memory=True)

    def e2e_run(n):
        cur = torch.cuda.current_stream()
        for e in freed:
            e.record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[0])
            dbuf[0].copy_(hx[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n):
            b = i % 2
            if i + 1 < n:  # prefetch the next batch while this one computes
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[1 - b])
                    dbuf[1 - b].copy_(hx[(i + 1) % 2], non_blocking=True)
                    ready[1 - b].record(copy_stream)
            cur.wait_event(ready[b])
            sig = AudioSignal(dbuf[b], SR)
            sig.normalize(TARGET_DB)
            logmel = sig.mel_spectrogram(n_mels=N_MELS, window_length=N_FFT, hop_length=HOP, window_type="hann",
                                         log=True)
            y = sig.audio_data  # normalised waveform (came out of the same pass)
            assert y.data_ptr() != dbuf[b].data_ptr()
            h_metric.copy_(logmel.mean(dim=(1, 2, 3)), non_blocking=True)  # the step's result, D2H
            freed[b].record(cur)
        return logmel

    e2e_steps = max(3, min(args.steps, 20))
    e2e_run(3)
    barrier()
    w0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - w0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(te.item())
    if exchange is not None:
        exchange.close()  # collective (barrier inside): all ranks, before the non-zero ranks leave

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = B * (2 * BYTES_X + BYTES_MEL)  # read x + write y + write log-mel, each once
    achieved = alg_bytes / (spec_ms * 1e-3) / 1e9
    lufs_ms = ms / args.steps - spec_ms
    roof = {"kernel": "spectral_warp_kernel<10> (gain + STFT + |.| + mel + log10, fused)", "bound": "hbm",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at B=64 from the ncu --set full capture
            # profiles/r01h_prof_spectral_v7_ncu_full_summary.csv (226.0 + 229.2 MB), scaled to this batch
            "traffic": (226.3e6 + 229.2e6) * B / 64,  # dram__bytes_read + write of one launch: profiles/r01r_prof_spectral_final_ncu_full_summary.csv
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": spec_ms,
            "rest_of_step_ms": lufs_ms,
            "rest_of_step": "lufs kernels (read x once: %.0f GB/s algorithmic)" % (B * BYTES_X / max(lufs_ms, 1e-9) / 1e6)}

    # ---- CPU baseline (bounded sample, rank 0 at any N; cheap)
    cpu = None
    if not args.no_cpu:
        n_clips = 8
        ts, cores = time_cpu(n_clips, reps=2, warmup=1)
        cpu = {"value": n_clips * len(ts) / sum(ts), "unit": "clips/s", "cores": cores, "kind": "port",
               "sample": f"{n_clips} clips per rep, {len(ts)} reps after 1 warm-up; torch threads calibrated over "
                         f"{{all cores, 64, 32, 16, 8}}, fastest used"}

    line = {
        "metric": "clips/sec (10s@44.1kHz) log-mel+LUFS pipeline", "value": value, "unit": "clips/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "per_gpu_batch": B,
                   "parallelism": f"batch-sharded x{world}, no data-path collective"
                                  + (f" (+ per-item LUFS exchange: {exchange_kind})" if world > 1 else ""),
                   "l2": f"inputs rotate over {NBUF} distinct {B * BYTES_X / 1e6:.0f} MB batches (> 126 MB L2)"},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * BYTES_X,
                "d2h_bytes_per_step": B * 4, "steps": e2e_steps,
                "api": "AudioSignal(x).normalize(-24).mel_spectrogram(..., log=True)"},
        "gpu_launches": launches, "clocks": clocks.summary(),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

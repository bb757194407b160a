#!/usr/bin/env python
"""bench.py -- audio-seconds per second of the denoise + LUFS + log-mel front-end (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1|cfg2|cfg2s|cfg3|cfg4|cfg5]

Workloads: cfg2 (default, the headline: BASELINE.json configs[1]); cfg1 / cfg3 / cfg4 / cfg5 are the other BASELINE configs
(cfg4 = one long recording time-sharded over the ranks with NCCL collectives, cfg5 = latency of one 30 s chunk of 48 kHz
stereo through the front door, replayed as a CUDA graph); cfg2s = the stationary gate on cfg2's recording.

One "step" = one pass of the whole hot path (noise detect / threshold, STFT, mask, iSTFT, peak normalise, loudness,
log-mel) over one batch of synthetic recordings resident in HBM.  N = 1 runs BASELINE.json configs[1]: one 1 h
16 kHz mono "meeting" recording, non-stationary denoise, 128-mel log-mel (30 s Whisper padding).  For N > 1 every rank
processes its own recording (file sharding, no data-path collective): weak scaling.
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the reference's libraries are not
installable offline) on a bounded sample of the same workload, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

# Rank 0 prints exactly ONE line on stdout, the JSON.  Native libraries write there too (NCCL prints its version banner with
# printf on communicator creation), so file descriptor 1 is pointed at stderr and the JSON goes out through a private
# duplicate of the real stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w") if __name__ == "__main__" else sys.stdout
if __name__ == "__main__":
    sys.stdout.flush()
    os.dup2(2, 1)


def emit(line: dict) -> None:
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 16000
WORKLOADS = {
    # name: (description, seconds per recording, recordings per GPU, stationary, n_mels, mel padding, meeting noise)
    "cfg1": ("cfg1: 60 s 16 kHz mono, stationary denoise + LUFS + 80-mel (30 s pad)", 60.0, 1, True, 80, 480000, False),
    "cfg2": ("cfg2: 1 h 16 kHz mono meeting, non-stationary denoise + LUFS + 128-mel (30 s pad)", 3600.0, 1, False, 128, 480000, True),
    "cfg3": ("cfg3: 1024 x 30 s clips, stationary denoise + LUFS + 80-mel, file-sharded", 30.0, 1024, True, 80, 0, False),
    # not a BASELINE config: the reference's own Protocol method (stationary gate) on the 1 h recording of config 2
    "cfg2s": ("cfg2s: 1 h 16 kHz mono meeting, STATIONARY denoise (auto-detected noise clip) + LUFS + 80-mel (30 s pad)", 3600.0, 1, True, 80, 480000, True),
}
CFG5 = "cfg5: 30 s chunk of 48 kHz stereo (speech + music bed + heavy noise) -> front door (peak normalise, 48k->16k, downmix) -> stationary denoise + LUFS + 80-mel; latency per chunk"



CPU_SAMPLE_S = 600.0     # 16 whole 600 000-sample denoise chunks: the reference's chunk loop does no wasted work on a partial chunk


def config_of(workload: str, world: int) -> dict:
    """The `config` object of the JSON line -- identical for `--impl ours` and `--impl reference`."""
    desc, seconds, per_gpu, stationary, n_mels, padding, meeting = WORKLOADS[workload]
    return {"workload": desc, "recordings_per_gpu": per_gpu, "seconds_per_recording": seconds,
            "l2": ("working set (STFT 4 KB/frame) exceeds the 126 MB L2; no explicit flush" if seconds * per_gpu >= 600 else
                   "working set fits the 126 MB L2: a latency figure, not a roofline one"),
            "parallelism": f"file-sharded x{world}"}


def cpu_sample_of(workload: str, sample_seconds: float):
    """(seconds, mel padding) of the recordings the CPU arm processes: the workload's own recording when it is short, else a
    600 s excerpt of it with the Whisper padding scaled by the same ratio (30 s on 1 h -> 5 s on 600 s), so that the share of
    log-mel work on padding zeros and the chunk geometry of the denoise (whole chunks, multi-chunk path) are the workload's."""
    desc, seconds, per_gpu, stationary, n_mels, padding, meeting = WORKLOADS[workload]
    if seconds <= sample_seconds:
        return seconds, padding
    chunk_s = 600000 / SR
    sample = max(chunk_s, round(sample_seconds / chunk_s) * chunk_s)
    pad = int(round(padding * sample / seconds / 160.0)) * 160
    return sample, pad


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (via NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


STAGE_KERNELS = {"stft": ["stft1024_kernel"], "ns_mask": ["ns_scan_kernel", "ns_mask_smooth_kernel"], "smooth_mask": ["smooth_mask_kernel"],
                 "istft": ["istft_kernel"], "kweight": ["kweight_energy_kernel"], "gate": ["gate_kernel"], "logmel": ["logmel_kernel"],
                 "mel_finalize": ["mel_finalize_kernel"], "noise": ["frame_stats_kernel", "select_clip_kernel", "noise_pass_kernel", "noise_finalize_kernel"]}


def ncu_traffic(workload, stage):
    """DRAM bytes per launch of the stage's kernels from the committed `ncu --set full` capture of this workload
    (profiles/traffic_<workload>.json, written by profiles/summarize.py); None when there is no capture."""
    path = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f)
    vals = [t["bytes_per_launch"].get(k) for k in STAGE_KERNELS.get(stage, [])]
    vals = [v for v in vals if v is not None]
    return (float(sum(vals)) if vals else None), t.get("tag")


def ncu_step_totals(workload):
    """(DRAM bytes of one step, tag, {kernel: warp instructions per launch}) from the committed ncu capture of this workload."""
    path = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
    if not os.path.exists(path):
        return None, None, None
    with open(path) as f:
        t = json.load(f)
    per_step = t.get("launches_per_step", {})
    total = sum(v * per_step.get(k, 1) for k, v in t["bytes_per_launch"].items())
    return float(total), t.get("tag"), t.get("warp_inst_per_launch")


def stage_bytes(plan, lengths, stationary, n_mels, padding, touched_tiles=None):
    """Algorithmic bytes per launch of every stage = each kernel's mandatory inputs + outputs."""
    rows = 0
    need_rows = 0
    pad_samples = 0
    tiles16 = 0
    tiles32 = 0
    for n in lengths:
        if n > 600000:
            chunks = [(min(600000, n - s), 660000) for s in range(0, n, 600000)]
        else:
            chunks = [(n, n + 60000)]
        for keep, L in chunks:
            rows += L // 256 + 1
            tiles16 += (L // 256 + 1 + 31) // 32     # STFT_TILE frames per scan partial
            tiles32 += (L // 256 + 1 + 3 + 31) // 32
            need_rows += (30000 + keep - 1) // 256 - 30000 // 256 + 1 + 3
            pad_samples += L
    n_tot = sum(lengths)
    mel_frames = sum((n + padding) // 160 for n in lengths)
    b = {}
    b["noise"] = 2 * n_tot if stationary else 0
    # non-stationary: the STFT kernel also writes the scan partials P, Q (one float each per 16-frame tile and bin); the
    # mask stage = scan (reads P, Q, writes the tile-end states) + mask kernel (reads S once, the states, writes M)
    b["stft"] = 2 * pad_samples + rows * 4096 + (rows * 68 if stationary else tiles16 * 513 * 8)
    b["ns_mask"] = 0 if stationary else rows * (4096 + 513 * 4) + tiles16 * 513 * 16 + tiles32 * 513 * 8
    b["smooth_mask"] = (rows * 68 + rows * 513 * 4) if stationary else 0
    b["istft"] = need_rows * (4096 + 513 * 4) + 2 * n_tot
    b["kweight"] = 2 * n_tot
    b["gate"] = 0
    b["logmel"] = 2 * n_tot + 2 * n_tot + 4 * n_mels * mel_frames
    # mel_finalize rewrites only the 32-frame tiles whose minimum lies below max - 8 (`touched`, counted on the device after a
    # step); every CTA reads its tile minimum (4 B)
    tiles = sum(((n + padding) // 160 + 31) // 32 for n in lengths)
    b["mel_finalize"] = (8 * n_mels * 32 * touched_tiles if touched_tiles is not None else 8 * n_mels * mel_frames) + 4 * tiles
    return b


def _cpu_worker(conn, wl, sample_seconds, seed):
    """One host thread of the CPU baseline: owns one recording, runs the oracle front-end on it on request."""
    import torch as _torch
    _torch.set_num_threads(1)
    from modular_audio_pipeline_b200.synth import speech_noise
    from oracle import reference_path as O
    desc, seconds, per_gpu, stationary, n_mels, padding, meeting = WORKLOADS[wl]
    sample, pad = cpu_sample_of(wl, sample_seconds)
    pcm = speech_noise(sample, seed, meeting=meeting).numpy()
    conn.send(len(pcm) / SR)
    while conn.recv() == "go":
        t0 = time.perf_counter()
        O.front_end(pcm, SR, stationary, n_mels, pad)
        conn.send(time.perf_counter() - t0)


class CpuPool:
    """The CPU arm: the oracle (restated reference path; `kind: "port"`) on ALL host threads the process may use -- one
    single-threaded worker process per core, each with its own recording of the workload's kind; a step = every worker
    processes its recording once, concurrently."""

    def __init__(self, wl, sample_seconds, n_workers=None):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")                      # the parent may hold a CUDA context: never fork it
        self.n = n_workers or max(1, len(os.sched_getaffinity(0)))
        self.conns, self.procs = [], []
        for i in range(self.n):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(b, wl, sample_seconds, 1000 + i), daemon=True)
            p.start()
            self.conns.append(a); self.procs.append(p)
        self.audio_s = sum(c.recv() for c in self.conns)   # workers report the seconds they hold once they are ready

    def step(self):
        t0 = time.perf_counter()
        for c in self.conns:
            c.send("go")
        for c in self.conns:
            c.recv()
        return time.perf_counter() - t0

    def close(self):
        for c in self.conns:
            try:
                c.send("stop")
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=10)


def cpu_sample_text(pool, workload, sample_seconds):
    sample, pad = cpu_sample_of(workload, sample_seconds)
    seconds = WORKLOADS[workload][1]
    what = "whole recordings of the workload" if sample >= seconds else (
        f"{sample:.0f} s excerpts of the workload's recording = {sample * SR / 600000:.0f} whole 600 000-sample denoise chunks, Whisper padding "
        f"scaled by the same ratio ({pad} samples)")
    return (f"{pool.n} x {sample:.0f} s per step through oracle/reference_path.front_end (restated reference path), one single-threaded "
            f"process per host thread; {what}")


def run_reference(args):
    """The CPU arm: the reference's path (oracle port -- its libraries cannot be installed offline) on every host thread, the same
    `config` as our arm, each step a bounded sample of the workload (cpu_sample_of)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload in ("cfg4", "cfg5"):
        args.workload = "cfg3"          # the CPU arm has no sharded / 48 kHz front door leg: 30 s stationary clips as in config 3
    pool = CpuPool(args.workload, args.cpu_sample)
    times = []
    for i in range(args.warmup + args.steps):
        dt = pool.step()
        if i >= args.warmup:
            times.append(dt)
    pool.close()
    ms = 1e3 * float(np.mean(times))
    value = pool.audio_s / (ms / 1e3)
    text = cpu_sample_text(pool, args.workload, args.cpu_sample)
    line = {"impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload, args.gpus),
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": pool.n, "kind": "port", "sample": text},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_cfg5(args):
    """BASELINE config 5: latency of ONE 30 s chunk, host buffer in -> host buffers out, synchronous (p50 / p99)."""
    from modular_audio_pipeline_b200 import FrontEnd
    from modular_audio_pipeline_b200.synth import music_bed_stereo
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    fe = FrontEnd(dev)
    h_in = music_bed_stereo(30.0, 5).pin_memory()
    n16 = 30 * SR
    plan = fe.plan([n16], stationary=True, n_mels=80, mel_padding=0)
    d_in = torch.empty_like(h_in, device=dev)
    out_pcm = torch.empty(max(plan.total_out, 8), dtype=torch.int16, device=dev)
    mel = torch.empty(max(plan.total_mel, 4), dtype=torch.float32, device=dev)
    h_out = torch.empty(plan.total_out, dtype=torch.int16).pin_memory()
    h_lufs = torch.empty(1, dtype=torch.float64).pin_memory()
    _ = plan.workspace

    def issue():                                              # everything asynchronous on the current stream
        d_in.copy_(h_in, non_blocking=True)
        mono = fe.front_door(d_in, 2, 48000, SR, normalize=False)     # the plan's own peak normalise follows the denoise (pipeline.py:346)
        res = plan.process(mono[:n16].contiguous() if mono.numel() >= n16 else torch.nn.functional.pad(mono, (0, n16 - mono.numel())),
                           out_pcm=out_pcm, mel=mel)
        h_out.copy_(out_pcm[:plan.total_out], non_blocking=True)
        h_lufs.copy_(res.lufs, non_blocking=True)

    def measure(run):
        for _ in range(max(args.warmup, 5)):
            run()
        runs = max(args.steps, 200)
        lat = []
        for _ in range(runs):
            t0 = time.perf_counter()
            run()
            lat.append((time.perf_counter() - t0) * 1e3)
        return np.sort(np.array(lat)), runs

    def eager():
        issue()
        torch.cuda.synchronize()

    lat_eager, runs = measure(eager)
    launches_per_chunk = plan.launches
    eager()
    launches_per_chunk = plan.launches - launches_per_chunk
    from modular_audio_pipeline_b200 import capture_graph
    graph, _ = capture_graph(issue, dev)                      # the whole chunk as ONE graph launch

    def replay():
        graph.replay()
        torch.cuda.synchronize()

    lat, runs = measure(replay)
    p50, p99 = float(lat[len(lat) // 2]), float(lat[min(len(lat) - 1, int(0.99 * len(lat)))])
    line = {"metric": "audio_seconds_per_second", "value": 30.0 / (p50 / 1e3), "unit": "audio-s/s", "n_gpus": 1, "steps": runs,
            "warmup": max(args.warmup, 5), "ms_per_step": p50, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": CFG5, "l2": "one chunk is L2-resident: latency, not roofline"},
            "latency_ms": {"p50": p50, "p99": p99, "min": float(lat[0]), "runs": runs, "mode": "CUDA graph replay of the whole chunk",
                           "eager_p50": float(lat_eager[len(lat_eager) // 2]), "eager_p99": float(lat_eager[min(len(lat_eager) - 1, int(0.99 * len(lat_eager)))]),
                           "what": "pinned host 48 kHz stereo chunk in -> H2D -> front door -> denoise -> peak normalise -> loudness -> log-mel -> D2H of PCM + LUFS, synchronous"},
            "h2d_bytes_per_step": h_in.numel() * 2, "d2h_bytes_per_step": plan.total_out * 2 + 8, "gpu_launches": launches_per_chunk * runs}
    emit(line)


def measure_time_sharded(args, dev, world, rank, steps, warmup):
    """BASELINE config 4 on the ranks of this job: ONE long recording (1 h per rank: 8 h on 8 GPUs) time-sharded with 30 000-sample
    denoise halos and the five fused exchanges of sharded.py (stationary denoise, LUFS, 80-mel, 30 s pad).  The process group
    must exist.  Returns the `time_sharded` object: eager step, the step replayed as ONE CUDA graph (kernels + NCCL), the step
    without its collectives, and the same 1 h per GPU through the unsharded fused entry (the weak-scaling reference)."""
    import torch.distributed as dist
    from modular_audio_pipeline_b200 import FrontEnd
    from modular_audio_pipeline_b200.sharded import Comm, CudaShardBackend, NullComm, ShardGeometry, run_sharded
    from modular_audio_pipeline_b200.synth import speech_noise
    hour = 3600 * SR
    geo = ShardGeometry(hour * world, rank, world) if world > 1 else None
    fe = FrontEnd(dev)

    def timed(fn, n):
        torch.cuda.synchronize()
        dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1) / n], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # the weak-scaling reference: this rank's hour through the unsharded fused entry, no collective
    own = speech_noise(3600.0, 2 + rank, device=dev)
    plan1 = fe.plan([hour], stationary=True, n_mels=80, mel_padding=480000)
    pcm1 = plan1.pack([own])
    out1 = torch.empty(max(plan1.total_out, 8), dtype=torch.int16, device=dev)
    mel1 = torch.empty(max(plan1.total_mel, 4), dtype=torch.float32, device=dev)
    for _ in range(warmup):
        plan1.process(pcm1, out_pcm=out1, mel=mel1)
    ms_n1 = timed(lambda: plan1.process(pcm1, out_pcm=out1, mel=mel1), steps)
    plan1.close()
    del pcm1, out1, mel1
    torch.cuda.empty_cache()
    if world == 1:
        geo = None
        return {"ms_per_step_unsharded_1h": ms_n1, "note": "one rank: nothing to shard"}
    sl = geo.local_slice()
    # every rank synthesises the samples it holds (own range + halos); the seams need not be continuous for timing
    local_pcm = speech_noise((sl.stop - sl.start) / SR, 2 + rank, device=dev)[: sl.stop - sl.start]
    backend = CudaShardBackend(fe, geo, local_pcm, stationary=True, n_mels=80, mel_padding=480000)
    comm = Comm(None)
    for _ in range(warmup):
        run_sharded(backend, geo, stationary=True, comm=comm)
    launches0, calls0 = backend.plan.launches, comm.calls
    ms_eager = timed(lambda: run_sharded(backend, geo, stationary=True, comm=comm), steps)
    launches = (backend.plan.launches - launches0) / steps
    collectives = (comm.calls - calls0) / steps
    null = NullComm(None)
    run_sharded(backend, geo, stationary=True, comm=null)
    ms_nocomm = timed(lambda: run_sharded(backend, geo, stationary=True, comm=null), steps)
    graph_ms, graph_err = None, None
    try:
        replay = backend.capture(Comm(None))
        for _ in range(2):
            replay()
        graph_ms = timed(replay, steps)
    except Exception as e:                                    # the eager figure stands
        graph_err = f"{type(e).__name__}: {e}"[:300]
    backend.release_graph()                                   # before destroy_process_group: a live graph pins the NCCL communicator
    out, mel, lufs = run_sharded(backend, geo, stationary=True, comm=comm)
    torch.cuda.synchronize()
    best = min(ms_eager, graph_ms) if graph_ms else ms_eager
    return {"workload": f"cfg4: one {world} h recording time-sharded over {world} GPUs (whole 600 000-sample chunks, 30 000-sample halos), "
                        "stationary denoise + LUFS + 80-mel (30 s pad)",
            "ms_per_step": best, "value": 3600.0 * world / (best / 1e3), "unit": "audio-s/s",
            "ms_per_step_eager": ms_eager, "ms_per_step_graph": graph_ms, "graph_error": graph_err,
            "ms_per_step_unsharded_1h": ms_n1, "efficiency_vs_n1": ms_n1 / best,
            "collectives_per_step": collectives, "collectives_ms": ms_eager - ms_nocomm, "ms_per_step_without_collectives": ms_nocomm,
            "kernel_launches_per_step": launches, "lufs": float(lufs.cpu().item()),
            "how": "CUDA events around `steps` back-to-back steps, max over ranks; efficiency = (1 h per GPU, unsharded fused entry, same "
                   "run) / (sharded step); collectives_ms = eager step - the same step with the five collectives skipped"}


def run_cfg4(args):
    """`--workload cfg4`: the time-sharded mode as the headline line of this run."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    sampler = ClockSampler(local)
    sampler.start()
    ts = measure_time_sharded(args, dev, world, rank, args.steps, max(args.warmup, 3))
    sampler.stop_flag.set()
    if rank == 0:
        ms = ts.get("ms_per_step", ts["ms_per_step_unsharded_1h"])
        line = {"metric": "audio_seconds_per_second", "value": 3600.0 * world / (ms / 1e3), "unit": "audio-s/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": ts.get("workload", "cfg4 on one GPU = the unsharded fused entry"), "seconds_per_rank": 3600.0,
                           "l2": "working set exceeds the 126 MB L2", "parallelism": f"time-sharded x{world}"},
                "time_sharded": ts, "gpu_launches": int(ts.get("kernel_launches_per_step", 0) * args.steps), "clocks": sampler.summary()}
        emit(line)
    dist.destroy_process_group()


def bench_extras(fe, dev, args):
    """SURVEY.md 8(f) rows next to the headline: remove_silence on the device (row 1) and the file-batch entry with its
    prefetching reader / background writer from a RAM disk (row 4).  Wall-clock figures: both include their host work."""
    import shutil
    import tempfile
    from modular_audio_pipeline_b200 import GpuAudioPreprocessor, wavio
    from modular_audio_pipeline_b200._compat import TimestampMapping
    from modular_audio_pipeline_b200.silence import remove_silence
    from modular_audio_pipeline_b200.synth import speech_noise
    out = {}
    # ---- silence removal of a 1 h recording (pydub: O(N x 250) audioop.rms calls + a Python loop per range) ----
    x = speech_noise(3600.0, 1, sigma=0.0002, device=dev)
    remove_silence(fe, x, SR, 250, 40.0, 100, True, TimestampMapping)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cleaned, maps = remove_silence(fe, x, SR, 250, 40.0, 100, True, TimestampMapping)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    from oracle import reference_path as O
    xs = x[: 30 * SR].cpu().numpy()
    t0 = time.perf_counter()
    O.remove_silence(xs, SR)
    dc = time.perf_counter() - t0
    out["silence"] = {"ms_per_hour": dt * 1e3, "value": 3600.0 / dt, "unit": "audio-s/s", "chunks": len(maps), "frames_out": int(cleaned.numel()),
                      "cpu_value": 30.0 / dc, "cpu_sample": "30 s excerpt through oracle/reference_path.remove_silence (pydub's per-millisecond "
                                                            "audioop.rms loop restated on the stdlib audioop), one host thread",
                      "what": "remove_silence(min_silence_len=250, offset 40 dB, margin 100 ms, preserve_timestamps) on a device-resident 1 h "
                              "recording: window RMS from prefix sums, range merge, margins, cross-fade and compaction on the device; wall clock "
                              "incl. the two scalar-sized device->host reads and building the TimestampMapping list"}
    del x, cleaned
    # ---- file batch from a RAM disk: 256 x 30 s WAV files (config 3's clips) ----
    d = tempfile.mkdtemp(prefix="maf_files_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        n_files = 256
        base = [speech_noise(30.0, 100 + i, device=dev).cpu().contiguous() for i in range(16)]
        paths = []
        for i in range(n_files):
            p = os.path.join(d, f"clip{i:04d}.wav")
            wavio.write_pcm16(p, base[i % 16], SR)
            paths.append(p)
        pre = GpuAudioPreprocessor(SR, os.path.join(d, "out"), device=dev)
        pre._fe = fe
        res = {}
        for write in (True, False):
            list(pre.process_files(paths[:64], batch_size=64, n_mels=80, mel_padding=0, write_output=write))     # warm-up: plan, pinned buffers
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            k = sum(1 for _ in pre.process_files(paths, batch_size=64, n_mels=80, mel_padding=0, write_output=write))
            torch.cuda.synchronize()
            dtf = time.perf_counter() - t0
            res["with_output_files" if write else "mel_only"] = {"s": dtf, "files_per_s": k / dtf, "value": 30.0 * k / dtf, "unit": "audio-s/s"}
        # the two host-side ceilings of this entry on this box: reading the batch into page-locked memory, writing it back
        t0 = time.perf_counter()
        for _ in wavio.BatchReader(paths, 64):
            pass
        rd = time.perf_counter() - t0
        wr_pool = wavio.WavWriter(4)
        t0 = time.perf_counter()
        jobs = [wr_pool.submit(os.path.join(d, "out", f"w{i}.wav"), base[i % 16], SR, None) for i in range(n_files)]
        for j in jobs:
            j.wait()
        wrt = time.perf_counter() - t0
        wr_pool.close()
        res["read_only_ceiling"] = {"s": rd, "value": 30.0 * n_files / rd, "GBps": n_files * 960044 / rd / 1e9}
        res["write_only_ceiling"] = {"s": wrt, "value": 30.0 * n_files / wrt, "GBps": n_files * 960044 / wrt / 1e9}
        res["what"] = (f"GpuAudioPreprocessor.process_files on {n_files} x 30 s 16 kHz WAV files on a RAM disk, batches of 64: wavio.BatchReader "
                       "reads batch k+1 into page-locked memory while batch k runs, the background writer stores batch k-1's "
                       "`_denoised_norm_loudnorm.wav` files; the log-mel stays on the device for the transcriber")
        out["file_batch"] = res
        pre.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg4", "cfg5"])
    ap.add_argument("--cpu-sample", type=float, default=CPU_SAMPLE_S, help="seconds of audio EACH host thread of the CPU baseline processes per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-time-sharded", action="store_true", help="N > 1: skip the time-sharded (config 4) leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the silence-removal and file-batch legs")
    ap.add_argument("--e2e-depth", type=int, default=3, help="plans / streams in flight in the HostPipeline e2e leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "cfg5":
        run_cfg5(args)
        return
    if args.workload == "cfg4":
        run_cfg4(args)
        return

    import torch.distributed as dist
    from modular_audio_pipeline_b200 import FrontEnd
    from modular_audio_pipeline_b200.synth import speech_noise

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import faulthandler
        faulthandler.dump_traceback_later(900, exit=True)      # a stuck collective must not hold N GPUs until the driver's limit
        dist.init_process_group("nccl", device_id=dev)
    desc, seconds, per_gpu, stationary, n_mels, padding, meeting = WORKLOADS[args.workload]
    n = int(round(seconds * SR))
    lengths = [n] * per_gpu

    # synthetic input, generated on the device, one distinct recording set per rank
    if per_gpu == 1:
        clips = [speech_noise(seconds, 1 + rank, meeting=meeting, device=dev)]
    else:
        base = [speech_noise(seconds, 100 + rank * 16 + i, device=dev) for i in range(16)]   # 16 distinct clips, tiled
        clips = [base[i % 16] for i in range(per_gpu)]
    fe = FrontEnd(dev)
    plan = fe.plan(lengths, stationary=stationary, n_mels=n_mels, mel_padding=padding)
    pcm = plan.pack(clips)
    out_pcm = torch.empty(max(plan.total_out, 8), dtype=torch.int16, device=dev)
    mel = torch.empty(max(plan.total_mel, 4), dtype=torch.float32, device=dev)
    _ = plan.workspace
    torch.cuda.synchronize()

    def step():
        return plan.process(pcm, out_pcm=out_pcm, mel=mel)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    plan.profile(min(args.steps, 64))
    launches0 = plan.launches
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    sampler.stop_flag.set()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = plan.launches - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    audio_s = seconds * per_gpu * world
    value = audio_s / (ms_step / 1e3)

    # per-stage device times (CUDA events recorded by the library on the launching stream, inside the timed region)
    nprof = min(args.steps, 64)
    stage = {}
    for s in range(nprof):
        for k, v in plan.stage_ms(s).items():
            stage[k] = stage.get(k, 0.0) + v / nprof
    touched = None
    if n_mels:
        # tiles mel_finalize rewrites: minimum below (recording maximum - 8); both live in the workspace after a step
        tmin = plan.buffer("mel_tmin", torch.int32).view(len(lengths), -1)
        mmax = plan.buffer("mel_max", torch.int32)

        def unordered(t):                                   # maf_common.cuh float_to_ordered, inverted
            return torch.where(t >= 0, t, t ^ 0x7fffffff).view(torch.float32)
        valid = tmin != 0x7fffffff
        touched = int(((unordered(tmin) < (unordered(mmax)[:, None] - 8.0)) & valid).sum().item())
    sbytes = stage_bytes(plan, lengths, stationary, n_mels, padding, touched)
    hbm_peak, peak_kind = peaks()
    dom = max(stage, key=lambda k: stage[k])
    achieved = sbytes[dom] / (stage[dom] * 1e-3) / 1e9 if stage[dom] > 0 else 0.0
    traffic, traffic_tag = ncu_traffic(args.workload, dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": f"profiles/{traffic_tag}_ncu.md" if traffic_tag else None, "peak_kind": peak_kind, "algorithmic_bytes_per_launch": sbytes[dom],
                "kernel_ms": stage[dom],
                "stages": {k: {"ms": round(stage[k], 4), "GBps": round(sbytes[k] / (stage[k] * 1e-3) / 1e9, 1) if stage[k] > 0 else 0.0}
                           for k in stage}}
    # the three denoise kernels run within ~1 % of each other: which one is "dominant" can change from run to run, so every
    # stage within 2 % of the longest is listed with its own fraction
    roofline["near_dominant"] = [{"kernel": k, "ms": round(stage[k], 4), "frac": round(sbytes[k] / (stage[k] * 1e-3) / 1e9 / hbm_peak, 3)}
                                 for k in sorted(stage, key=lambda k: -stage[k]) if stage[k] >= 0.98 * stage[dom]]
    total_bytes = sum(sbytes.values())
    roofline["pipeline_GBps"] = total_bytes / (ms_step * 1e-3) / 1e9
    roofline["pipeline_frac"] = roofline["pipeline_GBps"] / hbm_peak
    # the same fraction from MEASURED DRAM traffic (ncu dram__bytes_read + write of every kernel of a step, committed capture)
    ncu_total, ncu_tag, warp_inst = ncu_step_totals(args.workload)
    if ncu_total:
        roofline["pipeline_dram_GBps"] = ncu_total / (ms_step * 1e-3) / 1e9
        roofline["pipeline_dram_frac"] = roofline["pipeline_dram_GBps"] / hbm_peak
        roofline["pipeline_dram_source"] = f"profiles/{ncu_tag}_ncu.md"
    # The kernels of this path are FP32-issue-bound, not DRAM-bound: the ceiling they run against is the
    # warp-instruction issue rate, 4 schedulers x 1 instruction / clock per SM.  Warp instructions per launch come from the
    # committed ncu capture (smsp__inst_executed.sum), the clock from the NVML samples of this run.
    clk = sampler.summary().get("sm_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    issue_peak = 4.0 * n_sm * clk * 1e6
    if warp_inst:
        kern = STAGE_KERNELS.get(dom, [])
        wi = sum(warp_inst.get(k, 0.0) for k in kern)
        if wi and stage[dom] > 0:
            roofline["issue"] = {"kernel": dom, "warp_inst_per_launch": wi, "achieved_Ginst_s": wi / (stage[dom] * 1e-3) / 1e9,
                                 "peak_Ginst_s": issue_peak / 1e9, "frac": wi / (stage[dom] * 1e-3) / issue_peak,
                                 "source": f"profiles/{ncu_tag}_ncu.md"}
        tot = sum(warp_inst.values())
        roofline["pipeline_issue_frac"] = tot / (ms_step * 1e-3) / issue_peak

    # end-to-end through the host-buffer entry: every step copies its PCM from pinned host memory to the device and
    # brings the final PCM, the log-mel and the LUFS back; HostPipeline overlaps step i's D2H with step i+1's H2D +
    # kernels (two plans / streams).  A second figure times the synchronous C-ABI call maf_process_host.
    # End to end through the host-buffer entry (HostPipeline): every step copies its PCM from pinned host memory to the
    # device and brings the results back, `depth` plans / streams overlap step i's copies with step i+1's kernels.
    #   primary   : what the drop-in hands back to the HOST is the final PCM (the `_loudnorm.wav` content) and the LUFS; the
    #               log-mel is the transcriber input and stays on the device, exactly where openai-whisper computes it
    #               (whisper.log_mel_spectrogram runs on the model's device) -- transcriber.py's adapters take it from there
    #   all_to_host: the log-mel is copied to the host as well (301 MB instead of 115 MB per hour of audio: PCIe-bound)
    # A third figure times the synchronous C-ABI call maf_process_host.
    e2e = None
    if not args.no_e2e:
        from modular_audio_pipeline_b200 import HostPipeline
        host_pcm = pcm[:plan.total_pcm].cpu()
        del mel, out_pcm
        plan._ws = None
        torch.cuda.empty_cache()
        ksteps = max(4, min(args.steps, 20))

        def run_pipe(depth, mel_to_host):
            pipe = HostPipeline(fe, lengths, depth=depth, mel_to_host=mel_to_host, encoder_windows=not mel_to_host, stationary=stationary,
                                n_mels=n_mels, mel_padding=padding)
            for h in pipe.h_in:
                h[:plan.total_pcm].copy_(host_pcm)
            for _ in range(3):
                pipe.submit()
            pipe.drain()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                pipe.submit()
            pipe.drain()
            dt = (time.perf_counter() - t0) / ksteps
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return pipe, te.item()

        pipe, dt = run_pipe(args.e2e_depth, False)
        e2e = {"value": audio_s / dt, "unit": "audio-s/s", "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
               "ms_per_step": dt * 1e3, "steps": ksteps,
               "api": "HostPipeline.submit: pinned host PCM in (H2D), kernels, final PCM + LUFS back to pinned host memory (D2H) every "
                      f"step, depth-{args.e2e_depth} overlap; the log-mel is handed to the transcriber on the device -- as the float16 30 s "
                      "encoder windows [windows, n_mels, 3000] it consumes (maf_mel_windows), produced inside the timed region"}
        del pipe
        torch.cuda.empty_cache()
        pipe, dt = run_pipe(2, True)
        e2e["all_to_host"] = {"value": audio_s / dt, "unit": "audio-s/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": pipe.h2d_bytes,
                              "d2h_bytes_per_step": pipe.d2h_bytes, "note": "log-mel copied to the host too (depth-2 overlap)"}
        hout, hmel, hl = pipe.h_out[0], pipe.h_mel[0], pipe.h_lufs[0]
        p0 = pipe.plans[0]
        for _ in range(2):
            p0.process_host(pipe.h_in[0], hout, hmel, hl)
        t0 = time.perf_counter()
        for _ in range(3):
            p0.process_host(pipe.h_in[0], hout, hmel, hl)
        e2e["sync_call_ms_per_step"] = (time.perf_counter() - t0) / 3 * 1e3
        del pipe, p0, hout, hmel, hl
        torch.cuda.empty_cache()

    # The drop-in's own call sequence (pipeline.py:331-347): WAV path in -> reduce_stationary_noise -> normalize_audio ->
    # normalize_loudness -> WAV path out, files on a RAM disk, one recording of the workload; plans are created per call.
    if e2e is not None and rank == 0 and world == 1:
        import shutil
        import tempfile
        from modular_audio_pipeline_b200 import GpuAudioPreprocessor, wavio
        d = tempfile.mkdtemp(prefix="maf_proto_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            src = os.path.join(d, "in.wav")
            wavio.write_pcm16(src, clips[0].cpu().contiguous(), SR)
            proto = {}
            for mode in ("sync", "deferred"):
                pre = GpuAudioPreprocessor(SR, os.path.join(d, "tmp_" + mode), device=dev, stationary=stationary, write_mode=mode)
                pre._fe = fe

                def protocol_once():
                    out = pre.normalize_loudness(pre.normalize_audio(pre.reduce_stationary_noise(src)))
                    if mode == "deferred":
                        torch.cuda.synchronize()             # results are on the device; the files complete in flush()
                    return out

                protocol_once()
                pre.flush()
                times = []
                for _ in range(4):
                    os.utime(src)                            # new mtime: the input is read from the file and uploaded again every time
                    t0 = time.perf_counter()
                    protocol_once()
                    times.append(time.perf_counter() - t0)
                    t1 = time.perf_counter()
                    pre.flush()
                    flush_ms = (time.perf_counter() - t1) * 1e3
                dtp = float(np.median(times))
                proto[mode] = {"ms_per_recording": dtp * 1e3, "value": seconds / dtp, "unit": "audio-s/s", "flush_ms_after": flush_ms}
                pre.close()
                del pre
            e2e["protocol_calls"] = dict(proto["sync"], deferred=proto["deferred"],
                                         what="GpuAudioPreprocessor.reduce_stationary_noise -> normalize_audio -> normalize_loudness on one "
                                              "recording of the workload: WAV path in (read + uploaded every time), three WAV files out on a RAM "
                                              "disk, one cached plan, stage outputs handed over on the device.  `sync` (default, the reference's "
                                              "contract): every call returns when its file is complete; `deferred`: calls return when the write is "
                                              "queued, files complete in flush() / remove_silence (flush_ms_after = what was still to be written)")
        finally:
            shutil.rmtree(d, ignore_errors=True)
        torch.cuda.empty_cache()

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            extras = bench_extras(fe, dev, args)
        except Exception as e:
            extras = {"extras_error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        pool = CpuPool(args.workload, args.cpu_sample)
        pool.step()                                        # warm-up (page-in, FFT plans)
        dts = [pool.step() for _ in range(2)]
        pool.close()
        dt = min(dts)
        cpu = {"value": pool.audio_s / dt, "unit": "audio-s/s", "cores": pool.n, "kind": "port",
               "sample": cpu_sample_text(pool, args.workload, args.cpu_sample) + f" ({dt:.2f} s wall per step, os.cpu_count {os.cpu_count()})"}

    time_sharded = None
    if world > 1 and not args.no_time_sharded:
        # BASELINE config 4 next to the file-sharded figure, so that the driver's 1 -> 8 scaling run times the mode that HAS
        # collectives too (every buffer of the file-sharded leg is released first)
        plan.close()
        plan = pcm = clips = None
        try:
            del mel, out_pcm
        except NameError:
            pass
        torch.cuda.empty_cache()
        try:
            time_sharded = measure_time_sharded(args, dev, world, rank, max(4, min(args.steps, 20)), max(args.warmup, 3))
        except Exception as e:
            time_sharded = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        line = {"metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": config_of(args.workload, world),
                "x_realtime": value, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": sampler.summary()}
        if time_sharded is not None:
            line["time_sharded"] = time_sharded
        if extras:
            line.update(extras)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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


def stage_bytes(plan, lengths, stationary, n_mels, padding):
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
    b["mel_finalize"] = 8 * n_mels * mel_frames
    return b


def _cpu_worker(conn, wl, sample_seconds, seed):
    """One host thread of the CPU baseline: owns one recording, runs the oracle front-end on it on request."""
    import torch as _torch
    _torch.set_num_threads(1)
    from modular_audio_pipeline_b200.synth import speech_noise
    from oracle import reference_path as O
    desc, seconds, per_gpu, stationary, n_mels, padding, meeting = WORKLOADS[wl]
    pcm = speech_noise(min(sample_seconds, seconds), seed, meeting=meeting).numpy()
    conn.send(len(pcm) / SR)
    while conn.recv() == "go":
        t0 = time.perf_counter()
        O.front_end(pcm, SR, stationary, n_mels, padding)
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload in ("cfg4", "cfg5"):
        args.workload = "cfg3"          # the CPU arm has no sharded / 48 kHz front door leg: 30 s stationary clips as in config 3
    desc, seconds, per_gpu, stationary, n_mels, padding, meeting = WORKLOADS[args.workload]
    sample = min(seconds, args.cpu_sample)
    pool = CpuPool(args.workload, sample)
    times = []
    for i in range(args.warmup + args.steps):
        dt = pool.step()
        if i >= args.warmup:
            times.append(dt)
    pool.close()
    ms = 1e3 * float(np.mean(times))
    value = pool.audio_s / (ms / 1e3)
    line = {"impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "sample": f"{pool.n} recordings of {sample:.0f} s per step, one per host thread"},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": pool.n, "kind": "port",
                             "sample": f"{pool.n} x {sample:.0f} s of the workload audio per step through oracle/reference_path.front_end, "
                                       f"one single-threaded process per host thread"},
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


def run_cfg4(args):
    """BASELINE config 4: ONE long recording (1 h per rank: 8 h on 8 GPUs) time-sharded across the ranks with 30 000-sample
    denoise halos and the six latency-class NCCL collectives of sharded.py (stationary denoise, LUFS, 80-mel)."""
    import torch.distributed as dist
    from modular_audio_pipeline_b200 import FrontEnd
    from modular_audio_pipeline_b200.sharded import Comm, CudaShardBackend, ShardGeometry, run_sharded
    from modular_audio_pipeline_b200.synth import speech_noise
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_init = world > 1
    if own_init:
        dist.init_process_group("nccl", device_id=dev)
    else:                                                     # the collectives still run (group of one)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    hour = 3600 * SR
    geo = ShardGeometry(hour * world, rank, world)
    sl = geo.local_slice()
    # every rank synthesises the samples it holds (own range + halos); the seams need not be continuous for timing
    local_pcm = speech_noise((sl.stop - sl.start) / SR, 2 + rank, device=dev)[: sl.stop - sl.start]
    fe = FrontEnd(dev)
    backend = CudaShardBackend(fe, geo, local_pcm, stationary=True, n_mels=80, mel_padding=480000)
    comm = Comm(None)
    for _ in range(max(args.warmup, 3)):
        run_sharded(backend, geo, stationary=True, comm=comm)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = backend.plan.launches
    ev0.record()
    for _ in range(args.steps):
        out, mel, lufs = run_sharded(backend, geo, stationary=True, comm=comm)
    ev1.record()
    torch.cuda.synchronize()
    sampler.stop_flag.set()
    dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    audio_s = 3600.0 * world
    if rank == 0:
        line = {"metric": "audio_seconds_per_second", "value": audio_s / (ms_step / 1e3), "unit": "audio-s/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"cfg4: one {world} h recording time-sharded over {world} GPU(s) (whole 600 000-sample chunks, 30 000-sample "
                                       "halos), stationary denoise + LUFS + 80-mel; NCCL: all-gather frame stats, SUM clip, MAX peak, "
                                       "all-gather edges / sub-block energies, MAX mel",
                           "seconds_per_rank": 3600.0, "l2": "working set exceeds the 126 MB L2", "parallelism": f"time-sharded x{world}"},
                "lufs": float(lufs.cpu().item()), "gpu_launches": backend.plan.launches - launches0, "clocks": sampler.summary()}
        emit(line)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg4", "cfg5"])
    ap.add_argument("--cpu-sample", type=float, default=120.0, help="seconds of audio EACH host thread of the CPU baseline processes per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
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
    sbytes = stage_bytes(plan, lengths, stationary, n_mels, padding)
    hbm_peak, peak_kind = peaks()
    dom = max(stage, key=lambda k: stage[k])
    achieved = sbytes[dom] / (stage[dom] * 1e-3) / 1e9 if stage[dom] > 0 else 0.0
    traffic, traffic_tag = ncu_traffic(args.workload, dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": f"profiles/{traffic_tag}_ncu.md" if traffic_tag else None, "peak_kind": peak_kind, "algorithmic_bytes_per_launch": sbytes[dom],
                "kernel_ms": stage[dom],
                "stages": {k: {"ms": round(stage[k], 4), "GBps": round(sbytes[k] / (stage[k] * 1e-3) / 1e9, 1) if stage[k] > 0 else 0.0}
                           for k in stage}}
    total_bytes = sum(sbytes.values())
    roofline["pipeline_GBps"] = total_bytes / (ms_step * 1e-3) / 1e9 * (1 if world == 1 else 1)

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
            pipe = HostPipeline(fe, lengths, depth=depth, mel_to_host=mel_to_host, stationary=stationary, n_mels=n_mels, mel_padding=padding)
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
                      f"step, depth-{args.e2e_depth} overlap; the log-mel is handed to the transcriber on the device"}
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
            pre = GpuAudioPreprocessor(SR, os.path.join(d, "tmp"), device=dev, stationary=stationary)

            def protocol_once():
                return pre.normalize_loudness(pre.normalize_audio(pre.reduce_stationary_noise(src)))

            protocol_once()
            t0 = time.perf_counter()
            for _ in range(3):
                protocol_once()
            torch.cuda.synchronize()
            dtp = (time.perf_counter() - t0) / 3
            e2e["protocol_calls"] = {"ms_per_recording": dtp * 1e3, "value": seconds / dtp, "unit": "audio-s/s",
                                     "what": "GpuAudioPreprocessor.reduce_stationary_noise -> normalize_audio -> normalize_loudness on one "
                                             "recording of the workload, WAV path in / WAV path out (three files written), RAM disk"}
            del pre
        finally:
            shutil.rmtree(d, ignore_errors=True)
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = min(seconds, args.cpu_sample)
        pool = CpuPool(args.workload, sample)
        pool.step()                                        # warm-up (page-in, FFT plans)
        dts = [pool.step() for _ in range(2)]
        pool.close()
        dt = min(dts)
        cpu = {"value": pool.audio_s / dt, "unit": "audio-s/s", "cores": pool.n, "kind": "port",
               "sample": f"{pool.n} x {sample:.0f} s of the same workload through oracle/reference_path.front_end, one single-threaded "
                         f"process per host thread ({dt:.2f} s wall per step, os.cpu_count {os.cpu_count()})"}

    if rank == 0:
        line = {"metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "recordings_per_gpu": per_gpu, "seconds_per_recording": seconds,
                           "l2": ("working set (STFT 4 KB/frame) exceeds the 126 MB L2; no explicit flush" if seconds * per_gpu >= 600 else
                                  "working set fits the 126 MB L2: a latency figure, not a roofline one"),
                           "parallelism": f"file-sharded x{world}"},
                "x_realtime": value, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": sampler.summary()}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

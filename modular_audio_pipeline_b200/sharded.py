"""Time-sharding of ONE long recording across the GPUs of a box (BASELINE config 4).

Shards are whole 600 000-sample denoise chunks, so every chunk keeps the frame grid noisereduce gives it; a boundary
is at the same time a multiple of the Whisper hop (160), of the loudness block hop (1 600) and of the noise-detector
hop (160).  Each rank holds its samples plus a 30 000-sample halo per side (the denoise padding, which the algorithm
defines) and the data path needs only latency-class collectives over NCCL / NVLink -- ONE per barrier of the algorithm, the
statistics that become available together travel together:

  1. all-gather  per-frame (energy, zero-crossing) statistics, packed            -> every rank selects the same noise clip
  2. all-reduce  (SUM) the clip's samples: every sample is owned by one rank, the buffer travels as int32 words holding
                 two int16 samples each                                           -> per-bin thresholds everywhere
  3. all-gather  [first 2 048 | last 2 048 denoised samples | local peak]         -> peak normalise, K-weighting warm-up /
                                                                                     STFT-400 halo
  4. all-gather  [local normalised peak | 100 ms sub-block energies]              -> gating, gain, limiter
  5. all-reduce  (MAX) the log-mel maximum                                        -> the global max - 8 clamp

(non-stationary mode needs no noise clip: exchanges 3-5 only).  The device side of every exchange is a pack / unpack kernel of
libmaf_b200.so on preallocated buffers; nothing in a step allocates or synchronises with the host, so a whole step can be
replayed as one CUDA graph (``CudaShardBackend.capture``).

``ShardGeometry`` is pure Python.  ``run_sharded`` drives a *backend* (the CUDA plan in production,
``CudaShardBackend``; tests drive the same logic with a CPU stand-in over gloo) and a ``torch.distributed`` group.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

CHUNK = 600000
PAD = 30000
HALO = 2048
W_HOP = 160


@dataclass
class ShardGeometry:
    n_global: int
    rank: int
    world: int
    sample_rate: int = 16000

    def __post_init__(self):
        if self.n_global <= CHUNK:
            raise ValueError("a recording of one denoise chunk cannot be time-sharded")
        n_chunks = -(-self.n_global // CHUNK)
        if self.world > n_chunks:
            raise ValueError(f"{self.world} ranks but only {n_chunks} chunks")
        base, extra = divmod(n_chunks, self.world)
        counts = [base + (1 if r < extra else 0) for r in range(self.world)]
        # keep the last shard at least HALO long so its neighbour's halo exists: move a chunk if the tail is tiny
        starts = [sum(counts[:r]) * CHUNK for r in range(self.world + 1)]
        starts[-1] = self.n_global
        if self.world > 1 and starts[-1] - starts[-2] < HALO:
            raise ValueError("last shard shorter than the halo; use fewer ranks")
        self.bounds = starts
        self.own_start, self.own_end = starts[self.rank], starts[self.rank + 1]
        self.buf_start = max(0, self.own_start - PAD)
        self.buf_end = min(self.n_global, self.own_end + PAD)
        hop = int(self.sample_rate * 0.010)
        flen = int(self.sample_rate * 0.025)
        fg = (self.n_global - flen) // hop + 1 if self.n_global >= flen else 0
        self.frames_global = fg
        self.frame_counts = [max(0, min(fg, -(-starts[r + 1] // hop)) - (-(-starts[r] // hop))) for r in range(self.world)]
        tile = int(0.4 * 0.25 * self.sample_rate)
        self.tile_len = tile
        self.tile_counts = [-(-(starts[r + 1] - starts[r]) // tile) for r in range(self.world)]
        T = self.n_global / self.sample_rate
        self.n_blocks = int(round((T - 0.4) / (0.4 * 0.25)) + 1)

    @property
    def own_len(self) -> int:
        return self.own_end - self.own_start

    @property
    def origin(self) -> int:
        return self.own_start - self.buf_start

    @property
    def lo_avail(self) -> int:
        return min(HALO, self.own_start)

    @property
    def hi_avail(self) -> int:
        return min(HALO, self.n_global - self.own_end)

    def local_slice(self) -> slice:
        """Samples of the full recording this rank must hold: own range plus the 30 000-sample denoise halo."""
        return slice(self.buf_start, self.buf_end)

    def mel_frames(self, padding: int) -> Tuple[int, int]:
        """(first global frame, number of frames) this rank writes."""
        total = (self.n_global + padding) // W_HOP
        f0 = self.own_start // W_HOP
        cnt = total - f0 if self.rank == self.world - 1 else self.own_len // W_HOP
        return f0, cnt


class Comm:
    """The collectives run_sharded needs.  Default: torch.distributed on the tensors' own device (NCCL over NVLink on
    GPUs, gloo on CPU).  ``via_cpu`` stages device tensors through host memory so that a gloo group can carry them
    (used to exercise the multi-rank path on a box with fewer GPUs than ranks)."""

    def __init__(self, group=None, via_cpu: bool = False):
        self.group, self.via_cpu = group, via_cpu
        self.calls = 0

    def all_reduce(self, t: torch.Tensor, op) -> None:
        self.calls += 1
        if self.via_cpu and t.is_cuda:
            h = t.cpu()
            dist.all_reduce(h, op=op, group=self.group)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=op, group=self.group)

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor) -> None:
        self.calls += 1
        if self.via_cpu and inp.is_cuda:
            ho, hi = torch.empty(out.shape, dtype=out.dtype), inp.cpu()
            dist.all_gather_into_tensor(ho, hi, group=self.group)
            out.copy_(ho)
        else:
            dist.all_gather_into_tensor(out, inp, group=self.group)


class NullComm(Comm):
    """Times a step WITHOUT its collectives (bench.py: collective cost = step - this).  Results are meaningless."""

    def all_reduce(self, t, op) -> None:
        self.calls += 1

    def all_gather(self, out, inp) -> None:
        self.calls += 1


def run_sharded(backend, geo: ShardGeometry, group=None, stationary: bool = True, comm: Optional[Comm] = None):
    """Drive one shard through the pipeline.  Returns (own int16 PCM, own log-mel [n_mels, frames], LUFS).

    The backend packs what it contributes to an exchange into a flat buffer of a length that is the same on every rank and
    unpacks the gathered buffer itself: ``*_contrib()`` -> (send, recv) tensors, ``apply_*()`` consumes recv."""
    comm = comm or Comm(group)
    if stationary:
        send, recv = backend.stats_contrib()                       # 1. frame statistics of the own frames
        comm.all_gather(recv, send)
        backend.select_clip()                                      #    identical selection on every rank
        clip = backend.clip_contrib()                              # 2. own samples of the selected range, zeros elsewhere
        comm.all_reduce(clip, dist.ReduceOp.SUM)
        backend.set_noise()
    backend.denoise()
    send, recv = backend.edge_contrib()                            # 3. denoised edges + local peak
    comm.all_gather(recv, send)
    backend.apply_edges()
    send, recv = backend.energy_contrib()                          # 4. K-weighted sub-block energies + normalised peak
    comm.all_gather(recv, send)
    backend.gate()
    backend.logmel()                                               # 5. gain + log-mel of the own frames, global maximum, clamp
    comm.all_reduce(backend.mel_max(), dist.ReduceOp.MAX)
    return backend.finalize()


class CudaShardBackend:
    """The CUDA plan of one shard behind the interface ``run_sharded`` drives.  Every buffer is allocated here, once."""

    def __init__(self, fe, geo: ShardGeometry, pcm_local: torch.Tensor, stationary: bool = True, n_mels: int = 80,
                 mel_padding: int = 0, target_lufs: float = -16.0, peak_normalize: bool = True,
                 noise_sample_duration_s: float = 0.5):
        from ._capi import ItemDesc
        self.fe, self.geo = fe, geo
        self.stationary, self.n_mels, self.padding = stationary, n_mels, mel_padding
        self.target_lufs, self.dur = target_lufs, noise_sample_duration_s
        dev = fe.device
        world = geo.world
        desc = ItemDesc(geo.buf_end - geo.buf_start, geo.origin, geo.own_len, geo.own_start, geo.n_global, geo.lo_avail, geo.hi_avail)
        self.plan = fe.plan([], sample_rate=geo.sample_rate, stationary=stationary, external_noise=stationary,
                            peak_normalize=peak_normalize, target_lufs=target_lufs, n_mels=n_mels, mel_padding=mel_padding,
                            shards=[desc])
        self.pcm = self.plan.pack([pcm_local.to(dev)])
        self.gate_buf = torch.zeros(32, dtype=torch.uint8, device=dev)
        self.out = torch.empty(max(self.plan.total_out, 8), dtype=torch.int16, device=dev)
        self.mel = torch.empty(max(self.plan.total_mel, 4), dtype=torch.float32, device=dev)
        _ = self.plan.workspace
        lib = fe._lib
        # 1. frame statistics: [energy bits | zero crossings] padded to the largest shard
        self.m_frames = max(max(geo.frame_counts), 1)
        self.stats_send = torch.zeros(2 * self.m_frames, dtype=torch.int32, device=dev)
        self.stats_recv = torch.empty(world * 2 * self.m_frames, dtype=torch.int32, device=dev)
        self.frame_counts = torch.tensor(geo.frame_counts, dtype=torch.int32, device=dev)
        self.n_frames_all = int(sum(geo.frame_counts))
        self.e_all = torch.empty(max(self.n_frames_all, 1), dtype=torch.float32, device=dev)
        self.z_all = torch.empty(max(self.n_frames_all, 1), dtype=torch.int32, device=dev)
        self.seg = torch.zeros(2, dtype=torch.int32, device=dev)
        # 2. noise clip: CHUNK int16 samples = CHUNK / 2 int32 words on the wire
        self.clip = torch.zeros(CHUNK, dtype=torch.int16, device=dev)
        # 3. edges
        self.edge_bytes = int(lib.maf_sharded_edge_bytes())
        self.edge_send = torch.zeros(self.edge_bytes, dtype=torch.uint8, device=dev)
        self.edge_recv = torch.empty(world * self.edge_bytes, dtype=torch.uint8, device=dev)
        # 4. energies
        self.m_tiles = max(max(geo.tile_counts), 1)
        self.en_send = torch.zeros(1 + self.m_tiles, dtype=torch.float64, device=dev)
        self.en_recv = torch.empty(world * (1 + self.m_tiles), dtype=torch.float64, device=dev)
        self.tile_counts = torch.tensor(geo.tile_counts, dtype=torch.int32, device=dev)
        self.sub_all = torch.empty(max(int(sum(geo.tile_counts)), 1), dtype=torch.float64, device=dev)
        self._mel_max = self.plan.buffer("mel_max", torch.int32)
        self._graph = None

    def _st(self):
        return torch.cuda.current_stream(self.fe.device).cuda_stream

    def _call(self, fn, *args):
        from ._capi import check
        with torch.cuda.device(self.fe.device):
            check(fn(*args))

    # 1 -------------------------------------------------------------------------------------------------------------
    def stats_contrib(self):
        lib = self.fe._lib
        send = self.stats_send                                     # maf_frame_stats writes the two arrays straight into the send buffer
        self._call(lib.maf_frame_stats, self.plan._h, self.pcm.data_ptr(), send.data_ptr(), send.data_ptr() + 4 * self.m_frames, self._st())
        return send, self.stats_recv

    def select_clip(self):
        lib, fe, g = self.fe._lib, self.fe, self.geo
        m = self.m_frames
        self._call(lib.maf_gather_compact, fe._h, self.stats_recv.data_ptr(), 4, g.world, 2 * m, 0, self.frame_counts.data_ptr(),
                   self.e_all.data_ptr(), self._st())
        self._call(lib.maf_gather_compact, fe._h, self.stats_recv.data_ptr(), 4, g.world, 2 * m, m, self.frame_counts.data_ptr(),
                   self.z_all.data_ptr(), self._st())
        self._call(lib.maf_select_clip, fe._h, self.e_all.data_ptr(), self.z_all.data_ptr(), self.n_frames_all, g.n_global,
                   g.sample_rate, self.dur, self.seg.data_ptr(), self._st())

    # 2 -------------------------------------------------------------------------------------------------------------
    def clip_contrib(self) -> torch.Tensor:
        self._call(self.fe._lib.maf_sharded_clip_contrib, self.plan._h, self.pcm.data_ptr(), self.seg.data_ptr(), self.clip.data_ptr(), self._st())
        return self.clip.view(torch.int32)

    def set_noise(self):
        self._call(self.fe._lib.maf_set_noise_clips_dev, self.plan._h, self.seg.data_ptr(), self._st())

    def denoise(self):
        lib = self.fe._lib
        self._call(lib.maf_sharded_denoise, self.plan._h, self.pcm.data_ptr(), self.clip.data_ptr() if self.stationary else None,
                   self.plan.workspace.data_ptr(), self._st())

    # 3 -------------------------------------------------------------------------------------------------------------
    def edge_contrib(self):
        self._call(self.fe._lib.maf_sharded_pack_edges, self.plan._h, self.plan.workspace.data_ptr(), self.edge_send.data_ptr(), self._st())
        return self.edge_send, self.edge_recv

    def apply_edges(self):
        self._call(self.fe._lib.maf_sharded_unpack_edges, self.plan._h, self.plan.workspace.data_ptr(), self.edge_recv.data_ptr(),
                   self.geo.world, self.geo.rank, self._st())

    # 4 -------------------------------------------------------------------------------------------------------------
    def energy_contrib(self):
        lib = self.fe._lib
        self._call(lib.maf_sharded_energy, self.plan._h, self.plan.workspace.data_ptr(), self._st())
        self._call(lib.maf_sharded_pack_energy, self.plan._h, self.plan.workspace.data_ptr(), self.en_send.data_ptr(), self.m_tiles, self._st())
        return self.en_send, self.en_recv

    def gate(self):
        lib, g = self.fe._lib, self.geo
        self._call(lib.maf_sharded_unpack_energy, self.plan._h, self.plan.workspace.data_ptr(), self.en_recv.data_ptr(), g.world, self.m_tiles,
                   self.tile_counts.data_ptr(), self.sub_all.data_ptr(), self._st())
        self._call(lib.maf_gate, self.fe._h, self.sub_all.data_ptr(), self.sub_all.numel(), g.n_blocks, g.sample_rate,
                   self.target_lufs, self.plan.buffer("peak_norm", torch.int32).data_ptr(), self.gate_buf.data_ptr(), self._st())

    # 5 -------------------------------------------------------------------------------------------------------------
    def logmel(self):
        lib = self.fe._lib
        self._call(lib.maf_sharded_logmel, self.plan._h, self.plan.workspace.data_ptr(), self.gate_buf.data_ptr(), self.out.data_ptr(),
                   self.mel.data_ptr(), self._st())

    def mel_max(self):
        return self._mel_max

    def finalize(self):
        lib = self.fe._lib
        self._call(lib.maf_mel_finalize, self.plan._h, self.plan.workspace.data_ptr(), self.mel.data_ptr(), self._st())
        lufs = self.gate_buf[:8].view(torch.float64)
        return self.plan.split_out(self.out)[0], self.plan.split_mel(self.mel)[0], lufs

    # whole-step CUDA graph -------------------------------------------------------------------------------------------
    def capture(self, comm: Optional[Comm] = None):
        """Capture one whole sharded step -- kernels AND collectives -- into a CUDA graph (NCCL operations are capturable; the
        step allocates nothing and never touches the host).  Returns a callable that replays it and returns finalize()'s tensors."""
        comm = comm or Comm(None)
        dev = self.fe.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                                     # warm-up outside capture (lazy function attributes, NCCL channels)
                run_sharded(self, self.geo, stationary=self.stationary, comm=comm)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ret = run_sharded(self, self.geo, stationary=self.stationary, comm=comm)
        self._graph = graph
        del graph

        def replay():
            self._graph.replay()
            return ret
        return replay

    def release_graph(self) -> None:
        """Destroy the captured step.  A live CUDA graph that holds NCCL operations keeps the communicator referenced:
        ``dist.destroy_process_group()`` blocks until the graph is gone, so call this (or ``close``) first."""
        if self._graph is not None:
            torch.cuda.synchronize(self.fe.device)
            self._graph.reset()
            self._graph = None

    def close(self) -> None:
        self.release_graph()
        self.plan.close()

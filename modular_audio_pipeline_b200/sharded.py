"""Time-sharding of ONE long recording across the GPUs of a box (BASELINE config 4).

Shards are whole 600 000-sample denoise chunks, so every chunk keeps the frame grid noisereduce gives it; a boundary
is at the same time a multiple of the Whisper hop (160), of the loudness block hop (1 600) and of the noise-detector
hop (160).  Each rank holds its samples plus a 30 000-sample halo per side (the denoise padding, which the algorithm
defines) and the data path needs only latency-class collectives over NCCL / NVLink:

  1. all-gather of the per-frame (energy, zero-crossing) statistics  -> every rank selects the same noise clip
  2. all-reduce(SUM) of the clip's samples (each rank fills the part it owns)      -> per-bin thresholds everywhere
  3. all-reduce(MAX) of the denoised peak                                         -> peak normalise
  4. all-gather of 2 x 2 048 denoised edge samples                                -> K-weighting warm-up / STFT-400 halo
  5. all-gather of the 100 ms sub-block energies + all-reduce(MAX) of the normalised peak -> gating, gain, limiter
  6. all-reduce(MAX) of the log-mel maximum                                       -> the global max - 8 clamp

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

    def all_reduce(self, t: torch.Tensor, op) -> None:
        if self.via_cpu and t.is_cuda:
            h = t.cpu()
            dist.all_reduce(h, op=op, group=self.group)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=op, group=self.group)

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor) -> None:
        if self.via_cpu and inp.is_cuda:
            ho, hi = torch.empty(out.shape, dtype=out.dtype), inp.cpu()
            dist.all_gather_into_tensor(ho, hi, group=self.group)
            out.copy_(ho)
        else:
            dist.all_gather_into_tensor(out, inp, group=self.group)


def _gather_var(x: torch.Tensor, counts: List[int], comm: Comm) -> torch.Tensor:
    """all-gather of per-rank 1-D tensors of different lengths (padded to the maximum), concatenated in rank order."""
    world = len(counts)
    m = max(max(counts), 1)
    buf = torch.zeros(m, dtype=x.dtype, device=x.device)
    buf[: x.numel()] = x
    out = torch.empty(world * m, dtype=x.dtype, device=x.device)
    comm.all_gather(out, buf)
    return torch.cat([out[r * m: r * m + counts[r]] for r in range(world)])


def run_sharded(backend, geo: ShardGeometry, group=None, stationary: bool = True, comm: Optional[Comm] = None):
    """Drive one shard through the pipeline.  Returns (own int16 PCM, own log-mel [n_mels, frames], LUFS)."""
    world, rank = geo.world, geo.rank
    comm = comm or Comm(group)
    if stationary:
        # 1. noise clip: gather frame statistics, select identically on every rank
        e, z = backend.frame_stats()
        e_all = _gather_var(e, geo.frame_counts, comm)
        z_all = _gather_var(z, geo.frame_counts, comm)
        if hasattr(backend, "clip_contrib"):
            # device path: the selected range never visits the host (no stream sync in the whole step)
            seg = backend.select_clip_dev(e_all, z_all)
            clip = backend.clip_contrib(seg)                       # int32 [CHUNK]: own samples of the range, zeros elsewhere
            comm.all_reduce(clip, dist.ReduceOp.SUM)
            backend.set_noise_dev(clip.to(torch.int16), seg)
        else:
            gs, ge = backend.select_clip(e_all, z_all)             # global sample range (host ints, one sync)
            clip_len = min(ge - gs, CHUNK)
            # 2. every rank contributes the clip samples it owns; the sum is the clip
            clip = torch.zeros(clip_len, dtype=torch.int32, device=e.device)
            lo, hi = max(gs, geo.own_start), min(gs + clip_len, geo.own_end)
            if hi > lo:
                clip[lo - gs: hi - gs] = backend.own_samples(lo - geo.own_start, hi - geo.own_start).to(torch.int32)
            comm.all_reduce(clip, dist.ReduceOp.SUM)
            backend.set_noise(clip.to(torch.int16))
    # 3. denoise the own chunks; global peak
    backend.denoise()
    comm.all_reduce(backend.peak(), dist.ReduceOp.MAX)
    # 4. exchange 2 x HALO denoised edge samples with the neighbours
    edges = backend.den_edges()                                     # [2, HALO] int16: first / last own samples
    all_edges = torch.empty(world * 2 * HALO, dtype=torch.int16, device=edges.device)
    # neither NCCL nor gloo has an int16 type: move the edges as bytes
    comm.all_gather(all_edges.view(torch.uint8), edges.reshape(-1).contiguous().view(torch.uint8))
    all_edges = all_edges.view(world, 2, HALO)
    left = all_edges[rank - 1, 1] if rank > 0 else None             # left neighbour's tail
    right = all_edges[rank + 1, 0] if rank < world - 1 else None    # right neighbour's head
    backend.set_den_halo(left, right)
    # 5. loudness: local sub-block energies -> gathered, gated identically everywhere
    sub = backend.energy()
    comm.all_reduce(backend.peak_norm(), dist.ReduceOp.MAX)
    sub_all = _gather_var(sub, geo.tile_counts, comm)
    backend.gate(sub_all, geo.n_blocks)
    # 6. gain + log-mel of the own frames, global maximum, clamp
    backend.logmel()
    comm.all_reduce(backend.mel_max(), dist.ReduceOp.MAX)
    return backend.finalize()


class CudaShardBackend:
    """The CUDA plan of one shard behind the interface ``run_sharded`` drives."""

    def __init__(self, fe, geo: ShardGeometry, pcm_local: torch.Tensor, stationary: bool = True, n_mels: int = 80,
                 mel_padding: int = 0, target_lufs: float = -16.0, peak_normalize: bool = True,
                 noise_sample_duration_s: float = 0.5):
        from ._capi import ItemDesc
        self.fe, self.geo = fe, geo
        self.stationary, self.n_mels, self.padding = stationary, n_mels, mel_padding
        self.target_lufs, self.dur = target_lufs, noise_sample_duration_s
        desc = ItemDesc(geo.buf_end - geo.buf_start, geo.origin, geo.own_len, geo.own_start, geo.n_global, geo.lo_avail, geo.hi_avail)
        self.plan = fe.plan([], sample_rate=geo.sample_rate, stationary=stationary, external_noise=stationary,
                            peak_normalize=peak_normalize, target_lufs=target_lufs, n_mels=n_mels, mel_padding=mel_padding,
                            shards=[desc])
        self.pcm = self.plan.pack([pcm_local.to(fe.device)])
        self.noise: Optional[torch.Tensor] = None
        self.gate_buf = torch.zeros(32, dtype=torch.uint8, device=fe.device)
        self.out = torch.empty(max(self.plan.total_out, 8), dtype=torch.int16, device=fe.device)
        self.mel = torch.empty(max(self.plan.total_mel, 4), dtype=torch.float32, device=fe.device)
        _ = self.plan.workspace

    def _st(self):
        return torch.cuda.current_stream(self.fe.device).cuda_stream

    def _call(self, fn, *args):
        from ._capi import check
        with torch.cuda.device(self.fe.device):
            check(fn(*args))

    def frame_stats(self):
        return self.plan.frame_stats(self.pcm)

    def select_clip_dev(self, e_all, z_all):
        seg = torch.zeros(2, dtype=torch.int32, device=self.fe.device)
        lib = self.fe._lib
        self._e_all, self._z_all = e_all, z_all                 # keep alive until the kernel has run
        self._call(lib.maf_select_clip, self.fe._h, e_all.data_ptr(), z_all.data_ptr(), e_all.numel(), self.geo.n_global,
                   self.geo.sample_rate, self.dur, seg.data_ptr(), self._st())
        return seg

    def select_clip(self, e_all, z_all):
        s = self.select_clip_dev(e_all, z_all).cpu()
        return int(s[0]), int(s[1])

    def clip_contrib(self, seg: torch.Tensor) -> torch.Tensor:
        clip = torch.empty(CHUNK, dtype=torch.int32, device=self.fe.device)
        self._call(self.fe._lib.maf_sharded_clip_contrib, self.plan._h, self.pcm.data_ptr(), seg.data_ptr(), clip.data_ptr(), self._st())
        return clip

    def set_noise_dev(self, clip: torch.Tensor, seg: torch.Tensor):
        self.noise, self._seg = clip.contiguous(), seg
        self._call(self.fe._lib.maf_set_noise_clips_dev, self.plan._h, seg.data_ptr(), self._st())

    def own_samples(self, a: int, b: int) -> torch.Tensor:
        o = self.plan.layout[0].pcm_off + self.geo.origin
        return self.pcm[o + a: o + b]

    def set_noise(self, clip: torch.Tensor):
        self.noise = clip.contiguous()
        self.plan.set_noise_clips([0], [clip.numel()])

    def denoise(self):
        lib = self.fe._lib
        self._call(lib.maf_sharded_denoise, self.plan._h, self.pcm.data_ptr(), None if self.noise is None else self.noise.data_ptr(),
                   self.plan.workspace.data_ptr(), self._st())

    def peak(self):
        return self.plan.buffer("peak", torch.int32)

    def peak_norm(self):
        return self.plan.buffer("peak_norm", torch.int32)

    def mel_max(self):
        return self.plan.buffer("mel_max", torch.int32)

    def den_edges(self):
        den = self.plan.buffer("den16", torch.int16)
        o, n = self.plan.den_offset(0), self.geo.own_len
        head = torch.zeros(HALO, dtype=torch.int16, device=den.device)
        tail = torch.zeros(HALO, dtype=torch.int16, device=den.device)
        k = min(HALO, n)
        head[:k] = den[o: o + k]
        tail[HALO - k:] = den[o + n - k: o + n]
        return torch.stack([head, tail])

    def set_den_halo(self, left, right):
        den = self.plan.buffer("den16", torch.int16)
        o, n = self.plan.den_offset(0), self.geo.own_len
        if left is not None:
            den[o - HALO: o] = left
        if right is not None:
            den[o + n: o + n + HALO] = right

    def energy(self):
        lib = self.fe._lib
        self._call(lib.maf_sharded_energy, self.plan._h, self.plan.workspace.data_ptr(), self._st())
        return self.plan.buffer("sub_energy", torch.float64)[: self.geo.tile_counts[self.geo.rank]]

    def gate(self, sub_all: torch.Tensor, n_blocks: int):
        lib = self.fe._lib
        self.sub_all = sub_all.contiguous()
        self._call(lib.maf_gate, self.fe._h, self.sub_all.data_ptr(), self.sub_all.numel(), n_blocks, self.geo.sample_rate,
                   self.target_lufs, self.peak_norm().data_ptr(), self.gate_buf.data_ptr(), self._st())

    def logmel(self):
        lib = self.fe._lib
        self._call(lib.maf_sharded_logmel, self.plan._h, self.plan.workspace.data_ptr(), self.gate_buf.data_ptr(), self.out.data_ptr(),
                   self.mel.data_ptr(), self._st())

    def finalize(self):
        lib = self.fe._lib
        self._call(lib.maf_mel_finalize, self.plan._h, self.plan.workspace.data_ptr(), self.mel.data_ptr(), self._st())
        lufs = self.gate_buf[:8].view(torch.float64)
        return self.plan.split_out(self.out)[0], self.plan.split_mel(self.mel)[0], lufs

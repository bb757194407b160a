"""World-size-2 gloo test of the time-sharding host logic (modular_audio_pipeline_b200/sharded.py) on CPU.

``run_sharded`` is driven with a CPU stand-in backend built from the oracle's primitives, so what is under test is
the geometry (chunk-aligned shards, 30 000-sample denoise halo, 2 048-sample denoised halo), the order and content
of the collectives, and that stitching the shards reproduces the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import scipy.signal as sps
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modular_audio_pipeline_b200.sharded import CHUNK, HALO, PAD, ShardGeometry, run_sharded
from modular_audio_pipeline_b200.synth import speech_noise
from oracle import reference_path as O

SR = 16000


class OracleShardBackend:
    """Shard-local processing with the oracle's float64 primitives (test-only; mirrors CudaShardBackend's interface)."""

    def __init__(self, geo, pcm_local, n_mels=80, padding=0):
        self.geo, self.x, self.n_mels, self.padding = geo, pcm_local.astype(np.int16), n_mels, padding
        self._peak = torch.zeros(1, dtype=torch.int32)
        self._peak_norm = torch.zeros(1, dtype=torch.int32)
        self._mel_max = torch.zeros(1, dtype=torch.float32)
        self.left = self.right = None

    # 1. frame statistics: [energy bits | zero crossings], padded to the largest shard, gathered, un-padded
    def stats_contrib(self):
        g = self.geo
        hop, flen = 160, 400
        cnt = g.frame_counts[g.rank]
        seg = self.x[g.origin: g.origin + (cnt - 1) * hop + flen].astype(np.float32)
        e, z = O.frame_stats(seg, SR)
        m = max(g.frame_counts)
        send = torch.zeros(2 * m, dtype=torch.int32)
        send[:cnt] = torch.from_numpy(e[:cnt].astype(np.float32)).view(torch.int32)
        send[m: m + cnt] = torch.from_numpy(np.round(z[:cnt] * flen).astype(np.int32))
        self._stats = torch.empty(g.world * 2 * m, dtype=torch.int32)
        return send, self._stats

    def select_clip(self):
        g = self.geo
        m = max(g.frame_counts)
        rows = self._stats.view(g.world, 2, m)
        e_all = torch.cat([rows[r, 0, :g.frame_counts[r]] for r in range(g.world)]).view(torch.float32)
        z_all = torch.cat([rows[r, 1, :g.frame_counts[r]] for r in range(g.world)])
        segs = O.segments_from_stats(e_all.numpy().astype(np.float64), z_all.numpy() / 400.0, 160, SR)
        self.seg = max(segs, key=lambda s: s[1] - s[0]) if segs else (0, min(int(SR * 0.5), g.n_global))

    # 2. the clip: own samples of the range as int16 (zeros elsewhere), summed as int32 words of two samples
    def clip_contrib(self):
        g = self.geo
        gs, ge = self.seg
        n = min(ge - gs, CHUNK)
        clip = np.zeros(CHUNK, dtype=np.int16)
        lo, hi = max(gs, g.own_start), min(gs + n, g.own_end)
        if hi > lo:
            clip[lo - gs: hi - gs] = self.x[g.origin + lo - g.own_start: g.origin + hi - g.own_start]
        self._clip = torch.from_numpy(clip)
        return self._clip.view(torch.int32)

    def set_noise(self):
        gs, ge = self.seg
        self.thresh = O.stationary_threshold(self._clip.numpy()[: min(ge - gs, CHUNK)].astype(np.float32))

    def denoise(self):
        g = self.geo
        K = O.smoothing_filter(SR)
        y = self.x.astype(np.float32)
        out = np.zeros(g.own_len)
        for k in range(-(-g.own_len // CHUNK)):
            s = g.origin + k * CHUNK                     # local index of the chunk start
            keep = min(CHUNK, g.own_len - k * CHUNK)
            c = np.zeros(CHUNK + 2 * PAD)
            lo, hi = max(s - PAD, 0), min(s + CHUNK + PAD, len(y))
            c[lo - (s - PAD): hi - (s - PAD)] = y[lo:hi]
            out[k * CHUNK: k * CHUNK + keep] = O._filter_stationary(c, self.thresh, K, 0.8)[PAD: PAD + keep]
        self.den = O.to_int16_trunc(out.astype(np.float32))
        self._peak[0] = int(np.abs(self.den.astype(np.int32)).max())

    def mel_max(self): return self._mel_max

    # 3. edges: [first HALO | last HALO denoised samples | local peak]
    def edge_contrib(self):
        g = self.geo
        send = torch.zeros(2 * HALO + 2, dtype=torch.int32)
        send[:HALO] = torch.from_numpy(self.den[:HALO].astype(np.int32))
        send[HALO: 2 * HALO] = torch.from_numpy(self.den[-HALO:].astype(np.int32))
        send[2 * HALO] = int(self._peak[0])
        self._edges = torch.empty(g.world * (2 * HALO + 2), dtype=torch.int32)
        return send, self._edges

    def apply_edges(self):
        g = self.geo
        rows = self._edges.view(g.world, 2 * HALO + 2)
        self._peak[0] = int(rows[:, 2 * HALO].max())
        self.left = rows[g.rank - 1, HALO: 2 * HALO].numpy().astype(np.int16) if g.rank > 0 else None
        self.right = rows[g.rank + 1, :HALO].numpy().astype(np.int16) if g.rank < g.world - 1 else None

    def _norm_with_halo(self):
        parts = [p for p in (self.left, self.den, self.right) if p is not None]
        full = np.concatenate(parts)
        f = O.peak_normalize_factor(int(self._peak[0]))
        v = np.floor(np.clip(full.astype(np.float64) * f, -32768.0, 32767.0)).astype(np.int16)
        return v, (0 if self.left is None else len(self.left))

    # 4. [local normalised peak | sub-block energies]
    def energy_contrib(self):
        g = self.geo
        v, off = self._norm_with_halo()
        x = v.astype(np.float32) / np.float32(32768.0)
        d = x.copy()
        for (b, a) in O.k_weighting_coeffs(SR):
            d[:] = sps.lfilter(b, a, d)                  # starts from zero state HALO samples early: the warm-up
        own = d[off: off + g.own_len]
        nt = g.tile_counts[g.rank]
        sub = np.array([np.sum(np.square(own[t * 1600:(t + 1) * 1600]).astype(np.float64)) for t in range(nt)])
        m = max(g.tile_counts)
        send = torch.zeros(1 + m, dtype=torch.float64)
        send[0] = float(np.abs(v[off: off + g.own_len].astype(np.int32)).max())
        send[1: 1 + nt] = torch.from_numpy(sub)
        self._en = torch.empty(g.world * (1 + m), dtype=torch.float64)
        return send, self._en

    def gate(self):
        g = self.geo
        m = max(g.tile_counts)
        rows = self._en.view(g.world, 1 + m)
        self._peak_norm[0] = int(rows[:, 0].max())
        sb = torch.cat([rows[r, 1: 1 + g.tile_counts[r]] for r in range(g.world)]).numpy()
        z = np.array([sb[j:j + 4].sum() / 6400.0 for j in range(g.n_blocks)])
        self.lufs = O.gate_loudness(z)
        self.gain = np.float32(10.0 ** ((-16.0 - self.lufs) / 20.0))
        p = self.gain * np.float32(int(self._peak_norm[0]) / 32768.0)
        self.limp = p if p > 1.0 else None

    def logmel(self):
        import torch as T
        v, off = self._norm_with_halo()
        yv = self.gain * (v.astype(np.float32) / np.float32(32768.0))
        if self.limp is not None:
            yv = yv / self.limp
        q = np.clip(yv * np.float32(32768), -32768, 32767).astype(np.int16)
        g = self.geo
        self.out = q[off: off + g.own_len]
        a = q.astype(np.float32) / np.float32(32768.0)
        f0, cnt = g.mel_frames(self.padding)
        # frames centred on 160 f: gather 400-sample windows from (halo + own + halo), reflecting only at true ends
        gidx = (np.arange(f0, f0 + cnt)[:, None] * 160 - 200 + np.arange(400)[None, :])
        npad = g.n_global + self.padding
        gidx = np.abs(gidx)
        gidx = np.where(gidx >= npad, 2 * (npad - 1) - gidx, gidx)
        local = gidx - (g.own_start - off)
        valid = (gidx < g.n_global)
        frames = np.where(valid, a[np.clip(local, 0, len(a) - 1)], 0.0).astype(np.float32)
        win = T.hann_window(400).numpy()
        spec = np.abs(np.fft.rfft(frames.astype(np.float64) * win, axis=1)) ** 2
        mel = O.mel_filters(self.n_mels).astype(np.float64) @ spec.T
        self.logspec = np.log10(np.maximum(mel, 1e-10)).astype(np.float32)
        self._mel_max[0] = float(self.logspec.max())

    def finalize(self):
        ls = np.maximum(self.logspec, float(self._mel_max[0]) - 8.0)
        return self.out, ((ls + 4.0) / 4.0).astype(np.float32), self.lufs


def _worker(rank, world, port, pcm, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        geo = ShardGeometry(len(pcm), rank, world)
        backend = OracleShardBackend(geo, pcm[geo.local_slice()])
        out, mel, lufs = run_sharded(backend, geo)
        results[rank] = (geo.own_start, out, mel, lufs)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_geometry():
    g0, g1 = ShardGeometry(1500000, 0, 2), ShardGeometry(1500000, 1, 2)
    assert (g0.own_start, g0.own_end, g0.buf_start, g0.buf_end) == (0, 1200000, 0, 1230000)
    assert (g1.own_start, g1.own_end, g1.buf_start, g1.buf_end, g1.origin) == (1200000, 1500000, 1170000, 1500000, 30000)
    assert sum(g0.frame_counts) == (1500000 - 400) // 160 + 1 and sum(g0.tile_counts) == -(-1500000 // 1600)
    assert g0.mel_frames(0) == (0, 7500) and g1.mel_frames(0) == (7500, 1875)
    assert (g0.lo_avail, g0.hi_avail, g1.lo_avail, g1.hi_avail) == (0, HALO, HALO, 0)
    with pytest.raises(ValueError):
        ShardGeometry(500000, 0, 2)


@pytest.mark.timeout(600)
def test_two_rank_gloo_matches_unsharded_oracle():
    pcm = speech_noise(1500000 / SR, 8).numpy()
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), pcm, results), nprocs=world, join=True)
    ref = O.front_end(pcm, SR, True, 80, 0)
    out = np.concatenate([results[r][1] for r in range(world)])
    mel = np.concatenate([results[r][2] for r in range(world)], axis=1)
    assert results[0][0] == 0 and results[1][0] == 1200000
    assert abs(results[0][3] - ref["lufs"]) < 1e-6 and results[0][3] == results[1][3]
    assert np.array_equal(out, ref["loud_i16"])
    assert mel.shape == ref["log_mel"].shape and np.abs(mel - ref["log_mel"]).max() <= 1e-4

"""TEST DOUBLE: the ``FrontEnd`` / ``Plan`` surface that ``GpuAudioPreprocessor`` drives, computed by the CPU oracle.

Lets the host side of the drop-in (path bookkeeping, plan / resident caches, background writer, error mapping, call order under
the reference's own ``AudioPipeline``) run in the CPU container where there is no CUDA device.  Lives under ``tests/`` because it
routes through ``oracle/`` -- the product never does (``modular_audio_pipeline_b200`` has no CPU path: ``FrontEnd()`` raises
without a GPU).  Tensors are CPU tensors; layouts follow include/maf.h (items 8-sample aligned).
"""
from __future__ import annotations

import types
from typing import List, Optional

import numpy as np
import torch

from modular_audio_pipeline_b200._capi import MAF_ERR_INVALID, MAF_ERR_TOO_SHORT, MafError
from oracle import reference_path as O


def _align8(n):
    return (int(n) + 7) // 8 * 8


class FakePlan:
    def __init__(self, fe, lengths, sample_rate=16000, stationary=True, prop_decrease=0.8, auto_detect_noise=True,
                 noise_sample_duration_s=0.5, external_noise=False, peak_normalize=True, target_lufs=-16.0, n_mels=80,
                 mel_padding=0, shards=None):
        self.fe = fe
        self.lengths = [int(n) for n in lengths]
        self.params = types.SimpleNamespace(sample_rate=sample_rate, stationary=int(stationary), prop_decrease=prop_decrease,
                                            auto_detect_noise=int(auto_detect_noise), noise_sample_duration_s=noise_sample_duration_s,
                                            external_noise=int(external_noise), peak_normalize=int(peak_normalize),
                                            target_lufs=target_lufs, n_mels=n_mels, mel_padding=mel_padding)
        self.n_items = len(self.lengths)
        self.layout, off = [], 0
        for n in self.lengths:
            self.layout.append(types.SimpleNamespace(pcm_off=off, out_off=off, mel_off=0, mel_frames=(n + mel_padding) // 160, n_chunks=1))
            off += _align8(n)
        self.total_pcm = self.total_out = off
        self.closed = False
        self.noise_len = None
        fe.plans_created += 1

    device = torch.device("cpu")

    def close(self):
        self.closed = True

    def out_buffer(self):
        return torch.zeros(max(self.total_out, 8), dtype=torch.int16)

    def pack(self, items, out=None):
        if out is None:
            out = torch.zeros(max(self.total_pcm, 8), dtype=torch.int16)
        for lay, n, x in zip(self.layout, self.lengths, items):
            assert x.dtype == torch.int16 and x.numel() == n
            out[lay.pcm_off: lay.pcm_off + n] = x
        return out

    def _items(self, buf) -> List[np.ndarray]:
        assert not self.closed, "plan used after close()"
        assert buf.dtype == torch.int16 and buf.numel() >= self.total_pcm
        return [buf[lay.pcm_off: lay.pcm_off + n].numpy() for lay, n in zip(self.layout, self.lengths)]

    def split_out(self, buf):
        return [buf[lay.out_off: lay.out_off + n] for lay, n in zip(self.layout, self.lengths)]

    def set_noise_clips(self, offsets, lengths):
        if any(n <= 768 for n in lengths):
            raise MafError(MAF_ERR_INVALID, "noise clip of <= 768 samples")
        self.noise_len = list(lengths)

    def denoise(self, pcm, noise_pcm=None, want_float=False, out=None):
        p = self.params
        out = self.out_buffer() if out is None else out
        clips = []
        for k, x in enumerate(self._items(pcm)):
            if p.external_noise:
                noise = noise_pcm[: self.noise_len[k]].numpy().astype(np.float32)
                red = O.reduce_noise(x.astype(np.float32), p.sample_rate, noise, p.prop_decrease, bool(p.stationary))
                clips.append((0, self.noise_len[k]))
            else:
                red, _, c = O.denoise_int16(x, p.sample_rate, bool(p.stationary), bool(p.auto_detect_noise), p.noise_sample_duration_s,
                                            p.prop_decrease)
                clips.append(c)
            o = self.layout[k].out_off
            out[o: o + len(x)] = torch.from_numpy(O.to_int16_trunc(red))
        return self.split_out(out), None, torch.tensor(clips, dtype=torch.int32)

    def peak_normalize(self, pcm, out=None):
        out = self.out_buffer() if out is None else out
        for k, x in enumerate(self._items(pcm)):
            o = self.layout[k].out_off
            out[o: o + len(x)] = torch.from_numpy(O.peak_normalize(x))
        return self.split_out(out)

    def loudness(self, pcm, out=None):
        p = self.params
        out = self.out_buffer() if out is None else out
        lufs = []
        for k, x in enumerate(self._items(pcm)):
            if len(x) < 0.4 * p.sample_rate:
                raise MafError(MAF_ERR_TOO_SHORT, "Audio must have length greater than the block size")
            y, l, _ = O.normalize_loudness(x, p.sample_rate, p.target_lufs)
            o = self.layout[k].out_off
            out[o: o + len(x)] = torch.from_numpy(y)
            lufs.append(l)
        return self.split_out(out), torch.tensor(lufs, dtype=torch.float64)

    def frame_stats(self, pcm):
        e, z = O.frame_stats(self._items(pcm)[0].astype(np.float32), self.params.sample_rate)
        return torch.from_numpy(e.astype(np.float32)), torch.from_numpy((z * int(self.params.sample_rate * 0.025)).round().astype(np.int32))


class FakeFrontEnd:
    device = torch.device("cpu")

    def __init__(self):
        self.plans_created = 0
        self.calls: List[str] = []

    def plan(self, lengths, **kw):
        return FakePlan(self, lengths, **kw)

    def front_door(self, pcm, channels, in_rate, out_rate, normalize=True):
        self.calls.append("front_door")
        return torch.from_numpy(O.front_door(pcm.numpy(), channels, in_rate, out_rate, normalize))

    def loudness_interleaved(self, pcm, channels, sample_rate, target_lufs=-16.0):
        self.calls.append("loudness_interleaved")
        y, l, _ = O.normalize_loudness(pcm.numpy(), sample_rate, target_lufs, channels=channels)
        return torch.from_numpy(np.ascontiguousarray(y[: pcm.numel() // channels])), torch.tensor([l], dtype=torch.float64)

    def silence(self, x, sr, window_ms, offset_db, margin_ms=0, compact=False):
        self.calls.append("silence")
        pcm = x.numpy()
        silent = np.array(O.detect_silence_ms(pcm, sr, window_ms, offset_db), dtype=np.int32).reshape(-1, 2)
        if not compact:
            return 0, silent, np.zeros((0, 4), dtype=np.int32), None
        try:
            res = O.remove_silence(pcm, sr, window_ms, offset_db, margin_ms)
        except ValueError as e:
            return (3 if "TooManyMissingFrames" in str(e) else 2), silent, np.zeros((0, 4), dtype=np.int32), None
        if res is None:
            return 1, silent, np.zeros((0, 4), dtype=np.int32), None
        cleaned, maps = res
        table = np.array([[round(a * 1000), round((b - a) * 1000), round(c * 1000), round(d * 1000)] for a, b, c, d in maps],
                         dtype=np.int32).reshape(-1, 4)
        return 0, silent, table, torch.from_numpy(cleaned.copy())

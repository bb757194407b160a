"""remove_silence / detect_silence_segments (reference ``audio_pipeline/preprocessor.py:290-405``).

The O(N x window) part -- pydub's 1 ms sliding ``audioop.rms`` scan -- runs on the GPU with exact integer sums
(csrc/silence.cu).  What is left is bookkeeping on per-millisecond flags (range merging of
``pydub.silence.detect_silence`` / ``detect_nonsilent``), the margin / cross-fade arithmetic of the reference's loop
(:330-358) and the compaction of the kept ranges, done with device tensor slices.  pydub is not installable here, so
its cross-fade (``AudioSegment.append`` -> per-frame ``audioop.mul`` fades + saturating ``audioop.add``) is restated
from its published source: parity for this next-row component is unpinned.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Tuple

import numpy as np
import torch

from ._capi import check


def _ms_len(n: int, sr: int) -> int:
    return int(round(1000.0 * n / sr))                        # pydub len(seg)


def _flags(fe, x: torch.Tensor, sr: int, window_ms: int, offset_db: float):
    """(per-start silent flags as numpy bool, length in ms) -- or (None, ms) when the clip is shorter than the window."""
    n = x.numel()
    n_ms = _ms_len(n, sr)
    if n_ms < window_ms:
        return None, n_ms
    dev = fe.device
    bucket = torch.zeros(n_ms + window_ms + 1, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        check(fe._lib.maf_ms_energy(fe._h, x.data_ptr(), n, sr, n_ms + 1, bucket.data_ptr(), total.data_ptr(), st))
    rms = int(math.sqrt(total.item() / n))                    # audioop.rms truncates
    dbfs = 20 * math.log10(rms / 32768.0) if rms > 0 else -float("inf")   # seg.dBFS
    thresh = (10 ** ((dbfs - offset_db) / 20)) * 32768.0      # db_to_float(silence_thresh) * max_possible_amplitude
    n_starts = n_ms - window_ms + 1
    flags = torch.empty(n_starts, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(fe._lib.maf_silence_flags(fe._h, bucket.data_ptr(), n, sr, n_starts, window_ms, C.c_double(thresh), flags.data_ptr(), st))
    return flags.cpu().numpy().astype(bool), n_ms


def detect_silence_ms(fe, x: torch.Tensor, sr: int, min_silence_len: int, offset_db: float) -> List[List[int]]:
    """pydub.silence.detect_silence(seg, min_silence_len, seg.dBFS - offset_db, seek_step=1) -> [[start_ms, end_ms]]."""
    flags, n_ms = _flags(fe, x, sr, min_silence_len, offset_db)
    if flags is None or not flags.any():
        return []
    starts = np.nonzero(flags)[0]
    # a new range begins where the previous silent start is neither adjacent nor within min_silence_len
    gap = np.diff(starts)
    brk = np.nonzero((gap != 1) & (gap > min_silence_len))[0]
    firsts = np.concatenate(([starts[0]], starts[brk + 1]))
    lasts = np.concatenate((starts[brk], [starts[-1]]))
    return [[int(a), int(b) + min_silence_len] for a, b in zip(firsts, lasts)]


def detect_nonsilent_ms(fe, x, sr, min_silence_len, offset_db) -> List[List[int]]:
    n_ms = _ms_len(x.numel(), sr)
    silent = detect_silence_ms(fe, x, sr, min_silence_len, offset_db)
    if not silent:
        return [[0, n_ms]]
    if silent[0][0] == 0 and silent[0][1] == n_ms:
        return []
    out, prev_end, end_i = [], 0, 0
    for start_i, end_i in silent:
        out.append([prev_end, start_i])
        prev_end = end_i
    if end_i != n_ms:
        out.append([prev_end, n_ms])
    if out[0] == [0, 0]:
        out.pop(0)
    return out


def _slice(x: torch.Tensor, sr: int, a_ms: int, b_ms: int) -> torch.Tensor:
    r = sr / 1000.0
    return x[int(a_ms * r): int(b_ms * r)]                     # AudioSegment.__getitem__ -> frame_count(ms)


def _crossfade(a_tail: torch.Tensor, b_head: torch.Tensor) -> torch.Tensor:
    """seg1[-cf:].fade(to_gain=-120) overlaid with seg2[:cf].fade(from_gain=-120): per-frame audioop.mul (floor) fades,
    saturating audioop.add."""
    n = a_tail.numel()
    i = torch.arange(n, device=a_tail.device, dtype=torch.float64)
    lo = 10 ** (-120 / 20)
    out_gain = 1.0 + ((lo - 1.0) / n) * i
    in_gain = lo + ((1.0 - lo) / n) * i
    fa = torch.floor(torch.clamp(a_tail.to(torch.float64) * out_gain, -32768, 32767))
    fb = torch.floor(torch.clamp(b_head.to(torch.float64) * in_gain, -32768, 32767))
    return torch.clamp(fa + fb, -32768, 32767).to(torch.int16)


def remove_silence(fe, x: torch.Tensor, sr: int, min_silence_len: int, offset_db: float, margin: int, preserve: bool,
                   mapping_cls) -> Optional[Tuple[torch.Tensor, list]]:
    """The loop of preprocessor.py:330-358 on device-resident samples.  None = "no non-silent segments"."""
    ranges = detect_nonsilent_ms(fe, x, sr, min_silence_len, offset_db)
    if not ranges:
        return None
    n_ms = _ms_len(x.numel(), sr)
    pieces: List[torch.Tensor] = []
    mappings = []
    pos_ms = 0
    have = False
    for start_ms, end_ms in ranges:
        s = max(0, start_ms - margin)
        e = min(n_ms, end_ms + margin)
        chunk = _slice(x, sr, s, e)
        dur = _ms_len(chunk.numel(), sr)
        if preserve:
            mappings.append(mapping_cls(processed_start=pos_ms / 1000.0, processed_end=(pos_ms + dur) / 1000.0,
                                        original_start=s / 1000.0, original_end=e / 1000.0))
        if not have:
            pieces = [chunk]
            have = chunk.numel() > 0
        else:
            cf = min(20, dur // 4)
            ncf = int(cf * (sr / 1000.0))
            if ncf > 0 and pieces[-1].numel() >= ncf and chunk.numel() >= ncf:
                tail = pieces[-1][-ncf:]
                pieces[-1] = pieces[-1][:-ncf]
                pieces.append(_crossfade(tail, chunk[:ncf]))
                pieces.append(chunk[ncf:])
            else:
                pieces.append(chunk)
            pos_ms -= cf
        pos_ms += dur
    return torch.cat(pieces), mappings

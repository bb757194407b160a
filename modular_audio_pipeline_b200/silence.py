"""remove_silence / detect_silence_segments (reference ``audio_pipeline/preprocessor.py:290-405``).

Everything numerical runs on the GPU (csrc/silence.cu through ``FrontEnd.silence``): pydub's 1 ms sliding ``audioop.rms`` scan as
prefix sums of exact integer energies, its merge of silent starts into ranges, ``detect_nonsilent``'s complement, the reference's
margin / ``min(20, dur // 4)`` ms cross-fade / running-position loop (:330-358) and the compaction of the kept chunks with
pydub's cross-fade (per-frame ``audioop.mul`` fades, saturating ``audioop.add``).  This file only turns the small result
tables into the reference's return types.  pydub is not installable here, so ``AudioSegment.__getitem__`` / ``append`` /
``fade`` are restated from its published source (0.25.1): parity for this next-row component is unpinned beyond the stdlib
``audioop`` calls the oracle makes.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

_PYDUB_ERRORS = {2: "Crossfade is longer than the original AudioSegment",
                 3: "TooManyMissingFrames: you should never be filling in more than 2 ms with silence here"}


def detect_silence_ms(fe, x: torch.Tensor, sr: int, min_silence_len: int, offset_db: float) -> List[List[int]]:
    """pydub.silence.detect_silence(seg, min_silence_len, seg.dBFS - offset_db, seek_step=1) -> [[start_ms, end_ms]]."""
    _status, silent, _table, _ = fe.silence(x, sr, min_silence_len, offset_db, 0, compact=False)
    return [[int(a), int(b)] for a, b in silent]


def remove_silence(fe, x: torch.Tensor, sr: int, min_silence_len: int, offset_db: float, margin: int, preserve: bool,
                   mapping_cls) -> Optional[Tuple[torch.Tensor, list]]:
    """The loop of preprocessor.py:311-358 on device-resident samples.  None = "no non-silent segments" (:321-323)."""
    status, _silent, table, cleaned = fe.silence(x, sr, min_silence_len, offset_db, margin, compact=True)
    if status == 1:
        return None
    if status != 0:                                             # what pydub raises; the caller wraps it like :374-375
        raise ValueError(_PYDUB_ERRORS.get(status, f"silence removal failed (status {status})"))
    mappings = []
    if preserve:
        for pos, dur, s, e in table.tolist():
            mappings.append(mapping_cls(processed_start=pos / 1000.0, processed_end=(pos + dur) / 1000.0,
                                        original_start=s / 1000.0, original_end=e / 1000.0))
    return cleaned, mappings

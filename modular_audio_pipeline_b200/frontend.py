"""Device-resident front-end: batch plans over libmaf_b200.so with torch tensors as the memory hand-off.

This is the in-memory "fast path beyond the Protocol" of SURVEY.md section 8(b): ``FrontEnd.plan(...)`` fixes the
geometry of a batch (lengths, mode, mel size), ``Plan.process(pcm)`` runs denoise -> peak normalise -> loudness
normalise -> log-mel entirely on the GPU and returns device tensors (the log-mel is what Whisper's encoder eats).
PyTorch is used for device memory, streams and the tensor hand-off only; all arithmetic is in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _capi
from ._capi import ItemDesc, Layout, MafError, Params, check

CHUNK = 600000
PAD = 30000
KW_HALO = 2048


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("modular_audio_pipeline_b200 needs a CUDA device (B200); there is no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("modular_audio_pipeline_b200 runs on CUDA devices only")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


@dataclass
class FrontEndResult:
    pcm: List[torch.Tensor]            # final int16 PCM per item (views into one device buffer)
    log_mel: List[torch.Tensor]        # float32 [n_mels, frames] per item (views); empty when n_mels == 0
    lufs: torch.Tensor                 # float64 [n_items], loudness measured before the gain
    noise_clip: Optional[torch.Tensor]  # int32 [n_items, 2] sample range of the noise clip (stationary mode)


class FrontEnd:
    """One per process and GPU: owns the constant tables (windows, twiddles, mel banks) on the device."""

    def __init__(self, device=None):
        self.device = _require_cuda(device)
        self._lib = _capi.lib()
        h = C.c_void_p()
        check(self._lib.maf_create(self.device.index, C.byref(h)))
        self._h = h
        self._plans = weakref.WeakSet()

    def front_door(self, pcm: torch.Tensor, channels: int, in_rate: int, out_rate: int, normalize: bool = True) -> torch.Tensor:
        """normalize_audio's conversions for 16-bit input that is not yet mono at the target rate (preprocessor.py:227-233:
        pydub normalize -> set_frame_rate -> set_channels(1), i.e. audioop.max / mul / ratecv / tomono): interleaved int16
        CUDA tensor in, mono int16 CUDA tensor at ``out_rate`` out."""
        if pcm.dtype != torch.int16 or not pcm.is_cuda or not pcm.is_contiguous():
            raise MafError(_capi.MAF_ERR_INVALID, "front_door: contiguous int16 CUDA tensor expected")
        if channels not in (1, 2) or pcm.numel() % channels:
            raise MafError(_capi.MAF_ERR_UNSUPPORTED, "front_door: 1 or 2 interleaved channels expected")
        n_in = pcm.numel() // channels
        n_out = int(self._lib.maf_front_door_out_len(n_in, int(in_rate), int(out_rate)))
        out = torch.empty(max(n_out, 1), dtype=torch.int16, device=self.device)
        if n_in > 0:
            with torch.cuda.device(self.device):
                rc = self._lib.maf_front_door(self._h, pcm.data_ptr(), n_in, int(channels), int(in_rate), int(out_rate), int(bool(normalize)),
                                              out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
            check(rc)
        return out[:n_out]

    def mel_windows(self, mel: torch.Tensor, seeks: torch.Tensor, lens: Optional[torch.Tensor], window_frames: int,
                    out: torch.Tensor) -> torch.Tensor:
        """Encoder windows ``out[w, m, f] = mel[m, seeks[w] + f]`` (zero past ``lens[w]`` / the end of the recording), float32 or
        float16 -- maf_mel_windows in include/maf.h; the whisper seek loop behind transcriber.py:229-238."""
        n = int(seeks.numel())
        if mel.dtype != torch.float32 or mel.dim() != 2 or not mel.is_cuda or not mel.is_contiguous():
            raise MafError(_capi.MAF_ERR_INVALID, "mel_windows: contiguous float32 [n_mels, frames] CUDA tensor expected")
        if seeks.dtype != torch.int32 or not seeks.is_cuda or (lens is not None and (lens.dtype != torch.int32 or not lens.is_cuda)):
            raise MafError(_capi.MAF_ERR_INVALID, "mel_windows: seeks / lens must be int32 CUDA tensors")
        if lens is not None and int(lens.numel()) != n:
            raise MafError(_capi.MAF_ERR_INVALID, "mel_windows: seeks and lens differ in length")
        if out.dtype not in (torch.float16, torch.float32) or not out.is_cuda or out.numel() != n * int(mel.shape[0]) * int(window_frames):
            raise MafError(_capi.MAF_ERR_INVALID, "mel_windows: out must be a float16 / float32 CUDA tensor [n_windows, n_mels, window_frames]")
        with torch.cuda.device(self.device):
            rc = self._lib.maf_mel_windows(self._h, mel.data_ptr(), int(mel.shape[0]), int(mel.shape[1]), seeks.data_ptr(),
                                           lens.data_ptr() if lens is not None else None, n, int(window_frames),
                                           int(out.dtype == torch.float16), out.data_ptr(),
                                           torch.cuda.current_stream(self.device).cuda_stream)
        check(rc)
        return out

    def loudness_interleaved(self, pcm: torch.Tensor, channels: int, sample_rate: int, target_lufs: float = -16.0):
        """normalize_loudness on multi-channel 16-bit input (preprocessor.py:256-282): float32 channel mean -> BS.1770 loudness ->
        gain / limiter / int16 of the MONO signal.  Returns ``(mono int16 CUDA tensor, float64 LUFS CUDA tensor [1])``; the
        output is meaningless when the LUFS is non-finite or below -70 (the reference returns its input path then)."""
        if pcm.dtype != torch.int16 or not pcm.is_cuda or not pcm.is_contiguous() or channels < 1 or pcm.numel() % channels:
            raise MafError(_capi.MAF_ERR_INVALID, "loudness_interleaved: contiguous int16 CUDA tensor of whole frames expected")
        n = pcm.numel() // channels
        dev = self.device
        ws = torch.empty(max(256, int(self._lib.maf_loudness_interleaved_ws(n, int(sample_rate)))), dtype=torch.uint8, device=dev)
        out = torch.empty(max(n, 1), dtype=torch.int16, device=dev)
        lufs = torch.empty(1, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(self._lib.maf_loudness_interleaved(self._h, pcm.data_ptr(), n, int(channels), int(sample_rate), float(target_lufs),
                                                     ws.data_ptr(), out.data_ptr(), lufs.data_ptr(),
                                                     torch.cuda.current_stream(dev).cuda_stream))
        return out[:n], lufs

    def silence(self, x: torch.Tensor, sr: int, window_ms: int, offset_db: float, margin_ms: int = 0, compact: bool = False):
        """pydub's silence scan and (``compact``) the reference's remove_silence loop (preprocessor.py:311-358) on device-resident
        int16 samples -- maf_ms_energy + maf_silence in include/maf.h.  Returns ``(status, silent_ranges_ms [K, 2] int32 numpy,
        chunk_table [C, 4] int32 numpy (processed_start_ms, duration_ms, original_start_ms, original_end_ms), cleaned int16 CUDA
        tensor or None)``.  Two scalar-sized device->host reads: the recording's sum of squares (the threshold is the reference's
        own float64 expression, evaluated on the host) and the result block; no per-millisecond data leaves the device."""
        import math
        import numpy as np
        if x.dtype != torch.int16 or not x.is_cuda or not x.is_contiguous() or x.dim() != 1:
            raise MafError(_capi.MAF_ERR_INVALID, "silence: contiguous 1-D int16 CUDA tensor expected")
        n = int(x.numel())
        if n == 0:
            raise MafError(_capi.MAF_ERR_INVALID, "silence: empty recording")
        n_ms = int(round(1000 * (float(n) / sr)))                  # pydub len(seg)
        dev = self.device
        st = torch.cuda.current_stream(dev).cuda_stream
        bucket = torch.empty(n_ms + 1, dtype=torch.int64, device=dev)
        total = torch.empty(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            check(self._lib.maf_ms_energy(self._h, x.data_ptr(), n, int(sr), n_ms, bucket.data_ptr(), total.data_ptr(), st))
        rms = int(math.sqrt(total.item() / n))                    # audioop.rms truncates
        dbfs = 20 * math.log10(rms / 32768.0) if rms > 0 else -float("inf")   # seg.dBFS
        thresh = (10 ** ((dbfs - offset_db) / 20)) * 32768.0      # db_to_float(silence_thresh) * seg.max_possible_amplitude
        ws_b, res_b, cap, off_s, off_m = (C.c_int64() for _ in range(5))
        check(self._lib.maf_silence_layout(n_ms, int(window_ms), C.byref(ws_b), C.byref(res_b), C.byref(cap), C.byref(off_s), C.byref(off_m)))
        ws = torch.empty(ws_b.value, dtype=torch.uint8, device=dev)
        out = None
        out_cap = 0
        if compact:
            # chunks overlap only when the margins exceed the silent gap between them; + the sub-millisecond padding per chunk
            per_chunk = max(0, 2 * int(margin_ms) - int(window_ms)) + 3
            out_cap = n + cap.value * per_chunk * int(math.ceil(sr / 1000.0))
            out = torch.empty(out_cap, dtype=torch.int16, device=dev)
        with torch.cuda.device(dev):
            check(self._lib.maf_silence(self._h, x.data_ptr(), n, int(sr), bucket.data_ptr(), n_ms, int(window_ms), C.c_double(thresh),
                                        int(margin_ms), ws.data_ptr(), None if out is None else out.data_ptr(), out_cap, st))
        res = ws[: res_b.value].cpu().numpy()                      # header + range / mapping tables (a few KB)
        status, n_silent, n_chunks, _j_base = (int(v) for v in res[:16].view(np.int32))
        n_out = int(res[16:24].view(np.int64)[0])
        silent = res[off_s.value: off_s.value + 8 * n_silent].view(np.int32).reshape(-1, 2).copy()
        table = res[off_m.value: off_m.value + 16 * n_chunks].view(np.int32).reshape(-1, 4).copy()
        return status, silent, table, (out[:n_out] if out is not None and status == 0 else None)

    def close(self) -> None:
        if getattr(self, "_h", None):
            for p in list(self._plans):          # plans point into the context: release them first
                p.close()
            self._lib.maf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plan(self, lengths: Sequence[int], sample_rate: int = 16000, stationary: bool = True,
             prop_decrease: float = 0.8, auto_detect_noise: bool = True, noise_sample_duration_s: float = 0.5,
             external_noise: bool = False, peak_normalize: bool = True, target_lufs: float = -16.0,
             n_mels: int = 80, mel_padding: int = 0, shards: Optional[Sequence[ItemDesc]] = None) -> "Plan":
        params = Params(sample_rate, int(stationary), prop_decrease, int(auto_detect_noise), noise_sample_duration_s,
                        int(external_noise), int(peak_normalize), target_lufs, n_mels, mel_padding)
        if shards is None:
            descs = (ItemDesc * len(lengths))(*[ItemDesc(int(n), 0, int(n), 0, int(n), 0, 0) for n in lengths])
        else:
            descs = (ItemDesc * len(shards))(*shards)
        return Plan(self, params, descs)


class Plan:
    """Geometry + workspace of one batch.  Reusable: call process() as often as you like with new PCM."""

    def __init__(self, fe: FrontEnd, params: Params, descs):
        self.fe = fe
        self.params = params
        self.n_items = len(descs)
        self._lib = fe._lib
        h = C.c_void_p()
        check(self._lib.maf_plan_create(fe._h, C.byref(params), descs, self.n_items, C.byref(h)))
        self._h = h
        fe._plans.add(self)
        self.descs = descs
        self.layout: List[Layout] = []
        for i in range(self.n_items):
            lay = Layout()
            check(self._lib.maf_plan_layout(h, i, C.byref(lay)))
            self.layout.append(lay)
        self.total_pcm = int(self._lib.maf_plan_total_pcm(h))
        self.total_out = int(self._lib.maf_plan_total_out(h))
        self.total_mel = int(self._lib.maf_plan_total_mel(h))
        self.workspace_bytes = int(self._lib.maf_plan_workspace_bytes(h))
        self._ws: Optional[torch.Tensor] = None

    # ---- memory ------------------------------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.fe.device

    @property
    def workspace(self) -> torch.Tensor:
        if self._ws is None:
            self._ws = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.maf_plan_destroy(self._h)
            self._h = None
        self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self._lib.maf_plan_launches(self._h))

    STAGES = ("noise", "stft", "ns_mask", "smooth_mask", "istft", "kweight", "gate", "logmel", "mel_finalize")

    def profile(self, max_steps: int) -> None:
        """Record CUDA events between the stages of the next process() calls (ring of ``max_steps`` sets)."""
        check(self._lib.maf_plan_profile(self._h, max_steps))

    def stage_ms(self, step: int) -> dict:
        out = {}
        ms = C.c_float()
        for i, name in enumerate(self.STAGES, start=1):
            check(self._lib.maf_plan_stage_ms(self._h, step, i, C.byref(ms)))
            out[name] = ms.value
        return out

    def buffer(self, name: str, dtype: torch.dtype) -> torch.Tensor:
        """View of a named workspace buffer (for the collectives of the time-sharded mode)."""
        off, size = C.c_int64(), C.c_int64()
        check(self._lib.maf_plan_buffer(self._h, name.encode(), C.byref(off), C.byref(size)))
        return self.workspace[off.value: off.value + size.value].view(dtype)

    def workspace_layout(self):
        """[(offset, size)] of every buffer carved out of the workspace (``maf_plan_ws_layout``); the bounds-check test reads the
        gaps between them (MAF_DEBUG_REDZONE)."""
        n = int(self._lib.maf_plan_ws_layout(self._h, None, None, 0))
        offs, sizes = (C.c_int64 * n)(), (C.c_int64 * n)()
        self._lib.maf_plan_ws_layout(self._h, offs, sizes, n)
        return [(int(offs[i]), int(sizes[i])) for i in range(n)]

    def den_offset(self, item: int = 0) -> int:
        return int(self._lib.maf_plan_den_offset(self._h, item))

    def pack(self, items: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Lay int16 recordings out in the plan's input layout (device tensor unless ``out`` is given)."""
        if out is None:
            out = torch.zeros(max(self.total_pcm, 8), dtype=torch.int16, device=self.device)
        for lay, d, x in zip(self.layout, self.descs, items):
            if x.dtype != torch.int16 or x.numel() != d.buf_len:
                raise ValueError("pack: every item must be int16 with the planned length")
            out[lay.pcm_off: lay.pcm_off + d.buf_len].copy_(x, non_blocking=True)
        return out

    def split_out(self, buf: torch.Tensor) -> List[torch.Tensor]:
        return [buf[lay.out_off: lay.out_off + d.own_len] for lay, d in zip(self.layout, self.descs)]

    def split_mel(self, buf: torch.Tensor) -> List[torch.Tensor]:
        nm = self.params.n_mels
        return [buf[lay.mel_off: lay.mel_off + nm * lay.mel_frames].view(nm, lay.mel_frames) for lay in self.layout]

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_noise_clips(self, offsets: Sequence[int], lengths: Sequence[int]) -> None:
        n = self.n_items
        check(self._lib.maf_set_noise_clips(self._h, (C.c_int64 * n)(*offsets), (C.c_int32 * n)(*lengths)))

    # ---- fused hot path ----------------------------------------------------------------------------------------
    def process(self, pcm: torch.Tensor, noise_pcm: Optional[torch.Tensor] = None,
                out_pcm: Optional[torch.Tensor] = None, mel: Optional[torch.Tensor] = None) -> FrontEndResult:
        """denoise -> peak normalise -> loudness normalise -> log-mel on device-resident int16 PCM (plan layout)."""
        self._check_in(pcm)
        dev = self.device
        if out_pcm is None:
            out_pcm = torch.empty(max(self.total_out, 8), dtype=torch.int16, device=dev)
        nm = self.params.n_mels
        if nm and mel is None:
            mel = torch.empty(max(self.total_mel, 4), dtype=torch.float32, device=dev)
        lufs = torch.empty(self.n_items, dtype=torch.float64, device=dev)
        clip = torch.zeros(self.n_items, 2, dtype=torch.int32, device=dev) if self.params.stationary else None
        with torch.cuda.device(dev):
            check(self._lib.maf_process(self._h, _ptr(pcm), _ptr(noise_pcm), _ptr(self.workspace), _ptr(out_pcm),
                                        _ptr(mel) if nm else None, _ptr(lufs), _ptr(clip), self._stream()))
        return FrontEndResult(self.split_out(out_pcm), self.split_mel(mel) if nm else [], lufs, clip)

    def process_host(self, pcm: torch.Tensor, out_pcm: torch.Tensor, mel: Optional[torch.Tensor],
                     lufs: torch.Tensor, clip: Optional[torch.Tensor] = None,
                     noise_pcm: Optional[torch.Tensor] = None) -> None:
        """Same call on HOST tensors (pinned for full PCIe speed); returns when the results are in host memory."""
        for t in (pcm, out_pcm, mel, lufs, clip, noise_pcm):
            if t is not None and t.device.type != "cpu":
                raise ValueError("process_host takes CPU tensors")
        with torch.cuda.device(self.device):
            check(self._lib.maf_process_host(self._h, _ptr(pcm), _ptr(noise_pcm),
                                             0 if noise_pcm is None else noise_pcm.numel(), _ptr(out_pcm), _ptr(mel),
                                             _ptr(lufs), _ptr(clip)))

    # ---- single stages (what the Protocol methods call) --------------------------------------------------------------
    def _check_in(self, pcm: torch.Tensor) -> None:
        if pcm.dtype != torch.int16 or pcm.device != self.device or pcm.numel() < self.total_pcm or not pcm.is_contiguous():
            raise ValueError("expected a contiguous int16 tensor in the plan's layout on the plan's device")

    def out_buffer(self) -> torch.Tensor:
        """A fresh int16 device buffer in the plan's OUTPUT layout.  For a single-item plan the output layout equals the input
        layout, so a stage's output buffer can be handed to the next stage as is (no re-pack)."""
        return torch.empty(max(self.total_out, 8), dtype=torch.int16, device=self.device)

    def denoise(self, pcm: torch.Tensor, noise_pcm: Optional[torch.Tensor] = None, want_float: bool = False,
                out: Optional[torch.Tensor] = None):
        self._check_in(pcm)
        dev = self.device
        if out is None:
            out = self.out_buffer()
        outf = torch.empty(max(self.total_out, 8), dtype=torch.float32, device=dev) if want_float else None
        clip = torch.zeros(self.n_items, 2, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(self._lib.maf_denoise(self._h, _ptr(pcm), _ptr(noise_pcm), _ptr(self.workspace), _ptr(out), _ptr(outf),
                                        _ptr(clip), self._stream()))
        return self.split_out(out), (self.split_out(outf) if want_float else None), clip

    def peak_normalize(self, pcm: torch.Tensor, out: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        self._check_in(pcm)
        if out is None:
            out = self.out_buffer()
        with torch.cuda.device(self.device):
            check(self._lib.maf_peak_normalize(self._h, _ptr(pcm), _ptr(self.workspace), _ptr(out), self._stream()))
        return self.split_out(out)

    def loudness(self, pcm: torch.Tensor, out: Optional[torch.Tensor] = None):
        self._check_in(pcm)
        if out is None:
            out = self.out_buffer()
        lufs = torch.empty(self.n_items, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.maf_loudness(self._h, _ptr(pcm), _ptr(self.workspace), _ptr(out), _ptr(lufs), self._stream()))
        return self.split_out(out), lufs

    def frame_stats(self, pcm: torch.Tensor):
        """S0 per-frame statistics of every item: (float32 energies, int32 zero-crossing counts), concatenated."""
        self._check_in(pcm)
        hop = int(self.params.sample_rate * 0.010)
        flen = int(self.params.sample_rate * 0.025)
        total = 0
        for d in self.descs:
            fg = (d.n_global - flen) // hop + 1 if d.n_global >= flen else 0
            first = -(-d.g0 // hop)
            last = min(fg, -(-(d.g0 + d.own_len) // hop))
            total += max(0, last - first)
        e = torch.empty(max(total, 1), dtype=torch.float32, device=self.device)
        z = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.maf_frame_stats(self._h, _ptr(pcm), _ptr(e), _ptr(z), self._stream()))
        return e[:total], z[:total]

    def log_mel(self, pcm: torch.Tensor) -> List[torch.Tensor]:
        self._check_in(pcm)
        mel = torch.empty(max(self.total_mel, 4), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.maf_logmel(self._h, _ptr(pcm), _ptr(self.workspace), _ptr(mel), self._stream()))
        return self.split_mel(mel)


def capture_graph(fn, device=None):
    """Capture the asynchronous work ``fn()`` issues (copies, front door, ``Plan.process`` ...) into ONE CUDA graph and
    return ``(graph, fn's return value)``; ``graph.replay()`` re-issues the whole launch sequence with a single call.  Meant
    for launch-bound small batches (BASELINE config 5: one 30 s chunk is ~20 kernels of a few microseconds each).  Every
    tensor ``fn`` touches must stay alive and keep its address; tensors ``fn`` allocates live in the graph's memory pool."""
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        fn()                                               # warm-up outside capture (lazy module load, function attributes)
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize(device)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ret = fn()
    return graph, ret


class HostPipeline:
    """Host-buffer entry with copy / compute overlap: a three-stage pipeline -- upload stream, kernel stream, download stream,
    chained by events -- over ``depth`` slots (plan + device / pinned buffers each), so the H2D copy of batch i+1 and the D2H
    copy of batch i-1 run while the kernels of batch i do.  All kernels are issued to ONE stream in submission order: steps
    never time-slice the SMs against each other, each stage is busy with exactly one batch.

    ``submit(slot_inputs)`` is asynchronous; ``results(slot)`` waits for that slot and returns pinned host tensors
    (final int16 PCM, float32 log-mel, float64 LUFS) in the plan's layouts.
    """

    def __init__(self, fe: FrontEnd, lengths: Sequence[int], depth: int = 2, mel_to_host: bool = True, encoder_windows: bool = False,
                 **plan_kw):
        self.fe = fe
        self.depth = depth
        # encoder_windows=True additionally cuts every item's log-mel into the 30 s float16 windows a Whisper encoder on the same
        # GPU consumes (maf_mel_windows; whisper's seek loop at full-window stride): ``d_win[slot][item]`` = [windows, n_mels, 3000]
        self.encoder_windows = encoder_windows
        # mel_to_host=False keeps the log-mel on the device (``d_mel[slot]``) for the hand-off into a Whisper encoder on the
        # same GPU (transcriber.py adapters); only the final PCM and the LUFS travel back over PCIe
        self.mel_to_host = mel_to_host
        self.plans = [fe.plan(lengths, **plan_kw) for _ in range(depth)]
        p0 = self.plans[0]
        dev = fe.device
        self.s_in, self.s_k, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_k = [torch.cuda.Event() for _ in range(depth)]
        self.d_in = [torch.empty(max(p0.total_pcm, 8), dtype=torch.int16, device=dev) for _ in range(depth)]
        self.d_out = [torch.empty(max(p0.total_out, 8), dtype=torch.int16, device=dev) for _ in range(depth)]
        self.d_mel = [torch.empty(max(p0.total_mel, 4), dtype=torch.float32, device=dev) for _ in range(depth)]
        self.h_in = [torch.empty(max(p0.total_pcm, 8), dtype=torch.int16).pin_memory() for _ in range(depth)]
        self.h_out = [torch.empty(max(p0.total_out, 8), dtype=torch.int16).pin_memory() for _ in range(depth)]
        self.h_mel = [torch.empty(max(p0.total_mel, 4), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.h_lufs = [torch.empty(p0.n_items, dtype=torch.float64).pin_memory() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.d_win: List[List[torch.Tensor]] = [[] for _ in range(depth)]
        self._seeks: List[torch.Tensor] = []
        if encoder_windows and p0.params.n_mels:
            for lay in p0.layout:
                self._seeks.append(torch.arange(0, max(lay.mel_frames, 1), 3000, dtype=torch.int32, device=dev))
            for s in range(depth):
                self.d_win[s] = [torch.empty(int(sk.numel()), p0.params.n_mels, 3000, dtype=torch.float16, device=dev) for sk in self._seeks]
        self._n = 0
        for p in self.plans:
            _ = p.workspace

    @property
    def h2d_bytes(self) -> int:
        return 2 * self.plans[0].total_pcm

    @property
    def d2h_bytes(self) -> int:
        p = self.plans[0]
        return 2 * p.total_out + (4 * p.total_mel if self.mel_to_host else 0) + 8 * p.n_items

    def submit(self) -> int:
        """Process the batch currently in ``h_in[slot]`` (slot = submission count mod depth).  Returns the slot."""
        slot = self._n % self.depth
        self._n += 1
        plan = self.plans[slot]
        self.done[slot].synchronize()                      # the slot's previous results must have been consumed
        with torch.cuda.stream(self.s_in):
            self.d_in[slot].copy_(self.h_in[slot], non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        self.s_k.wait_event(self.ev_in[slot])
        with torch.cuda.stream(self.s_k):
            res = plan.process(self.d_in[slot], out_pcm=self.d_out[slot], mel=self.d_mel[slot] if plan.params.n_mels else None)
            if self._seeks:
                for m, sk, w in zip(res.log_mel, self._seeks, self.d_win[slot]):
                    self.fe.mel_windows(m, sk, None, 3000, w)
            self.ev_k[slot].record(self.s_k)
        res.lufs.record_stream(self.s_out)
        self.s_out.wait_event(self.ev_k[slot])
        with torch.cuda.stream(self.s_out):
            self.h_out[slot].copy_(self.d_out[slot], non_blocking=True)
            if plan.params.n_mels and self.mel_to_host:
                self.h_mel[slot].copy_(self.d_mel[slot], non_blocking=True)
            self.h_lufs[slot].copy_(res.lufs, non_blocking=True)
            self.done[slot].record(self.s_out)
        return slot

    def next_slot(self) -> int:
        return self._n % self.depth

    def results(self, slot: int):
        self.done[slot].synchronize()
        p = self.plans[slot]
        mel_src = self.h_mel[slot] if self.mel_to_host else self.d_mel[slot]       # device tensors when the mel is handed over on the GPU
        return p.split_out(self.h_out[slot]), (p.split_mel(mel_src) if p.params.n_mels else []), self.h_lufs[slot]

    def drain(self) -> None:
        for e in self.done:
            e.synchronize()


__all__ = ["FrontEnd", "Plan", "FrontEndResult", "HostPipeline", "MafError", "ItemDesc", "capture_graph"]

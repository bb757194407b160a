"""GpuAudioPreprocessor -- drop-in for ``audio_pipeline.preprocessor.AudioPreprocessor`` (reference
``audio_pipeline/preprocessor.py:35-405``) behind ``PreprocessorProtocol`` (``protocols.py:68-90``).

Same constructor, same method names, argument meaning, output file names and error behaviour as the reference
class, so ``AudioPipeline(config, preprocessor=GpuAudioPreprocessor.from_config(config))`` runs unchanged.
Every numerical stage is a CUDA kernel of libmaf_b200.so; this file only does WAV I/O and path bookkeeping.

Beyond the Protocol ("fast path"):

* every stage output stays resident on the GPU keyed by its output path (a small LRU), so the next Protocol call on that
  path skips the disk read and the PCIe upload -- the orchestrator's chain (pipeline.py:331-357) uploads a recording once;
* batch plans are cached by geometry / mode, so a call costs its kernels, not a plan build;
* WAV files are written by a background writer from page-locked landing buffers.  ``write_mode="sync"`` (default) keeps the
  reference's contract -- every method returns when its output file is complete, which a vocal separator between the denoise
  and the normalise calls relies on (it re-opens ``_denoised.wav``; pipeline.py:338-342); ``write_mode="deferred"`` lets
  ``reduce_stationary_noise`` / ``normalize_audio`` / ``normalize_loudness`` return as soon as their write is queued and
  completes every file when ``remove_silence`` -- the orchestrator's last Protocol call, after which VAD / transcriber open
  the files -- returns, or in ``flush()``: for pipelines whose separator is the no-op one;
* ``prefetch()`` reads and uploads the next files while the current one is processed; ``process()`` / ``process_files()`` are
  the fused in-memory / file-batch entries (denoise -> peak normalise -> loudness -> log-mel in one launch sequence).
"""
from __future__ import annotations

import contextlib
import logging
import os
import wave
from collections import OrderedDict
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from . import wavio
from .frontend import FrontEnd, MafError

logger = logging.getLogger(__name__)

try:  # use the reference's own types when its package is importable, so isinstance / except clauses line up
    from audio_pipeline.exceptions import AudioProcessingError  # type: ignore
    from audio_pipeline.protocols import TimestampMapping  # type: ignore
except Exception:  # pragma: no cover - the reference package needs pydub / noisereduce at import time
    from ._compat import AudioProcessingError, TimestampMapping

from ._compat import NoiseReductionConfig

__all__ = ["GpuAudioPreprocessor"]


def _align8(n: int) -> int:
    return (int(n) + 7) // 8 * 8


class _Resident:
    """One recording on the device: ``buf`` holds ``n`` int16 samples (capacity a multiple of 8, the plans' item alignment)."""
    __slots__ = ("buf", "n", "sr", "job", "stat", "channels", "width")

    def __init__(self, buf: torch.Tensor, n: int, sr: int, job=None, stat=None, channels: int = 1, width: int = 2):
        self.buf, self.n, self.sr, self.job, self.stat, self.channels, self.width = buf, n, sr, job, stat, channels, width

    @property
    def pcm(self) -> torch.Tensor:
        return self.buf[: self.n]


class GpuAudioPreprocessor:
    """Noise reduction, peak / loudness normalisation and silence removal on a B200."""

    def __init__(self, sample_rate: int, temp_dir: str, noise_config=None, device=None, stationary: bool = True,
                 write_mode: str = "sync", max_resident_bytes: int = 2 << 30, max_plans: int = 6):
        if write_mode not in ("sync", "deferred"):
            raise ValueError("write_mode must be 'sync' or 'deferred'")
        self.sample_rate = sample_rate
        self.temp_dir = temp_dir
        self.noise_config = noise_config or NoiseReductionConfig()
        self.stationary = stationary
        self.write_mode = write_mode
        self._device = device
        self._fe: Optional[FrontEnd] = None
        self._resident: "OrderedDict[str, _Resident]" = OrderedDict()
        self._max_resident_bytes = int(max_resident_bytes)
        self._plans: "OrderedDict[tuple, object]" = OrderedDict()
        self._max_plans = max(1, int(max_plans))
        self._in_pool = wavio.PinnedPool()       # page-locked staging: file -> pinned -> device
        self._out_slots = [[wavio.PinnedPool(), None] for _ in range(4)]      # [landing buffer, job using it]
        self._out_next = 0
        self._writer: Optional[wavio.WavWriter] = None
        self._jobs: List[wavio.WriteJob] = []
        os.makedirs(temp_dir, exist_ok=True)

    @classmethod
    def from_config(cls, config, **kw) -> "GpuAudioPreprocessor":
        """preprocessor.py:62-69."""
        return cls(sample_rate=config.audio.sample_rate, temp_dir=config.temp_dir, noise_config=config.noise_reduction, **kw)

    # ---- plumbing ---------------------------------------------------------------------------------------------------
    @property
    def frontend(self) -> FrontEnd:
        if self._fe is None:
            self._fe = FrontEnd(self._device)
        return self._fe

    def _sync(self) -> None:
        dev = self.frontend.device
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()

    def _event(self):
        dev = self.frontend.device
        if dev.type != "cuda":
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        return ev

    def read_wave(self, path: str) -> Tuple[bytes, int]:
        """preprocessor.py:71-79 (channel count and sample width are ignored there too)."""
        try:
            with contextlib.closing(wave.open(path, "rb")) as wf:
                sr = wf.getframerate()
                pcm = wf.readframes(wf.getnframes())
            return pcm, sr
        except Exception as e:
            raise AudioProcessingError(f"Failed to read WAV file: {path}", details=str(e))

    def write_wave(self, path: str, audio: bytes, sample_rate: int) -> None:
        """preprocessor.py:81-90: always mono, 16-bit."""
        try:
            with wave.open(path, "wb") as wf:
                wf.setnchannels(1)
                wf.setsampwidth(2)
                wf.setframerate(sample_rate)
                wf.writeframes(audio)
        except Exception as e:
            raise AudioProcessingError(f"Failed to write WAV file: {path}", details=str(e))

    # -- plans ----------------------------------------------------------------------------------------------------------
    def _plan(self, lengths, **kw):
        """A cached batch plan for this geometry / mode (plan creation builds and uploads the item / chunk tables and
        allocates the workspace: milliseconds the kernels of a short file do not take)."""
        key = (tuple(int(n) for n in lengths), tuple(sorted(kw.items())))
        plan = self._plans.get(key)
        if plan is not None:
            self._plans.move_to_end(key)
            return plan
        plan = self.frontend.plan(list(lengths), **kw)
        self._plans[key] = plan
        while len(self._plans) > self._max_plans:
            _, old = self._plans.popitem(last=False)
            old.close()
        return plan

    def _stage_plan(self, n: int, sr: int, **kw):
        """The plan the three stage methods of one recording share (denoise / peak normalise / loudness are entry points of
        the same plan; n_mels = 0)."""
        base = dict(sample_rate=sr, stationary=self.stationary, auto_detect_noise=bool(self.noise_config.auto_detect_noise),
                    noise_sample_duration_s=float(self.noise_config.noise_sample_duration_s), target_lufs=-16.0, n_mels=0)
        base.update(kw)
        return self._plan([n], **base)

    # -- residents ------------------------------------------------------------------------------------------------------
    def _remember(self, path: str, res: _Resident) -> None:
        self._resident.pop(path, None)
        self._resident[path] = res
        total = sum(2 * r.buf.numel() for r in self._resident.values())
        while total > self._max_resident_bytes and len(self._resident) > 1:
            _, old = self._resident.popitem(last=False)
            total -= 2 * old.buf.numel()

    def _lookup(self, path: str) -> Optional[_Resident]:
        """The device copy of `path` if it still describes the file: our own write still in flight, or a finished file whose
        (mtime, size) are the ones recorded when it was written / read.  temp_dir may be wiped at any time
        (media_handler.py:98-102) and foreign stages may rewrite a file."""
        res = self._resident.get(path)
        if res is None:
            return None
        ok = False
        try:
            if res.job is not None and not res.job.done.is_set():
                ok = True
            elif res.job is not None and res.job.error is not None:
                ok = False
            else:
                st = os.stat(path)
                want = res.job.stat if res.job is not None else res.stat
                ok = want is not None and (st.st_mtime_ns, st.st_size) == want
        except OSError:
            ok = False
        if not ok:
            self._resident.pop(path, None)
            return None
        self._resident.move_to_end(path)
        return res

    def _device_buffer(self, n: int) -> torch.Tensor:
        return torch.empty(max(_align8(n), 8), dtype=torch.int16, device=self.frontend.device)

    def _load(self, path: str):
        """(_Resident, channels, sample width) of `path`: the resident copy of an earlier stage's output (mono, 16-bit), or the
        file read straight into page-locked memory (wavio.py) and uploaded with one DMA.  read_wave's semantics: the data
        chunk as is -- the caller decides what channel count / width mean."""
        hit = self._lookup(path)
        if hit is not None:
            return hit, hit.channels, hit.width
        try:
            x, info = wavio.read_pcm16(path, self._in_pool)
            st = os.stat(path)
        except Exception as e:
            raise AudioProcessingError(f"Failed to read WAV file: {path}", details=str(e))
        n = x.numel()
        buf = self._device_buffer(n)
        buf[:n].copy_(x, non_blocking=True)
        if _align8(n) > n:
            buf[n:].zero_()
        self._sync()                                 # the pinned buffer is reused by the next read
        res = _Resident(buf, n, info.rate, None, (st.st_mtime_ns, st.st_size), info.channels, info.sampwidth)
        self._remember(path, res)
        return res, info.channels, info.sampwidth

    def _upload(self, path: str) -> Tuple[torch.Tensor, int]:
        """int16 samples of `path` on the device and its rate (read_wave semantics)."""
        res, _, _ = self._load(path)
        return res.pcm, res.sr

    def prefetch(self, paths: Iterable[str], depth: int = 2) -> int:
        """Read and upload `paths` ahead of the Protocol calls that will name them (wavio.WavPrefetcher: the next file is read
        into a second page-locked buffer while the previous one is on its way to the device).  Returns the number of files now
        resident; later calls on those paths start from device memory."""
        n_ok = 0
        dev = self.frontend.device
        for path, x, info in wavio.WavPrefetcher([str(p) for p in paths], depth=max(2, depth)):
            st = os.stat(path)
            n = x.numel()
            buf = self._device_buffer(n)
            buf[:n].copy_(x, non_blocking=True)
            if _align8(n) > n:
                buf[n:].zero_()
            if dev.type == "cuda":
                torch.cuda.current_stream(dev).synchronize()       # one file in flight; the reader is already on the next one
            self._remember(path, _Resident(buf, n, info.rate, None, (st.st_mtime_ns, st.st_size), info.channels, info.sampwidth))
            n_ok += 1
        return n_ok

    # -- output files -----------------------------------------------------------------------------------------------------
    def _store(self, path: str, buf: torch.Tensor, n: int, sr: int) -> None:
        """write_wave's file (:81-90: mono, 16-bit, caller's rate; byte-identical header) written by the background writer from a
        page-locked landing buffer; the device copy stays resident under `path`."""
        if self._writer is None:
            self._writer = wavio.WavWriter(4)
        slot = self._out_slots[self._out_next]
        self._out_next = (self._out_next + 1) % len(self._out_slots)
        if slot[1] is not None:
            slot[1].done.wait()                      # the landing buffer's previous file must be on disk
        host = slot[0].get(n)
        host.copy_(buf[:n], non_blocking=True)
        job = self._writer.submit(path, host, sr, self._event())
        slot[1] = job
        self._jobs = [j for j in self._jobs if not j.done.is_set() or j.error is not None] + [job]
        self._remember(path, _Resident(buf, n, sr, job))
        if self.write_mode == "sync":
            try:
                job.wait()
            except Exception as e:
                raise AudioProcessingError(f"Failed to write WAV file: {path}", details=str(e))

    def flush(self) -> None:
        """Wait for every queued output file (``write_mode="deferred"``); raises AudioProcessingError for a failed write."""
        jobs, self._jobs = self._jobs, []
        first = None
        for j in jobs:
            j.done.wait()
            if j.error is not None and first is None:
                first = j
        if first is not None:
            raise AudioProcessingError(f"Failed to write WAV file: {first.path}", details=str(first.error))

    def close(self) -> None:
        try:
            self.flush()
        finally:
            if self._writer is not None:
                self._writer.close()
                self._writer = None
            for p in self._plans.values():
                p.close()
            self._plans.clear()
            self._resident.clear()

    def __del__(self):
        try:
            if self._writer is not None:
                self._writer.close()
        except Exception:
            pass

    # ---- noise-clip detection (preprocessor.py:92-147) ---------------------------------------------------------------
    def _detect_noise_segments(self, audio: np.ndarray, sr: int) -> List[Tuple[int, int]]:
        """All (start, end) noise segments.  Frame energies / zero crossings come from the S0 kernel; the percentile
        thresholds and the run decoding of :125-145 are repeated here on the (small) per-frame arrays."""
        x = torch.from_numpy(np.asarray(audio).astype(np.int16)).to(self.frontend.device)
        plan = self._plan([x.numel()], sample_rate=sr, n_mels=0)
        e, z = plan.frame_stats(plan.pack([x]))
        energies = e.cpu().numpy().astype(np.float64)
        frame_length, hop_length = int(sr * 0.025), int(sr * 0.010)
        zcrs = z.cpu().numpy() / frame_length
        if len(energies) == 0:
            return []
        noise = (energies < np.percentile(energies, 20)) & (zcrs > np.percentile(zcrs, 50) * 0.5)
        segs, in_noise, start = [], False, 0
        for i, is_noise in enumerate(noise):
            if is_noise and not in_noise:
                start, in_noise = i, True
            elif not is_noise and in_noise:
                if (i - start) * hop_length / sr >= 0.1:
                    segs.append((start * hop_length, i * hop_length))
                in_noise = False
        return segs

    # ---- Protocol methods -----------------------------------------------------------------------------------------
    def reduce_stationary_noise(self, input_wav: str, noise_sample_path: Optional[str] = None) -> str:
        """preprocessor.py:149-214: spectral-gating denoise with an auto-detected, explicit or leading noise clip."""
        if not self.noise_config.enabled:
            logger.info("Noise reduction disabled, skipping")
            return input_wav
        res, _, _ = self._load(input_wav)            # like read_wave (:71-77): channel count and width are not looked at
        n, sr = res.n, res.sr
        try:
            if noise_sample_path:
                noise, _, _ = self._load(noise_sample_path)
                plan = self._stage_plan(n, sr, external_noise=True)
                plan.set_noise_clips([0], [noise.n])
                out = plan.out_buffer()
                plan.denoise(res.buf, noise_pcm=noise.buf, out=out)
            else:
                plan = self._stage_plan(n, sr)
                out = plan.out_buffer()
                _, _, clip = plan.denoise(res.buf, out=out)
                if self.stationary and logger.isEnabledFor(logging.INFO):
                    c = clip.cpu().numpy()[0]
                    logger.info(f"Noise profile: {c[0] / sr:.2f}s - {c[1] / sr:.2f}s")
        except (MafError, RuntimeError, ValueError) as e:
            raise AudioProcessingError("Noise reduction failed", details=str(e))
        out_path = os.path.join(self.temp_dir, f"{Path(input_wav).stem}_denoised.wav")
        self._store(out_path, out, n, sr)
        logger.info(f"Noise reduced: {out_path}")
        return out_path

    def normalize_audio(self, input_wav: str) -> str:
        """preprocessor.py:216-242: pydub peak normalise (-0.1 dBFS), then mono / 16-bit at the target rate.

        16-bit mono input at the target rate (the pipeline's contract, media_handler.py:219-223) takes the plan's peak
        normalise; stereo input and / or another sample rate go through the front-door kernel, which reproduces pydub's
        call order over audioop: max / mul on the interleaved samples, ratecv, tomono.  Other sample widths and more than
        two channels raise AudioProcessingError (out of scope)."""
        try:
            res, ch, width = self._load(input_wav)
            if width != 2:
                raise ValueError(f"unsupported sample width {width} (16-bit PCM expected)")
            sr = res.sr
            if sr != self.sample_rate or ch > 1:
                if ch > 2:
                    raise ValueError(f"{ch} channels: only mono / stereo input is supported")
                y = self.frontend.front_door(res.pcm, ch, sr, self.sample_rate, normalize=True)
                n = int(y.numel())
                out = self._device_buffer(n)
                out[:n].copy_(y)
            else:
                n = res.n
                plan = self._stage_plan(n, sr)
                out = plan.out_buffer()
                plan.peak_normalize(res.buf, out=out)
        except AudioProcessingError as e:            # a read failure keeps this method's message, like the reference's blanket except
            raise AudioProcessingError("Audio normalization failed", details=str(e))
        except Exception as e:
            raise AudioProcessingError("Audio normalization failed", details=str(e))
        out_path = os.path.join(self.temp_dir, f"{Path(input_wav).stem}_norm.wav")
        self._store(out_path, out, n, self.sample_rate)
        logger.info(f"Audio normalized: {out_path}")
        return out_path

    def normalize_loudness(self, input_wav: str, target_lufs: float = -16.0) -> str:
        """preprocessor.py:244-288: BS.1770 integrated loudness -> gain to target -> peak limiter -> int16.  Multi-channel input
        is metered and written as the float32 channel mean, like :260-261."""
        try:
            res, ch, width = self._load(input_wav)
            if width != 2:
                raise ValueError(f"unsupported sample width {width} (16-bit PCM expected)")
            if ch > 1:
                y, lufs = self.frontend.loudness_interleaved(res.pcm, ch, self.sample_rate, target_lufs)
                n = int(y.numel())
                out = self._device_buffer(n)
                out[:n].copy_(y)
            else:
                n = res.n
                plan = self._stage_plan(n, self.sample_rate, target_lufs=float(target_lufs))
                out = plan.out_buffer()
                _, lufs = plan.loudness(res.buf, out=out)
            loudness = float(lufs.cpu()[0])
            if not np.isfinite(loudness) or loudness < -70:
                logger.warning("Audio is too quiet for LUFS normalization, skipping")
                return input_wav
        except Exception as e:
            raise AudioProcessingError("Loudness normalization failed", details=str(e))
        out_path = os.path.join(self.temp_dir, f"{Path(input_wav).stem}_loudnorm.wav")
        self._store(out_path, out, n, self.sample_rate)
        logger.info(f"Loudness normalized to {target_lufs} LUFS: {out_path}")
        return out_path

    def remove_silence(self, input_wav: str, min_silence_len: int = 250, silence_offset_db: float = 40.0,
                       silence_margin: int = 100, preserve_timestamps: bool = True):
        """preprocessor.py:290-375.  See silence.py."""
        from .silence import remove_silence
        try:
            res, ch, width = self._load(input_wav)
            if ch != 1 or width != 2:
                raise ValueError(f"{ch} channel(s) / {8 * width}-bit input: 16-bit mono expected (normalize_audio produces it)")
            out = remove_silence(self.frontend, res.pcm, res.sr, min_silence_len, silence_offset_db, silence_margin,
                                 preserve_timestamps, TimestampMapping)
        except Exception as e:
            raise AudioProcessingError("Silence removal failed", details=str(e))
        if out is None:
            logger.warning("No non-silent segments found, returning original")
            if self.write_mode == "deferred":
                self.flush()                         # the returned path may be a file of ours that is still being written
            return input_wav, []
        cleaned, mappings = out
        n = int(cleaned.numel())
        base = cleaned._base if cleaned._base is not None else cleaned     # the compaction buffer (capacity >= n)
        if base.numel() < _align8(n) or cleaned.data_ptr() != base.data_ptr():
            base = self._device_buffer(n)
            base[:n].copy_(cleaned)
        out_path = os.path.join(self.temp_dir, f"{Path(input_wav).stem}_nosilence.wav")
        self._store(out_path, base, n, res.sr)
        if self.write_mode == "deferred":
            self.flush()                             # last Protocol call of pipeline.py:331-357: foreign stages read files from here on
        if logger.isEnabledFor(logging.INFO):
            original, processed = res.n / res.sr, n / res.sr
            removed = original - processed
            logger.info(f"Silence removed: {out_path} (removed {removed:.1f}s, {removed / max(original, 1e-9) * 100:.1f}%)")
        return out_path, mappings

    def detect_silence_segments(self, input_wav: str, min_silence_len: int = 500, silence_offset_db: float = 40.0):
        """preprocessor.py:377-405."""
        from .silence import detect_silence_ms
        x, sr = self._upload(input_wav)
        return [(s / 1000.0, e / 1000.0) for s, e in detect_silence_ms(self.frontend, x, sr, min_silence_len, silence_offset_db)]

    # ---- fused in-memory entry ------------------------------------------------------------------------------------------
    def process(self, pcm: torch.Tensor, n_mels: int = 80, mel_padding: int = 0, target_lufs: float = -16.0):
        """denoise -> peak normalise -> loudness normalise -> log-mel on one int16 recording already in memory.

        Returns (int16 device tensor, LUFS before the gain, float32 [n_mels, frames] device tensor)."""
        x = pcm.to(self.frontend.device)
        plan = self._plan([x.numel()], sample_rate=self.sample_rate, stationary=self.stationary,
                          auto_detect_noise=bool(self.noise_config.auto_detect_noise),
                          noise_sample_duration_s=float(self.noise_config.noise_sample_duration_s),
                          target_lufs=float(target_lufs), n_mels=n_mels, mel_padding=mel_padding)
        res = plan.process(plan.pack([x]))
        return res.pcm[0], float(res.lufs.cpu()[0]), (res.log_mel[0] if n_mels else None)

    def process_files(self, paths, batch_size: int = 64, n_mels: int = 80, mel_padding: int = 0, target_lufs: float = -16.0,
                      write_output: bool = True):
        """The orchestrator's chain (pipeline.py:331-347 + the transcriber's log-mel) for MANY files, one fused launch sequence per
        batch.  Batch k + 1 is read (``wavio.BatchReader``: every file's data chunk straight into a page-locked batch buffer at
        its plan offset, files in parallel) while batch k is uploaded and processed, and batch k's output files are written by
        the background writer (files in parallel) while batch k + 1 runs; with ``write_output`` every file gets what the chained
        Protocol calls would have left in ``temp_dir`` (``<stem>_denoised_norm_loudnorm.wav``; ``<stem>_norm_loudnorm.wav`` when
        noise reduction is disabled; the ``_norm`` file when the recording is too quiet for the loudness stage, :267-269).
        Files must be 16-bit mono at ``self.sample_rate`` (the pipeline's contract, media_handler.py:219-223).  Yields
        ``(output path or None, LUFS, log-mel CUDA tensor [n_mels, frames] or None)`` in input order, a batch at a time; a
        yielded path's file is complete.  Results are bit-identical to the chained calls."""
        paths = [str(p) for p in paths]
        fe = self.frontend
        dev = fe.device
        denoise = bool(self.noise_config.enabled)
        if write_output and self._writer is None:
            self._writer = wavio.WavWriter(4)
        out_pools = [wavio.PinnedPool(), wavio.PinnedPool()]
        pending = None                               # (group, lufs host tensor, mel views, jobs or None, stems, event)

        def finish(p):
            group, lufs_h, mels, jobs, outs, ev = p
            if ev is not None:
                ev.synchronize()
            lufs_np = lufs_h.numpy()
            res = []
            for k in range(len(group)):
                path = None
                if jobs is not None:
                    applied = bool(np.isfinite(lufs_np[k])) and not lufs_np[k] < -70
                    path = outs[k][0] if applied else outs[k][1]
                res.append((path, float(lufs_np[k]), mels[k] if n_mels else None))
            if jobs is not None:
                for j in jobs(lufs_np):
                    try:
                        j.wait()
                    except Exception as e:
                        raise AudioProcessingError(f"Failed to write WAV file: {j.path}", details=str(e))
            return res

        try:
            reader = wavio.BatchReader(paths, batch_size)
            for bi, (group, infos, lengths, offs, h_in) in enumerate(reader):
                for p, i in zip(group, infos):
                    if i.channels != 1 or i.sampwidth != 2 or i.rate != self.sample_rate:
                        raise ValueError(f"{p}: {i.channels} ch / {8 * i.sampwidth} bit / {i.rate} Hz; the fused batch entry takes "
                                         f"16-bit mono at {self.sample_rate} Hz (use normalize_audio for other input)")
                plan = self._plan(lengths, sample_rate=self.sample_rate, stationary=self.stationary,
                                  auto_detect_noise=bool(self.noise_config.auto_detect_noise),
                                  noise_sample_duration_s=float(self.noise_config.noise_sample_duration_s),
                                  target_lufs=float(target_lufs), n_mels=n_mels, mel_padding=mel_padding)
                d_in = h_in.to(dev, non_blocking=True)
                ev_in = self._event()
                if denoise:
                    res = plan.process(d_in)
                    pcm_out, mels, lufs = res.pcm, res.log_mel, res.lufs
                else:
                    norm = plan.peak_normalize(d_in)
                    pcm_out, lufs = plan.loudness(plan.pack(norm))
                    mels = plan.log_mel(plan.pack(pcm_out)) if n_mels else []
                lufs_h = torch.empty(len(group), dtype=torch.float64).pin_memory() if dev.type == "cuda" else torch.empty(len(group), dtype=torch.float64)
                lufs_h.copy_(lufs, non_blocking=True)
                jobs = outs = None
                if write_output:
                    base = pcm_out[0]._base if pcm_out[0]._base is not None else pcm_out[0]
                    h_out = out_pools[bi % 2].get(max(plan.total_out, 8))
                    h_out.copy_(base[: h_out.numel()], non_blocking=True)
                    outs = []
                    for p in group:
                        stem = Path(p).stem + ("_denoised" if denoise else "")
                        outs.append((os.path.join(self.temp_dir, f"{stem}_norm_loudnorm.wav"), os.path.join(self.temp_dir, f"{stem}_norm.wav")))
                ev_out = self._event()
                if write_output:
                    lays = [(lay.out_off, n) for lay, n in zip(plan.layout, lengths)]

                    def jobs(lufs_np, h_out=h_out, outs=outs, lays=lays):
                        # queued once the LUFS are on the host: the name of a file depends on whether its gain was applied
                        js = []
                        for k, (off, n) in enumerate(lays):
                            applied = bool(np.isfinite(lufs_np[k])) and not lufs_np[k] < -70
                            js.append(self._writer.submit(outs[k][0] if applied else outs[k][1], h_out[off: off + n], self.sample_rate, None))
                        return js
                if ev_in is not None:
                    ev_in.synchronize()              # the reader may now reuse this batch's page-locked input buffer
                done, pending = pending, (group, lufs_h, mels, jobs, outs, ev_out)
                if done is not None:                 # batch k - 1: its copies finished while batch k was being issued
                    yield from finish(done)
            if pending is not None:
                done, pending = pending, None
                yield from finish(done)
        except AudioProcessingError:
            raise
        except Exception as e:
            raise AudioProcessingError("Batch processing failed", details=str(e))

"""GpuAudioPreprocessor -- drop-in for ``audio_pipeline.preprocessor.AudioPreprocessor`` (reference
``audio_pipeline/preprocessor.py:35-405``) behind ``PreprocessorProtocol`` (``protocols.py:68-90``).

Same constructor, same method names, argument meaning, output file names and error behaviour as the reference
class, so ``AudioPipeline(config, preprocessor=GpuAudioPreprocessor.from_config(config))`` runs unchanged.
Every numerical stage is a CUDA kernel of libmaf_b200.so; this file only does WAV I/O and path bookkeeping.

Beyond the Protocol ("fast path"): results of a stage stay resident on the GPU keyed by the
output path, so the next Protocol call on that path skips the disk read and the PCIe upload; ``process()`` is the
fused in-memory entry (denoise -> peak normalise -> loudness -> log-mel in one call).
"""
from __future__ import annotations

import contextlib
import logging
import os
import wave
from pathlib import Path
from typing import Dict, List, Optional, Tuple

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


class GpuAudioPreprocessor:
    """Noise reduction, peak / loudness normalisation and silence removal on a B200."""

    def __init__(self, sample_rate: int, temp_dir: str, noise_config=None, device=None, stationary: bool = True):
        self.sample_rate = sample_rate
        self.temp_dir = temp_dir
        self.noise_config = noise_config or NoiseReductionConfig()
        self.stationary = stationary
        self._device = device
        self._fe: Optional[FrontEnd] = None
        self._resident: Dict[str, Tuple[torch.Tensor, int, float]] = {}   # path -> (int16 device tensor, sr, mtime)
        self._in_pool = wavio.PinnedPool()       # page-locked staging: file -> pinned -> device, device -> pinned -> file
        self._out_pool = wavio.PinnedPool()
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

    def _read_wav_info(self, path: str):
        with contextlib.closing(wave.open(path, "rb")) as wf:
            return wf.getnchannels(), wf.getsampwidth(), wf.getframerate(), wf.readframes(wf.getnframes())

    def _upload(self, path: str) -> Tuple[torch.Tensor, int]:
        """int16 samples of `path` on the device (resident copy of an earlier stage's output when still valid)."""
        hit = self._resident.get(path)
        if hit is not None:
            try:
                if os.path.getmtime(path) == hit[2]:
                    return hit[0], hit[1]
            except OSError:
                pass
            self._resident.pop(path, None)       # temp_dir may have been wiped (media_handler.py:98-102)
        # read_wave's semantics (:71-79: the data chunk as is, channel count and width ignored), but read straight into
        # page-locked memory so the upload is one DMA at full PCIe rate (wavio.py)
        dev = self.frontend.device
        try:
            x, info = wavio.read_pcm16(path, self._in_pool)
        except Exception as e:
            raise AudioProcessingError(f"Failed to read WAV file: {path}", details=str(e))
        xd = x.to(dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()         # the pinned buffer is reused by the next read
        return xd, info.rate

    def _store(self, path: str, dev: torch.Tensor, sr: int) -> None:
        """write_wave's file (:81-90: mono, 16-bit, caller's rate; byte-identical header) from a pinned landing buffer."""
        try:
            host = self._out_pool.get(dev.numel())
            host.copy_(dev.reshape(-1), non_blocking=True)
            torch.cuda.current_stream(dev.device).synchronize()
            wavio.write_pcm16(path, host, sr)
        except Exception as e:
            raise AudioProcessingError(f"Failed to write WAV file: {path}", details=str(e))
        self._resident.clear()                   # keep exactly one stage output resident
        self._resident[path] = (dev, sr, os.path.getmtime(path))

    # ---- noise-clip detection (preprocessor.py:92-147) ---------------------------------------------------------------
    def _detect_noise_segments(self, audio: np.ndarray, sr: int) -> List[Tuple[int, int]]:
        """All (start, end) noise segments.  Frame energies / zero crossings come from the S0 kernel; the percentile
        thresholds and the run decoding of :125-145 are repeated here on the (small) per-frame arrays."""
        x = torch.from_numpy(np.asarray(audio).astype(np.int16)).to(self.frontend.device)
        plan = self.frontend.plan([x.numel()], sample_rate=sr, n_mels=0)
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
        x, sr = self._upload(input_wav)
        try:
            fe = self.frontend
            if noise_sample_path:
                noise, _ = self._upload(noise_sample_path)
                plan = fe.plan([x.numel()], sample_rate=sr, stationary=self.stationary, external_noise=True, n_mels=0)
                plan.set_noise_clips([0], [noise.numel()])
                out, _, _ = plan.denoise(plan.pack([x]), noise_pcm=noise.contiguous())
            else:
                plan = fe.plan([x.numel()], sample_rate=sr, stationary=self.stationary,
                               auto_detect_noise=bool(self.noise_config.auto_detect_noise),
                               noise_sample_duration_s=float(self.noise_config.noise_sample_duration_s), n_mels=0)
                out, _, clip = plan.denoise(plan.pack([x]))
                if self.stationary:
                    c = clip.cpu().numpy()[0]
                    logger.info(f"Noise profile: {c[0] / sr:.2f}s - {c[1] / sr:.2f}s")
            den = out[0].clone()
            plan.close()
        except (MafError, RuntimeError, ValueError) as e:
            raise AudioProcessingError("Noise reduction failed", details=str(e))
        out_path = os.path.join(self.temp_dir, f"{Path(input_wav).stem}_denoised.wav")
        self._store(out_path, den, sr)
        logger.info(f"Noise reduced: {out_path}")
        return out_path

    def normalize_audio(self, input_wav: str) -> str:
        """preprocessor.py:216-242: pydub peak normalise (-0.1 dBFS), then mono / 16-bit at the target rate.

        16-bit mono input at the target rate (the pipeline's contract, media_handler.py:219-223) takes the plan's peak
        normalise; stereo input and / or another sample rate go through the front-door kernel, which reproduces pydub's
        call order over audioop: max / mul on the interleaved samples, ratecv, tomono.  Other sample widths and more than
        two channels raise AudioProcessingError."""
        try:
            hit = self._resident.get(input_wav)
            if hit is not None and os.path.exists(input_wav) and os.path.getmtime(input_wav) == hit[2]:
                x, sr, ch = hit[0], hit[1], 1
            else:
                ch, width, sr, raw = self._read_wav_info(input_wav)
                if width != 2:
                    raise ValueError(f"unsupported sample width {width} (16-bit PCM expected)")
                x = torch.frombuffer(bytearray(raw), dtype=torch.int16).to(self.frontend.device)
            if sr != self.sample_rate or ch > 1:
                if ch > 2:
                    raise ValueError(f"{ch} channels: only mono / stereo input is supported")
                y = self.frontend.front_door(x.contiguous(), ch, sr, self.sample_rate, normalize=True)
            else:
                plan = self.frontend.plan([x.numel()], sample_rate=sr, n_mels=0)
                y = plan.peak_normalize(plan.pack([x]))[0].clone()
                plan.close()
        except Exception as e:
            raise AudioProcessingError("Audio normalization failed", details=str(e))
        out_path = os.path.join(self.temp_dir, f"{Path(input_wav).stem}_norm.wav")
        self._store(out_path, y, self.sample_rate)
        logger.info(f"Audio normalized: {out_path}")
        return out_path

    def normalize_loudness(self, input_wav: str, target_lufs: float = -16.0) -> str:
        """preprocessor.py:244-288: BS.1770 integrated loudness -> gain to target -> peak limiter -> int16."""
        try:
            x, sr = self._upload(input_wav)
            plan = self.frontend.plan([x.numel()], sample_rate=self.sample_rate, target_lufs=target_lufs, n_mels=0)
            out, lufs = plan.loudness(plan.pack([x]))
            loudness = float(lufs.item())
            y = out[0].clone()
            plan.close()
            if not np.isfinite(loudness) or loudness < -70:
                logger.warning("Audio is too quiet for LUFS normalization, skipping")
                return input_wav
        except Exception as e:
            raise AudioProcessingError("Loudness normalization failed", details=str(e))
        out_path = os.path.join(self.temp_dir, f"{Path(input_wav).stem}_loudnorm.wav")
        self._store(out_path, y, self.sample_rate)
        logger.info(f"Loudness normalized to {target_lufs} LUFS: {out_path}")
        return out_path

    def remove_silence(self, input_wav: str, min_silence_len: int = 250, silence_offset_db: float = 40.0,
                       silence_margin: int = 100, preserve_timestamps: bool = True):
        """preprocessor.py:290-375.  See silence.py."""
        from .silence import remove_silence
        try:
            x, sr = self._upload(input_wav)
            res = remove_silence(self.frontend, x, sr, min_silence_len, silence_offset_db, silence_margin, preserve_timestamps,
                                 TimestampMapping)
        except Exception as e:
            raise AudioProcessingError("Silence removal failed", details=str(e))
        if res is None:
            logger.warning("No non-silent segments found, returning original")
            return input_wav, []
        cleaned, mappings = res
        out_path = os.path.join(self.temp_dir, f"{Path(input_wav).stem}_nosilence.wav")
        self._store(out_path, cleaned, sr)
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
        plan = self.frontend.plan([x.numel()], sample_rate=self.sample_rate, stationary=self.stationary,
                                  auto_detect_noise=bool(self.noise_config.auto_detect_noise),
                                  noise_sample_duration_s=float(self.noise_config.noise_sample_duration_s),
                                  target_lufs=target_lufs, n_mels=n_mels, mel_padding=mel_padding)
        res = plan.process(plan.pack([x]))
        out = res.pcm[0].clone(), float(res.lufs.item()), (res.log_mel[0].clone() if n_mels else None)
        plan.close()
        return out

    def process_files(self, paths, batch_size: int = 64, n_mels: int = 80, mel_padding: int = 0, target_lufs: float = -16.0,
                      write_output: bool = True):
        """The orchestrator's chain (pipeline.py:331-347 + the transcriber's log-mel) for MANY files, one fused launch sequence per
        batch: every file's data chunk is read straight into a page-locked batch buffer at its plan offset, one H2D copy, the
        kernels, one D2H copy of the final PCM, and -- with ``write_output`` -- the same file the chained Protocol calls
        would have left in ``temp_dir`` (``<stem>_denoised_norm_loudnorm.wav``; ``<stem>_norm_loudnorm.wav`` when noise
        reduction is disabled).  Files must be 16-bit mono at ``self.sample_rate`` (the pipeline's contract,
        media_handler.py:219-223).  Yields ``(output path or None, LUFS, log-mel CUDA tensor [n_mels, frames] or None)``
        in input order; results are bit-identical to the chained calls."""
        paths = [str(p) for p in paths]
        fe = self.frontend
        dev = fe.device
        denoise = bool(self.noise_config.enabled)
        for b0 in range(0, len(paths), max(1, batch_size)):
            group = paths[b0:b0 + max(1, batch_size)]
            try:
                infos = [wavio.read_info(p) for p in group]
                for p, i in zip(group, infos):
                    if i.channels != 1 or i.sampwidth != 2 or i.rate != self.sample_rate:
                        raise ValueError(f"{p}: {i.channels} ch / {8 * i.sampwidth} bit / {i.rate} Hz; the fused batch entry takes "
                                         f"16-bit mono at {self.sample_rate} Hz (use normalize_audio for other input)")
                lengths = [i.data_bytes // 2 for i in infos]
                plan = fe.plan(lengths, sample_rate=self.sample_rate, stationary=self.stationary,
                               auto_detect_noise=bool(self.noise_config.auto_detect_noise),
                               noise_sample_duration_s=float(self.noise_config.noise_sample_duration_s),
                               target_lufs=target_lufs, n_mels=n_mels, mel_padding=mel_padding)
                h_in = self._in_pool.get(max(plan.total_pcm, 8))
                h_in.zero_()
                for p, i, lay, n in zip(group, infos, plan.layout, lengths):
                    wavio.read_data_into(p, i, h_in[lay.pcm_off: lay.pcm_off + n])
                d_in = h_in.to(dev, non_blocking=True)
                if denoise:
                    res = plan.process(d_in)
                    pcm_out, mels, lufs = res.pcm, res.log_mel, res.lufs
                else:
                    norm = plan.peak_normalize(d_in)
                    pcm_out, lufs = plan.loudness(plan.pack(norm))
                    mels = plan.log_mel(plan.pack(pcm_out)) if n_mels else []
                lufs_h = lufs.cpu().numpy()
                mel_out = [m.clone() for m in mels] if n_mels else [None] * len(group)
                outs: List[Optional[str]] = [None] * len(group)
                if write_output:
                    for k, (p, x) in enumerate(zip(group, pcm_out)):
                        host = self._out_pool.get(x.numel())
                        host.copy_(x, non_blocking=True)
                        torch.cuda.current_stream(dev).synchronize()
                        stem = Path(p).stem + ("_denoised" if denoise else "") + "_norm_loudnorm"
                        outs[k] = os.path.join(self.temp_dir, f"{stem}.wav")
                        wavio.write_pcm16(outs[k], host, self.sample_rate)
                else:
                    torch.cuda.current_stream(dev).synchronize()
                plan.close()
            except AudioProcessingError:
                raise
            except Exception as e:
                raise AudioProcessingError("Batch processing failed", details=str(e))
            for k in range(len(group)):
                yield outs[k], float(lufs_h[k]), mel_out[k]

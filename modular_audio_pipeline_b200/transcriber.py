"""Transcriber-input side of the seam: Whisper's log-mel front-end on the GPU (reference call sites
``audio_pipeline/transcriber.py:229-238`` -> ``whisper.transcribe`` -> ``whisper.audio.log_mel_spectrogram`` and
``:421-427`` -> faster-whisper's ``FeatureExtractor``), handed to the model as a device tensor.

* ``GpuLogMelFrontEnd.log_mel_spectrogram(audio, n_mels=80, padding=0, device=None)`` has the signature of
  ``whisper.audio.log_mel_spectrogram`` and returns a ``torch.float32`` CUDA tensor ``[n_mels, frames]``.
* ``GpuLogMelFrontEnd.feature_extractor(waveform, padding=160, chunk_length=None)`` has the call signature of
  faster-whisper's ``FeatureExtractor.__call__`` and returns ``np.ndarray [n_mels, frames]``.
* ``GpuFrontEndTranscriber`` wraps a reference transcriber (``WhisperTranscriber`` / ``FasterWhisperTranscriber``,
  ``TranscriberProtocol``: ``transcribe`` / ``is_loaded`` / ``load_model``) and installs the GPU front-end into the
  loaded model.  The ASR encoder / decoder are not part of this path.
"""
from __future__ import annotations

import sys
import wave
from typing import Any, Dict, Optional, Union

import numpy as np
import torch

from .frontend import FrontEnd

__all__ = ["GpuLogMelFrontEnd", "GpuFeatureExtractor", "GpuFrontEndTranscriber"]

N_SAMPLES = 480000      # whisper.audio.N_SAMPLES: 30 s at 16 kHz
N_FRAMES = 3000


class GpuLogMelFrontEnd:
    def __init__(self, device=None):
        self._device = device
        self._fe: Optional[FrontEnd] = None

    @property
    def frontend(self) -> FrontEnd:
        if self._fe is None:
            self._fe = FrontEnd(self._device)
        return self._fe

    def _to_int16(self, audio: Union[str, np.ndarray, torch.Tensor]) -> torch.Tensor:
        """whisper.load_audio decodes to s16le and divides by 32768; the kernels take the int16 samples directly.
        Float input must therefore be int16-representable (it is, for anything that came from a 16-bit WAV)."""
        dev = self.frontend.device
        if isinstance(audio, str):
            with wave.open(audio, "rb") as wf:
                if wf.getsampwidth() != 2 or wf.getnchannels() != 1 or wf.getframerate() != 16000:
                    raise ValueError("expected a 16 kHz mono 16-bit WAV (the pipeline's contract, media_handler.py:219-223)")
                raw = wf.readframes(wf.getnframes())
            return torch.frombuffer(bytearray(raw), dtype=torch.int16).to(dev)
        t = torch.as_tensor(audio)
        if t.dtype == torch.int16:
            return t.to(dev)
        q = torch.round(t.to(dev, torch.float32) * 32768.0)
        if not torch.equal(q / 32768.0, t.to(dev, torch.float32)) or q.abs().max() > 32768:
            raise ValueError("float audio is not int16-representable; pass the int16 samples or a WAV path")
        return torch.clamp(q, -32768, 32767).to(torch.int16)

    def log_mel_spectrogram(self, audio, n_mels: int = 80, padding: int = 0, device=None) -> torch.Tensor:
        x = self._to_int16(audio)
        plan = self.frontend.plan([x.numel()], n_mels=n_mels, mel_padding=padding)
        mel = plan.log_mel(plan.pack([x]))[0].clone()
        plan.close()
        return mel if device is None else mel.to(device)

    def feature_extractor(self, waveform, padding: int = 160, chunk_length: Optional[int] = None) -> np.ndarray:
        """faster-whisper ``FeatureExtractor.__call__`` arithmetic: log-mel of the WHOLE waveform plus ``padding`` zeros.
        ``chunk_length`` does not touch the samples there -- it only re-sizes the extractor's ``n_samples`` /
        ``nb_max_frames`` bookkeeping, which ``GpuFeatureExtractor`` (below) mirrors on the wrapped object."""
        x = self._to_int16(waveform)
        return self.log_mel_spectrogram(x, self.n_mels, padding).cpu().numpy()

    n_mels = 80

    @staticmethod
    def pad_or_trim(mel: torch.Tensor, length: int = N_FRAMES) -> torch.Tensor:
        """whisper.audio.pad_or_trim along the frame axis (the 30 s window the encoder consumes)."""
        if mel.shape[-1] > length:
            return mel[..., :length]
        if mel.shape[-1] < length:
            return torch.nn.functional.pad(mel, (0, length - mel.shape[-1]))
        return mel

    def windows(self, mel: torch.Tensor, seeks, lens=None, length: int = N_FRAMES, dtype: torch.dtype = torch.float16) -> torch.Tensor:
        """All encoder windows of one recording in one launch: ``out[w] = pad_or_trim(mel[:, seeks[w] : seeks[w] + lens[w]],
        length).to(dtype)`` (the seek loop of whisper/transcribe.py, behind transcriber.py:229-238), float16 or float32,
        ``[n_windows, n_mels, length]`` on the device.  ``mel`` is a float32 CUDA tensor ``[n_mels, frames]``."""
        fe = self.frontend
        if mel.dtype != torch.float32 or mel.dim() != 2 or not mel.is_cuda or mel.device != fe.device:
            raise ValueError("expected a float32 [n_mels, frames] tensor on the front-end's device")
        if dtype not in (torch.float16, torch.float32):
            raise ValueError("dtype must be torch.float16 or torch.float32")
        mel = mel.contiguous()
        seeks_t = torch.as_tensor(seeks, dtype=torch.int32).reshape(-1).to(fe.device)
        lens_t = None if lens is None else torch.as_tensor(lens, dtype=torch.int32).reshape(-1).to(fe.device)
        if lens_t is not None and lens_t.numel() != seeks_t.numel():
            raise ValueError("seeks and lens differ in length")
        n = seeks_t.numel()
        out = torch.empty(n, mel.shape[0], length, dtype=dtype, device=fe.device)
        if n:
            fe.mel_windows(mel, seeks_t, lens_t, length, out)
        return out


class GpuFeatureExtractor:
    """Stands in for ``WhisperModel.feature_extractor`` (faster-whisper ``feature_extractor.FeatureExtractor``): the CALL runs the
    GPU log-mel, every attribute -- ``sampling_rate``, ``hop_length``, ``chunk_length``, ``n_samples``, ``nb_max_frames``,
    ``time_per_frame``, ``mel_filters``, ``n_fft``, which ``WhisperModel.transcribe`` / ``generate_segments`` read off the
    object -- is the wrapped extractor's own, reads and writes alike."""

    def __init__(self, original, front_end: "GpuLogMelFrontEnd"):
        object.__setattr__(self, "_original", original)
        object.__setattr__(self, "_front_end", front_end)

    def __getattr__(self, name):                       # only reached for names not found on the proxy itself
        return getattr(object.__getattribute__(self, "_original"), name)

    def __setattr__(self, name, value):
        setattr(object.__getattribute__(self, "_original"), name, value)

    @property
    def n_mels(self) -> int:
        mf = getattr(self._original, "mel_filters", None)
        return int(mf.shape[0]) if mf is not None else 80

    def __call__(self, waveform, padding: int = 160, chunk_length: Optional[int] = None) -> np.ndarray:
        o = self._original
        if chunk_length is not None:                   # FeatureExtractor.__call__: bookkeeping only, the waveform is not trimmed
            sr = int(getattr(o, "sampling_rate", 16000))
            hop = int(getattr(o, "hop_length", 160))
            o.n_samples = chunk_length * sr
            o.nb_max_frames = o.n_samples // hop
        fe = self._front_end
        x = fe._to_int16(waveform)
        return fe.log_mel_spectrogram(x, self.n_mels, padding).cpu().numpy()


class GpuFrontEndTranscriber:
    """TranscriberProtocol (protocols.py:119-133) adapter: same transcriber, GPU log-mel front-end."""

    def __init__(self, inner, front_end: Optional[GpuLogMelFrontEnd] = None):
        self.inner = inner
        self.front_end = front_end or GpuLogMelFrontEnd()
        self._installed = False
        self._saved_extractor = None                   # (model, original feature_extractor)
        self._saved_whisper = None                     # (module, original log_mel_spectrogram)

    def is_loaded(self) -> bool:
        return self.inner.is_loaded()

    def load_model(self) -> None:
        self.inner.load_model()
        self._install()

    def _uninstall(self) -> None:
        if self._saved_extractor is not None:
            model, orig = self._saved_extractor
            if isinstance(getattr(model, "feature_extractor", None), GpuFeatureExtractor):
                model.feature_extractor = orig
            self._saved_extractor = None
        if self._saved_whisper is not None:
            mod, orig = self._saved_whisper
            if getattr(mod, "log_mel_spectrogram", None) == self.front_end.log_mel_spectrogram:
                mod.log_mel_spectrogram = orig
            self._saved_whisper = None
        self._installed = False

    def unload_model(self) -> None:
        """pipeline.py:532-533 calls this from cleanup(): the model and the process-wide whisper module get their own
        front-end back before the wrapped transcriber drops the model."""
        self._uninstall()
        if hasattr(self.inner, "unload_model"):
            self.inner.unload_model()

    def _install(self) -> None:
        if self._installed:
            return
        model = getattr(self.inner, "_model", None)
        ext = getattr(model, "feature_extractor", None) if model is not None else None
        if ext is not None and not isinstance(ext, GpuFeatureExtractor):       # faster-whisper WhisperModel
            model.feature_extractor = GpuFeatureExtractor(ext, self.front_end)
            self._saved_extractor = (model, ext)
            self._installed = True
        mod = sys.modules.get("whisper.transcribe")                            # openai-whisper: patch the name transcribe() calls
        if mod is not None and hasattr(mod, "log_mel_spectrogram"):
            if mod.log_mel_spectrogram != self.front_end.log_mel_spectrogram:
                self._saved_whisper = (mod, mod.log_mel_spectrogram)
                mod.log_mel_spectrogram = self.front_end.log_mel_spectrogram
            self._installed = True

    def transcribe(self, input_wav: str) -> Dict[str, Any]:
        if not self.inner.is_loaded():
            self.load_model()
        self._install()
        return self.inner.transcribe(input_wav)

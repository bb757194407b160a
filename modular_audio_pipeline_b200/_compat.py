"""Stand-ins for the reference's small value types, used only when ``audio_pipeline`` itself cannot be imported
(its ``__init__`` pulls pydub / noisereduce / pyloudnorm / webrtcvad at import time).  Field names, defaults and
``__str__`` follow the reference: ``exceptions.py:30-61``, ``protocols.py:35-41``, ``config.py:76-82``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


class AudioPipelineError(Exception):
    def __init__(self, message: str, details: Optional[str] = None):
        self.message = message
        self.details = details
        super().__init__(self.message)

    def __str__(self) -> str:
        if self.details:
            return f"{self.message}\nDetails: {self.details}"
        return self.message


class AudioProcessingError(AudioPipelineError):
    """Raised when audio preprocessing fails."""


class TranscriptionError(AudioPipelineError):
    """Raised when transcription fails."""


class ModelLoadError(AudioPipelineError):
    """Raised when a model fails to load."""


@dataclass
class TimestampMapping:
    processed_start: float
    processed_end: float
    original_start: float
    original_end: float


@dataclass
class NoiseReductionConfig:
    enabled: bool = True
    auto_detect_noise: bool = True
    noise_sample_duration_s: float = 0.5
    noise_sample_path: Optional[str] = None

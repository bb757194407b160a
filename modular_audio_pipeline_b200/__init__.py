"""B200-native DSP front-end (spectral-gating denoise, BS.1770 loudness, Whisper log-mel) behind the
PreprocessorProtocol / Transcriber-input seam of AndersonGACFilho/Modular-Audio-Pipeline."""

__version__ = "0.1.0"

__all__ = ["FrontEnd", "Plan", "GpuAudioPreprocessor", "GpuLogMelFrontEnd", "__version__"]


def __getattr__(name):
    if name in ("FrontEnd", "Plan", "FrontEndResult", "HostPipeline", "MafError", "capture_graph"):
        from . import frontend
        return getattr(frontend, name)
    if name == "GpuAudioPreprocessor":
        from .preprocessor import GpuAudioPreprocessor
        return GpuAudioPreprocessor
    if name in ("WavPrefetcher", "PinnedPool"):
        from . import wavio
        return getattr(wavio, name)
    if name in ("GpuLogMelFrontEnd", "GpuFrontEndTranscriber"):
        from . import transcriber
        return getattr(transcriber, name)
    raise AttributeError(name)

"""The reference-facing surface: PreprocessorProtocol / TranscriberProtocol conformance, signatures equal to the
reference class, and (GPU) the path-in / path-out methods in the order AudioPipeline.run() calls them
(pipeline.py:331-357, :372) checked against the oracle."""
import importlib.util
import inspect
import os
import sys
import types
import wave

import numpy as np
import pytest
import torch

from modular_audio_pipeline_b200.synth import speech_noise
from oracle import reference_path as O

REF = "/root/reference"
SR = 16000


def _write_wav(path, pcm, sr=SR, channels=1):
    with wave.open(str(path), "wb") as wf:
        wf.setnchannels(channels)
        wf.setsampwidth(2)
        wf.setframerate(sr)
        wf.writeframes(np.asarray(pcm, dtype=np.int16).tobytes())


def _read_wav(path):
    with wave.open(str(path), "rb") as wf:
        return np.frombuffer(wf.readframes(wf.getnframes()), dtype=np.int16), wf.getframerate(), wf.getnchannels()


def _load_reference_package():
    """Import the reference package with empty stand-ins for its uninstallable numerical dependencies."""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present on this machine")
    for name in ["pydub", "pydub.effects", "pydub.silence", "noisereduce", "pyloudnorm", "webrtcvad"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.AudioSegment = object
            m.normalize = m.detect_nonsilent = m.detect_silence = lambda *a, **k: None
            sys.modules[name] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import audio_pipeline  # noqa: F401
    from audio_pipeline import preprocessor, protocols
    return preprocessor, protocols


def test_protocol_conformance_and_signatures(tmp_path):
    preprocessor, protocols = _load_reference_package()
    from modular_audio_pipeline_b200.preprocessor import GpuAudioPreprocessor
    from modular_audio_pipeline_b200.transcriber import GpuFrontEndTranscriber
    obj = GpuAudioPreprocessor(16000, str(tmp_path))
    assert isinstance(obj, protocols.PreprocessorProtocol)
    ref_cls = preprocessor.AudioPreprocessor
    for name in ["reduce_stationary_noise", "normalize_audio", "normalize_loudness", "remove_silence",
                 "detect_silence_segments", "read_wave", "write_wave", "_detect_noise_segments"]:
        want = inspect.signature(getattr(ref_cls, name))
        got = inspect.signature(getattr(GpuAudioPreprocessor, name))
        assert [(p.name, p.default) for p in want.parameters.values()] == [(p.name, p.default) for p in got.parameters.values()], name
    # constructor: same leading parameters (extra keyword-only-style extras come after)
    want = list(inspect.signature(ref_cls.__init__).parameters)
    got = list(inspect.signature(GpuAudioPreprocessor.__init__).parameters)
    assert got[:len(want)] == want

    class Inner:
        def transcribe(self, p): return {"segments": []}
        def is_loaded(self): return True
        def load_model(self): pass
    assert isinstance(GpuFrontEndTranscriber(Inner()), protocols.TranscriberProtocol)


def test_from_config_and_disabled_passthrough(tmp_path):
    from modular_audio_pipeline_b200._compat import NoiseReductionConfig
    from modular_audio_pipeline_b200.preprocessor import GpuAudioPreprocessor
    cfg = types.SimpleNamespace(audio=types.SimpleNamespace(sample_rate=16000), temp_dir=str(tmp_path / "t"),
                                noise_reduction=NoiseReductionConfig(enabled=False))
    pre = GpuAudioPreprocessor.from_config(cfg)
    assert os.path.isdir(cfg.temp_dir)
    assert pre.reduce_stationary_noise("whatever.wav") == "whatever.wav"          # preprocessor.py:166-168


def test_errors_are_audio_processing_errors(tmp_path):
    from modular_audio_pipeline_b200._compat import AudioProcessingError
    from modular_audio_pipeline_b200.preprocessor import GpuAudioPreprocessor
    pre = GpuAudioPreprocessor(16000, str(tmp_path))
    with pytest.raises(Exception) as ei:
        pre.read_wave(str(tmp_path / "missing.wav"))
    assert type(ei.value).__name__ == "AudioProcessingError" and "Failed to read WAV file" in str(ei.value)
    assert issubclass(AudioProcessingError, Exception)


def test_oracle_silence_detection_sanity():
    pcm = speech_noise(6, 2, sigma=0.0002).numpy()        # near-silent gaps: 40 dB below the clip's RMS
    rng = O.detect_nonsilent_ms(pcm, SR)
    assert rng and all(a < b for a, b in rng)
    res = O.remove_silence(pcm, SR)
    assert res is not None and len(res[0]) < len(pcm)


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_pipeline_call_sequence_against_oracle(tmp_path, frontend):
    """reduce_stationary_noise -> normalize_audio -> normalize_loudness -> remove_silence(preserve_timestamps=True),
    file in / file out, then the transcriber front-end on the result."""
    from modular_audio_pipeline_b200.preprocessor import GpuAudioPreprocessor
    from modular_audio_pipeline_b200.transcriber import GpuLogMelFrontEnd
    pcm = speech_noise(20, 6, sigma=0.001).numpy()
    src = tmp_path / "meeting.wav"
    _write_wav(src, pcm)
    pre = GpuAudioPreprocessor(SR, str(tmp_path / "tmp"))
    p1 = pre.reduce_stationary_noise(str(src))
    assert os.path.basename(p1) == "meeting_denoised.wav"
    d1, sr, ch = _read_wav(p1)
    _, want1, _ = O.denoise_int16(pcm, SR)
    diff = np.abs(d1.astype(np.int32) - want1.astype(np.int32))
    assert (sr, ch) == (SR, 1) and diff.max() <= 1 and np.mean(diff != 0) <= 1e-3
    p2 = pre.normalize_audio(p1)
    assert os.path.basename(p2) == "meeting_denoised_norm.wav"
    d2, _, _ = _read_wav(p2)
    assert np.array_equal(d2, O.peak_normalize(d1))
    p3 = pre.normalize_loudness(p2)
    assert os.path.basename(p3) == "meeting_denoised_norm_loudnorm.wav"
    d3, _, _ = _read_wav(p3)
    want3, _, applied = O.normalize_loudness(d2, SR)
    diff = np.abs(d3.astype(np.int32) - want3.astype(np.int32))
    assert applied and diff.max() <= 1 and np.mean(diff != 0) <= 1e-3
    p4, maps = pre.remove_silence(p3, preserve_timestamps=True)
    assert os.path.basename(p4) == "meeting_denoised_norm_loudnorm_nosilence.wav"
    d4, _, _ = _read_wav(p4)
    want4, want_maps = O.remove_silence(d3, SR)
    assert len(want_maps) > 1 and len(want4) < len(d3)
    assert np.array_equal(d4, want4)
    assert [(m.processed_start, m.processed_end, m.original_start, m.original_end) for m in maps] == want_maps
    assert pre.detect_silence_segments(p3, 250) == [(a / 1000.0, b / 1000.0) for a, b in O.detect_silence_ms(d3, SR, 250)]
    # transcriber front-end on the path, like whisper.transcribe (padding = 30 s)
    mel = GpuLogMelFrontEnd().log_mel_spectrogram(p4, 80, 480000)
    want = O.log_mel_spectrogram(d4.astype(np.float32) / 32768.0, 80, 480000)
    assert mel.is_cuda and np.abs(mel.cpu().numpy() - want).max() <= 1e-4


@pytest.mark.gpu
def test_explicit_noise_sample_and_quiet_passthrough(tmp_path, frontend):
    from modular_audio_pipeline_b200.preprocessor import GpuAudioPreprocessor
    pcm = speech_noise(8, 3).numpy()
    noise = speech_noise(1.0, 77, sigma=0.02).numpy()
    _write_wav(tmp_path / "a.wav", pcm)
    _write_wav(tmp_path / "n.wav", noise)
    pre = GpuAudioPreprocessor(SR, str(tmp_path / "tmp"))
    out, _, _ = _read_wav(pre.reduce_stationary_noise(str(tmp_path / "a.wav"), str(tmp_path / "n.wav")))
    want = O.to_int16_trunc(O.reduce_noise(pcm.astype(np.float32), SR, noise.astype(np.float32)))
    diff = np.abs(out.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1 and np.mean(diff != 0) <= 1e-3
    _write_wav(tmp_path / "q.wav", np.zeros(SR, dtype=np.int16))
    assert pre.normalize_loudness(str(tmp_path / "q.wav")) == str(tmp_path / "q.wav")     # preprocessor.py:267-269
    _write_wav(tmp_path / "s.wav", pcm[:3000])
    with pytest.raises(Exception) as ei:                                                  # pyloudnorm ValueError -> :287-288
        pre.normalize_loudness(str(tmp_path / "s.wav"))
    assert type(ei.value).__name__ == "AudioProcessingError"


@pytest.mark.gpu
def test_normalize_audio_front_door_48k_stereo(tmp_path, frontend):
    """normalize_audio on a 48 kHz stereo WAV (BASELINE config 5 input): pydub's normalize -> set_frame_rate(16000) ->
    set_channels(1) chain over audioop, bit for bit (the oracle restatement is pinned against stdlib audioop)."""
    from modular_audio_pipeline_b200.preprocessor import GpuAudioPreprocessor
    left = speech_noise(6.0, 21, sr=48000).numpy()
    right = np.roll(speech_noise(6.0, 22, sr=48000, sigma=0.05).numpy(), 777)
    inter = np.stack([left, right], axis=1).reshape(-1)
    _write_wav(tmp_path / "st48.wav", inter, sr=48000, channels=2)
    pre = GpuAudioPreprocessor(SR, str(tmp_path / "tmp"))
    out, sr, ch = _read_wav(pre.normalize_audio(str(tmp_path / "st48.wav")))
    assert (sr, ch) == (SR, 1)
    assert np.array_equal(out, O.front_door(inter, 2, 48000, SR))
    # and the rest of the chain runs on the result
    loud, _, _ = _read_wav(pre.normalize_loudness(pre.normalize_audio(str(tmp_path / "st48.wav"))))
    want16, _, _ = O.normalize_loudness(O.front_door(inter, 2, 48000, SR), SR)
    d = np.abs(loud.astype(np.int32) - want16.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) <= 5e-3


@pytest.mark.gpu
def test_transcriber_adapter_with_fake_models(frontend):
    from modular_audio_pipeline_b200.transcriber import GpuFrontEndTranscriber, GpuLogMelFrontEnd
    x = speech_noise(3, 1)

    class FakeFW:                                     # faster-whisper style model
        def __init__(self): self.feature_extractor = types.SimpleNamespace(mel_filters=np.zeros((128, 201)))

    class Inner:
        def __init__(self): self._model = FakeFW()
        def is_loaded(self): return True
        def load_model(self): pass
        def transcribe(self, p):
            feats = self._model.feature_extractor(x.numpy().astype(np.float32) / 32768.0, padding=160)
            return {"segments": [], "features": feats}
    t = GpuFrontEndTranscriber(Inner(), GpuLogMelFrontEnd())
    feats = t.transcribe("x.wav")["features"]
    want = O.log_mel_spectrogram(x.numpy().astype(np.float32) / 32768.0, 128, 160)
    assert feats.shape == want.shape and np.abs(feats - want).max() <= 1e-4
    m = GpuLogMelFrontEnd.pad_or_trim(torch.zeros(80, 10), 3000)
    assert m.shape == (80, 3000)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float16", "float32"])
def test_encoder_windows_bit_exact(frontend, dtype):
    """The 30 s windows the encoder consumes (whisper/transcribe.py seek loop: mel[:, seek:seek+size] -> pad_or_trim ->
    .to(dtype)) from ONE launch, bit-exact against the torch restatement -- incl. odd seeks, short last windows, a
    window past the end, explicit segment sizes, and an odd window length."""
    from modular_audio_pipeline_b200.transcriber import GpuLogMelFrontEnd
    dt = getattr(torch, dtype)
    fe = GpuLogMelFrontEnd(frontend.device)
    fe._fe = frontend
    g = torch.Generator().manual_seed(5)
    mel = (torch.randn(80, 9000 + 137, generator=g) * 0.7).to(frontend.device)
    seeks = [0, 3000, 2999, 6001, 8999, 9137, 12000, 1]
    lens = [3000, 3000, 1234, 3000, 138, 3000, 3000, 0]
    for length, ln in ((3000, None), (3000, lens), (1501, None)):
        got = fe.windows(mel, seeks, ln, length=length, dtype=dt)
        assert got.shape == (len(seeks), 80, length) and got.dtype == dt
        for w, sk in enumerate(seeks):
            size = length if ln is None else min(ln[w], length)
            seg = mel[:, sk:sk + size]
            want = GpuLogMelFrontEnd.pad_or_trim(seg, length).to(dt)
            assert torch.equal(got[w], want), (dtype, length, w)
    assert fe.windows(mel, [], None).shape == (0, 80, 3000)


@pytest.mark.gpu
def test_process_files_equals_chained_protocol_calls(tmp_path, frontend):
    """The fused file-batch entry leaves the same WAV the three chained Protocol calls leave (bit for bit), for ragged lengths
    and more files than one batch holds; the log-mel equals log_mel of that file."""
    from modular_audio_pipeline_b200 import GpuAudioPreprocessor
    from modular_audio_pipeline_b200.transcriber import GpuLogMelFrontEnd
    secs = [4.0, 7.31, 2.5, 11.0, 3.003]
    paths = []
    for i, s in enumerate(secs):
        p = tmp_path / f"in{i}.wav"
        _write_wav(p, speech_noise(s, 20 + i).numpy())
        paths.append(str(p))
    chained = GpuAudioPreprocessor(SR, str(tmp_path / "a"), device=frontend.device)
    fused = GpuAudioPreprocessor(SR, str(tmp_path / "b"), device=frontend.device)
    chained._fe = fused._fe = frontend
    lm = GpuLogMelFrontEnd(frontend.device); lm._fe = frontend
    got = list(fused.process_files(paths, batch_size=2, n_mels=80, mel_padding=160))
    assert len(got) == len(paths)
    for p, (out, lufs, mel) in zip(paths, got):
        want_path = chained.normalize_loudness(chained.normalize_audio(chained.reduce_stationary_noise(p)))
        assert os.path.basename(out) == os.path.basename(want_path)
        w, _, _ = _read_wav(want_path)
        g, sr, ch = _read_wav(out)
        assert sr == SR and ch == 1 and np.array_equal(g, w)
        assert np.isfinite(lufs)
        want_mel = lm.log_mel_spectrogram(torch.from_numpy(g.copy()), 80, 160)
        assert torch.equal(mel, want_mel)
    # no files written, no mel
    res = list(fused.process_files(paths[:1], n_mels=0, write_output=False))
    assert res[0][0] is None and res[0][2] is None

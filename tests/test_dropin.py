"""The reference-facing surface: PreprocessorProtocol / TranscriberProtocol conformance, signatures equal to the
reference class, and (GPU) the path-in / path-out methods in the order AudioPipeline.run() calls them
(pipeline.py:331-357, :372) checked against the oracle."""
import importlib.util
import inspect
import os
from pathlib import Path
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


class _FwExtractor:
    """The attribute surface of faster-whisper 1.2.1 ``feature_extractor.FeatureExtractor`` that ``WhisperModel.transcribe`` /
    ``generate_segments`` read off ``model.feature_extractor`` (recalled from its published source; the package is not
    installable here)."""

    def __init__(self, n_mels=128):
        self.n_fft, self.hop_length, self.chunk_length, self.sampling_rate = 400, 160, 30, 16000
        self.n_samples = self.chunk_length * self.sampling_rate
        self.nb_max_frames = self.n_samples // self.hop_length
        self.time_per_frame = self.hop_length / self.sampling_rate
        self.mel_filters = np.zeros((n_mels, 201), dtype=np.float32)

    def __call__(self, waveform, padding=160, chunk_length=None):
        raise AssertionError("the CPU extractor must not run once the GPU front-end is installed")


class _FwModel:
    """``WhisperModel.transcribe`` as far as the feature extractor is concerned: reads ``sampling_rate`` first, extracts features
    of the WHOLE recording with ``chunk_length`` forwarded, then walks windows of ``nb_max_frames`` with ``time_per_frame``."""

    def __init__(self, n_mels=128):
        self.feature_extractor = _FwExtractor(n_mels)

    def transcribe(self, audio, chunk_length=None):
        sr = self.feature_extractor.sampling_rate
        duration = audio.shape[0] / sr
        feats = self.feature_extractor(audio, chunk_length=chunk_length)
        nb = self.feature_extractor.nb_max_frames
        content_frames = feats.shape[-1] - 1
        windows = [(seek, min(nb, content_frames - seek)) for seek in range(0, content_frames, nb)]
        return feats, {"duration": duration, "windows": windows, "time_per_frame": self.feature_extractor.time_per_frame,
                       "n_samples": self.feature_extractor.n_samples}


def _adapter_roundtrip(front_end, x_f32, want_fn):
    from modular_audio_pipeline_b200.transcriber import GpuFeatureExtractor, GpuFrontEndTranscriber

    class Inner:
        def __init__(self): self._model = _FwModel(128)
        def is_loaded(self): return True
        def load_model(self): pass
        def unload_model(self): self.unloaded = True
        def transcribe(self, p):
            feats, info = self._model.transcribe(x_f32, chunk_length=20)
            return {"segments": [], "features": feats, "info": info}
    inner = Inner()
    original = inner._model.feature_extractor
    mod = types.ModuleType("whisper.transcribe")
    sentinel = mod.log_mel_spectrogram = lambda *a, **k: None
    saved = sys.modules.get("whisper.transcribe")
    sys.modules["whisper.transcribe"] = mod
    try:
        t = GpuFrontEndTranscriber(inner, front_end)
        out = t.transcribe("x.wav")
        assert isinstance(inner._model.feature_extractor, GpuFeatureExtractor)
        assert mod.log_mel_spectrogram == front_end.log_mel_spectrogram
        feats, info = out["features"], out["info"]
        # chunk_length re-sizes the bookkeeping of the WRAPPED extractor, it does not trim the waveform
        assert original.n_samples == 20 * 16000 and original.nb_max_frames == 2000 and info["n_samples"] == 320000
        assert info["duration"] == x_f32.shape[0] / 16000 and info["time_per_frame"] == 0.01
        want = want_fn(x_f32, 128, 160)
        assert feats.shape == want.shape and feats.shape[-1] == (x_f32.shape[0] + 160) // 160
        assert np.abs(feats - want).max() <= 1e-4
        assert info["windows"][0] == (0, 2000) and len(info["windows"]) == -(-(feats.shape[-1] - 1) // 2000)
        t.unload_model()                                  # pipeline.py:532-533
        assert inner._model.feature_extractor is original and mod.log_mel_spectrogram is sentinel and inner.unloaded
    finally:
        if saved is None:
            sys.modules.pop("whisper.transcribe", None)
        else:
            sys.modules["whisper.transcribe"] = saved


def test_transcriber_adapter_keeps_the_extractor_surface_cpu():
    """the adapter must not replace ``WhisperModel.feature_extractor`` by a bare callable -- transcribe() reads
    ``sampling_rate`` / ``nb_max_frames`` / ``time_per_frame`` off it.  Host logic only: the log-mel comes from the oracle."""
    class OracleFrontEnd:                                 # the two calls GpuFeatureExtractor makes on GpuLogMelFrontEnd
        def _to_int16(self, w): return torch.from_numpy(np.round(np.asarray(w, dtype=np.float32) * 32768.0).astype(np.int16))
        def log_mel_spectrogram(self, x, n_mels=80, padding=0, device=None):
            return torch.from_numpy(O.log_mel_spectrogram(x.numpy().astype(np.float32) / 32768.0, n_mels, padding))
    x = speech_noise(47.0, 1).numpy().astype(np.float32) / 32768.0
    _adapter_roundtrip(OracleFrontEnd(), x, O.log_mel_spectrogram)


@pytest.mark.gpu
def test_transcriber_adapter_with_fake_models(frontend):
    from modular_audio_pipeline_b200.transcriber import GpuLogMelFrontEnd
    fe = GpuLogMelFrontEnd(frontend.device)
    fe._fe = frontend
    x = speech_noise(47.0, 1).numpy().astype(np.float32) / 32768.0
    _adapter_roundtrip(fe, x, O.log_mel_spectrogram)
    m = GpuLogMelFrontEnd.pad_or_trim(torch.zeros(80, 10), 3000)
    assert m.shape == (80, 3000)
    # openai-whisper's call: log_mel_spectrogram(audio, n_mels, padding=N_SAMPLES) on a path or an array
    got = fe.log_mel_spectrogram(x, 80, 480000)
    assert got.is_cuda and np.abs(got.cpu().numpy() - O.log_mel_spectrogram(x, 80, 480000)).max() <= 1e-4


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


# ---- remove_silence / detect_silence_segments on the device (csrc/silence.cu) against the oracle's per-millisecond loop over
# ---- the real stdlib audioop ---------------------------------------------------------------------------------------------
def _gappy(seconds, seed, sr=SR, n_gaps=6, sigma=0.0002):
    """speech + faint noise with some stretches of near / exact digital silence of assorted lengths"""
    x = speech_noise(seconds, seed, sr=sr, sigma=sigma).numpy().copy()
    rng = np.random.default_rng(seed)
    for k in range(n_gaps):
        a = int(rng.integers(0, max(1, len(x) - sr)))
        b = a + int(rng.integers(sr // 50, sr))
        x[a:b] = 0 if k % 2 else (rng.integers(-2, 3, size=len(x[a:b]))).astype(np.int16)
    return x


@pytest.mark.gpu
@pytest.mark.parametrize("seconds,seed,sr,win,off,margin", [
    (20.0, 6, 16000, 250, 40.0, 100), (31.7, 7, 16000, 250, 40.0, 100), (12.3, 8, 16000, 100, 35.0, 30), (15.0, 9, 16000, 500, 40.0, 0),
    (9.0, 10, 16000, 50, 30.0, 10), (14.0, 11, 8000, 250, 40.0, 100), (11.0, 12, 48000, 250, 40.0, 100), (10.0, 13, 16000, 250, 40.0, 200),
    (6.0005, 14, 16000, 250, 40.0, 100), (7.00056, 15, 16000, 120, 38.0, 60)])
def test_remove_silence_device_path_bit_exact(frontend, seconds, seed, sr, win, off, margin):
    from modular_audio_pipeline_b200.silence import detect_silence_ms, remove_silence
    from modular_audio_pipeline_b200._compat import TimestampMapping
    pcm = _gappy(seconds, seed, sr=sr)
    x = torch.from_numpy(pcm).to(frontend.device)
    assert detect_silence_ms(frontend, x, sr, win, off) == O.detect_silence_ms(pcm, sr, win, off)
    want = O.remove_silence(pcm, sr, win, off, margin)
    got = remove_silence(frontend, x, sr, win, off, margin, True, TimestampMapping)
    assert (want is None) == (got is None)
    if want is not None:
        assert len(want[1]) >= 2, "the case should exercise the cross-fade"
        assert np.array_equal(got[0].cpu().numpy(), want[0])
        assert [(m.processed_start, m.processed_end, m.original_start, m.original_end) for m in got[1]] == want[1]


@pytest.mark.gpu
def test_remove_silence_edge_cases(frontend):
    from modular_audio_pipeline_b200.silence import detect_silence_ms, remove_silence
    from modular_audio_pipeline_b200._compat import TimestampMapping
    dev = frontend.device
    # all silence -> the reference returns its input (:321-323)
    z = torch.zeros(3 * SR, dtype=torch.int16, device=dev)
    assert remove_silence(frontend, z, SR, 250, 40.0, 100, True, TimestampMapping) is None
    assert detect_silence_ms(frontend, z, SR, 250, 40.0) == [[0, 3000]] == O.detect_silence_ms(np.zeros(3 * SR, dtype=np.int16), SR, 250, 40.0)
    # no silence at all -> one chunk, the whole recording, one mapping
    loud = speech_noise(4.0, 3, sigma=0.05).numpy()
    got = remove_silence(frontend, torch.from_numpy(loud).to(dev), SR, 250, 40.0, 100, True, TimestampMapping)
    want = O.remove_silence(loud, SR, 250, 40.0, 100)
    assert np.array_equal(got[0].cpu().numpy(), want[0]) and len(got[1]) == len(want[1]) == 1
    # shorter than the window: detect_silence returns [] -> everything is kept
    short = speech_noise(0.2, 4).numpy()
    got = remove_silence(frontend, torch.from_numpy(short).to(dev), SR, 250, 40.0, 100, False, TimestampMapping)
    assert np.array_equal(got[0].cpu().numpy(), short) and got[1] == []
    # silence at both ends, preserve_timestamps=False
    pcm = _gappy(8.0, 21)
    pcm[:SR] = 0; pcm[-SR:] = 0
    got = remove_silence(frontend, torch.from_numpy(pcm).to(dev), SR, 250, 40.0, 100, False, TimestampMapping)
    assert np.array_equal(got[0].cpu().numpy(), O.remove_silence(pcm, SR)[0]) and got[1] == []
    # margin 0 with tiny kept ranges: chunks shorter than the next cross-fade; pydub raises when the cross-fade is longer than
    # what has been accumulated -- both sides must agree on raise / no raise and on the samples
    rng = np.random.default_rng(5)
    for trial in range(6):
        y = np.zeros(4 * SR, dtype=np.int16)
        for k in range(8):
            a = int(rng.integers(0, len(y) - 4000)); ln = int(rng.integers(16, 1500))
            y[a:a + ln] = rng.integers(-9000, 9000, size=ln)
        try:
            want = O.remove_silence(y, SR, 40, 30.0, 0)
            err = None
        except ValueError as e:
            want, err = None, e
        if err is not None:
            with pytest.raises(ValueError):
                remove_silence(frontend, torch.from_numpy(y).to(dev), SR, 40, 30.0, 0, True, TimestampMapping)
        else:
            got = remove_silence(frontend, torch.from_numpy(y).to(dev), SR, 40, 30.0, 0, True, TimestampMapping)
            assert (want is None) == (got is None)
            if want is not None:
                assert np.array_equal(got[0].cpu().numpy(), want[0])
                assert [(m.processed_start, m.processed_end, m.original_start, m.original_end) for m in got[1]] == want[1]


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_remove_silence_one_hour_properties(frontend):
    """Full size (BASELINE config 2's recording): the oracle's per-millisecond Python loop needs minutes for an hour, so the device
    path is checked through properties -- the mappings tile the output, every un-faded output sample is the input sample its
    mapping names, and a one-minute excerpt equals the oracle bit for bit."""
    from modular_audio_pipeline_b200.silence import remove_silence
    from modular_audio_pipeline_b200._compat import TimestampMapping
    pcm = speech_noise(3600.0, 1, sigma=0.0002)
    x = pcm.to(frontend.device)
    out, maps = remove_silence(frontend, x, SR, 250, 40.0, 100, True, TimestampMapping)
    out = out.cpu().numpy()
    src = pcm.numpy()
    assert len(maps) > 1000 and out.size < src.size
    ms = lambda t: int(round(t * 1000))
    durs = [ms(m.processed_end) - ms(m.processed_start) for m in maps]
    cfs = [0] + [min(20, d // 4) for d in durs[1:]]                  # cross-fade INTO chunk i (preprocessor.py:349-354)
    for i, m in enumerate(maps):
        assert ms(m.original_end) - ms(m.original_start) == durs[i]
        nxt = cfs[i + 1] if i + 1 < len(maps) else 0
        if i + 1 < len(maps):                                       # the reference records a position BEFORE it subtracts that chunk's cross-fade
            assert ms(maps[i + 1].processed_start) == ms(m.processed_start) + durs[i] - cfs[i]
        a = ms(m.processed_start) * 16                              # chunk i sits at processed_start - cf_i; its un-faded body is a verbatim copy
        b = a + (durs[i] - cfs[i] - nxt) * 16
        o = (ms(m.original_start) + cfs[i]) * 16
        assert b > a and np.array_equal(out[a:b], src[o:o + (b - a)]), i
    assert out.size == (ms(maps[-1].processed_start) + durs[-1] - cfs[-1]) * 16
    one = src[: 60 * SR]
    got = remove_silence(frontend, x[: 60 * SR].contiguous(), SR, 250, 40.0, 100, True, TimestampMapping)
    want = O.remove_silence(one, SR)
    assert np.array_equal(got[0].cpu().numpy(), want[0])


@pytest.mark.gpu
@pytest.mark.parametrize("channels,sr,seconds", [(2, 16000, 12.0), (3, 16000, 5.0), (2, 48000, 6.0), (6, 8000, 9.0)])
def test_loudness_multichannel_parity(frontend, channels, sr, seconds):
    """normalize_loudness on interleaved multi-channel input (preprocessor.py:260-261: float32 channel mean)."""
    chans = [speech_noise(seconds, 50 + c, sr=sr).numpy() for c in range(channels)]
    inter = np.stack(chans, axis=1).reshape(-1)
    out, lufs = frontend.loudness_interleaved(torch.from_numpy(inter).to(frontend.device), channels, sr)
    want, want_lufs, applied = O.normalize_loudness(inter, sr, channels=channels)
    assert applied and abs(lufs.item() - want_lufs) <= 1e-4
    d = np.abs(out.cpu().numpy().astype(np.int32) - want.astype(np.int32))
    assert out.numel() == len(chans[0]) and d.max() <= 1 and np.mean(d != 0) <= 5e-3
    from modular_audio_pipeline_b200 import MafError
    with pytest.raises(MafError):
        frontend.loudness_interleaved(torch.zeros(2 * 1000, dtype=torch.int16, device=frontend.device), 2, sr)   # < 0.4 s


@pytest.mark.gpu
def test_protocol_path_reuses_plans_and_residents(tmp_path, frontend):
    """The Protocol chain on the real FrontEnd: one plan for the three stage calls of a recording, stage outputs handed over on the
    device (no file is re-read), deferred writes complete at remove_silence, prefetch() makes later calls start from HBM."""
    from modular_audio_pipeline_b200.preprocessor import GpuAudioPreprocessor
    paths = []
    for i in range(3):
        p = tmp_path / f"f{i}.wav"
        _write_wav(p, speech_noise(6.0 + i, 30 + i, sigma=0.001).numpy())
        paths.append(str(p))
    sync = GpuAudioPreprocessor(SR, str(tmp_path / "s"), device=frontend.device)
    lazy = GpuAudioPreprocessor(SR, str(tmp_path / "d"), device=frontend.device, write_mode="deferred")
    sync._fe = lazy._fe = frontend
    assert lazy.prefetch(paths) == 3 and all(lazy._lookup(p) is not None for p in paths)
    for p in paths:
        a = sync.remove_silence(sync.normalize_loudness(sync.normalize_audio(sync.reduce_stationary_noise(p))))
        b = lazy.remove_silence(lazy.normalize_loudness(lazy.normalize_audio(lazy.reduce_stationary_noise(p))))
        assert os.path.basename(a[0]) == os.path.basename(b[0]) and a[1] == b[1]
        for suffix in ("_denoised", "_denoised_norm", "_denoised_norm_loudnorm", "_denoised_norm_loudnorm_nosilence"):
            name = Path(p).stem + suffix + ".wav"
            assert (tmp_path / "s" / name).read_bytes() == (tmp_path / "d" / name).read_bytes()
    assert len(sync._plans) == 3                            # one stage plan per distinct length
    sync.close(); lazy.close()

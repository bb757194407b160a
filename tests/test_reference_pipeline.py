"""The reference's OWN orchestrator with the replacement injected:
``AudioPipeline(config, preprocessor=GpuAudioPreprocessor..., transcriber=Fake, vad=NoOp, ...).run("meeting.wav")``
(``/root/reference/audio_pipeline/pipeline.py:93-209`` constructor, ``:294-502`` run, ``:525-544`` cleanup).

The reference package is imported from ``/root/reference`` with empty stand-ins for its uninstallable numerical dependencies
(SURVEY.md section 4); it exists in the build container only, so these tests skip on the GPU box.  Without a CUDA device the
preprocessor's numerical back end is ``tests/fake_frontend.FakeFrontEnd`` (the oracle behind the FrontEnd / Plan surface): what
is under test here is everything the orchestrator can observe -- call order, arguments, output paths and files, the pass-through
and error contracts, timestamp remapping through the returned TimestampMappings -- plus the drop-in's own host machinery (plan /
resident caches, background writer, sync vs deferred writes).  With a GPU the same tests run on the real FrontEnd.
"""
import json
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


@pytest.fixture(scope="module")
def ref():
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
    import audio_pipeline
    from audio_pipeline import config, diarizer, pipeline, protocols, redundancy, separator, vad
    return types.SimpleNamespace(pkg=audio_pipeline, config=config, pipeline=pipeline, protocols=protocols, vad=vad, diarizer=diarizer,
                                 separator=separator, redundancy=redundancy)


def _preprocessor(temp_dir, noise_config=None, **kw):
    """GpuAudioPreprocessor on the real FrontEnd when there is a GPU, else on the oracle-backed test double."""
    from modular_audio_pipeline_b200.preprocessor import GpuAudioPreprocessor
    pre = GpuAudioPreprocessor(SR, str(temp_dir), noise_config=noise_config, **kw)
    if not torch.cuda.is_available():
        from fake_frontend import FakeFrontEnd
        pre._fe = FakeFrontEnd()
    return pre


class Recorder:
    """Wraps the preprocessor: records every Protocol call the orchestrator makes, in order."""

    def __init__(self, inner):
        self.inner, self.calls = inner, []

    def __getattr__(self, name):
        attr = getattr(self.inner, name)
        if name in ("reduce_stationary_noise", "normalize_audio", "normalize_loudness", "remove_silence", "detect_silence_segments"):
            def wrapped(*a, **k):
                out = attr(*a, **k)
                self.calls.append((name, a, k, out))
                return out
            return wrapped
        return attr


class FakeTranscriber:
    def __init__(self):
        self.paths, self.unloaded = [], False

    def transcribe(self, input_wav):
        self.paths.append(input_wav)
        assert os.path.isfile(input_wav), "the transcriber must find a complete file at the path it is given"
        n = _read_wav(input_wav)[0].size
        return {"segments": [{"start": 0.5, "end": 1.5, "text": "hello"}, {"start": 2.0, "end": min(3.0, n / SR), "text": "world"}]}

    def is_loaded(self):
        return True

    def load_model(self):
        pass

    def unload_model(self):
        self.unloaded = True


def _pipeline(ref, tmp_path, pre, **cfg_kw):
    media = tmp_path / "files"
    media.mkdir(exist_ok=True)
    cfg = ref.config.PipelineConfig(media_dir=str(media), **cfg_kw)
    cfg.vad.enabled = False
    cfg.diarization.enabled = False
    cfg.vocal_separation.enabled = False
    cfg.vocal_separation.auto_detect = False
    cfg.redundancy.enabled = False
    cfg.segment_merging.enabled = False
    tr = FakeTranscriber()
    pipe = ref.pipeline.AudioPipeline(cfg, preprocessor=pre, transcriber=tr, vad=ref.vad.NoOpVADFilter(),
                                      diarizer=ref.diarizer.NoOpDiarizer(), separator=ref.separator.NoOpVocalSeparator(),
                                      redundancy_remover=ref.redundancy.NoOpRedundancyRemover())
    return pipe, cfg, tr, media


@pytest.mark.parametrize("write_mode", ["sync", "deferred"])
def test_audio_pipeline_run_with_injected_preprocessor(ref, tmp_path, write_mode):
    pcm = speech_noise(9.0, 6, sigma=0.001).numpy()
    media = tmp_path / "files"
    media.mkdir()
    _write_wav(media / "meeting.wav", pcm)
    temp = media / "temp"
    inner = _preprocessor(temp, write_mode=write_mode)
    assert isinstance(inner, ref.protocols.PreprocessorProtocol)
    pre = Recorder(inner)
    pipe, cfg, tr, _ = _pipeline(ref, tmp_path, pre)
    assert pipe.preprocessor is pre
    result = pipe.run("meeting.wav")
    assert result.success, result.error
    # call order and arguments of pipeline.py:331-357
    names = [c[0] for c in pre.calls]
    assert names == ["reduce_stationary_noise", "normalize_audio", "normalize_loudness", "remove_silence"]
    src = str(media / "meeting.wav")
    p1, p2, p3 = (pre.calls[i][3] for i in range(3))
    p4, maps = pre.calls[3][3]
    assert pre.calls[0][1] == (src,) and pre.calls[1][1] == (p1,) and pre.calls[2][1] == (p2,) and pre.calls[3][1] == (p3,)
    assert pre.calls[3][2] == {"preserve_timestamps": True}
    assert [os.path.basename(p) for p in (p1, p2, p3, p4)] == ["meeting_denoised.wav", "meeting_denoised_norm.wav",
                                                               "meeting_denoised_norm_loudnorm.wav",
                                                               "meeting_denoised_norm_loudnorm_nosilence.wav"]
    assert all(os.path.dirname(p) == cfg.temp_dir for p in (p1, p2, p3, p4))
    assert tr.paths == [p4]
    if write_mode == "deferred":
        inner.flush()
    # the files are what the reference's chain leaves (oracle), stage by stage on the previous stage's file
    d1, d2, d3, d4 = (_read_wav(p)[0] for p in (p1, p2, p3, p4))
    _, want1, _ = O.denoise_int16(pcm, SR)
    diff = np.abs(d1.astype(np.int32) - want1.astype(np.int32))
    assert diff.max() <= 1 and np.mean(diff != 0) <= 1e-3
    assert np.array_equal(d2, O.peak_normalize(d1))
    want3, _, applied = O.normalize_loudness(d2, SR)
    diff = np.abs(d3.astype(np.int32) - want3.astype(np.int32))
    assert applied and diff.max() <= 1 and np.mean(diff != 0) <= 1e-3
    want4, want_maps = O.remove_silence(d3, SR)
    assert np.array_equal(d4, want4) and len(want_maps) > 1
    assert all(isinstance(m, ref.protocols.TimestampMapping) for m in maps)
    assert [(m.processed_start, m.processed_end, m.original_start, m.original_end) for m in maps] == want_maps
    # the orchestrator remapped the transcript through those mappings (pipeline.py:394-402) and wrote its JSON
    out = json.load(open(result.output_file))
    segs = out["segments"]
    assert [s["text"] for s in segs] == ["hello", "world"]
    for s in segs:
        assert s["original_start"] == pipe._map_timestamp_to_original(s["start"], maps)
        assert s["original_start"] >= s["start"]
    # the recording was uploaded once: every later stage started from the resident copy of the previous stage's output
    if not torch.cuda.is_available():
        assert inner._fe.plans_created == 1 and inner._fe.calls == ["silence"]
    pipe.cleanup()
    assert tr.unloaded and not os.path.exists(cfg.temp_dir)
    inner.close()


def test_audio_pipeline_noise_reduction_disabled_and_failure_is_reported(ref, tmp_path):
    pcm = speech_noise(6.0, 2, sigma=0.001).numpy()
    media = tmp_path / "files"
    media.mkdir()
    _write_wav(media / "a.wav", pcm)
    pre = Recorder(_preprocessor(media / "temp"))
    pipe, cfg, tr, _ = _pipeline(ref, tmp_path, pre)
    cfg.noise_reduction.enabled = False                      # pipeline.py:331-335: the denoise call is skipped altogether
    result = pipe.run("a.wav")
    assert result.success, result.error
    assert [c[0] for c in pre.calls] == ["normalize_audio", "normalize_loudness", "remove_silence"]
    assert os.path.basename(tr.paths[0]) == "a_norm_loudnorm_nosilence.wav"
    # a recording shorter than the loudness block: pyloudnorm's ValueError -> AudioProcessingError -> PipelineResult(success=False)
    _write_wav(media / "tiny.wav", pcm[:3000])
    pre2 = Recorder(_preprocessor(media / "temp2"))
    pipe2, cfg2, tr2, _ = _pipeline(ref, tmp_path, pre2)
    cfg2.noise_reduction.enabled = False
    r2 = pipe2.run("tiny.wav")
    assert not r2.success and "Loudness normalization failed" in r2.error and tr2.paths == []


def test_audio_pipeline_quiet_recording_passes_through_loudness(ref, tmp_path):
    """preprocessor.py:267-269: a too-quiet recording comes back as the INPUT path of normalize_loudness; remove_silence then finds
    no non-silent segment in digital silence?  No -- pydub reports the whole of an all-zero clip as one silent range, so the
    reference returns its input there too (:321-323) and the transcriber is handed the `_norm` file."""
    media = tmp_path / "files"
    media.mkdir()
    _write_wav(media / "quiet.wav", np.zeros(2 * SR, dtype=np.int16))
    from modular_audio_pipeline_b200._compat import NoiseReductionConfig
    pre = Recorder(_preprocessor(media / "temp", noise_config=NoiseReductionConfig(enabled=False)))
    pipe, cfg, tr, _ = _pipeline(ref, tmp_path, pre)
    cfg.noise_reduction.enabled = False
    result = pipe.run("quiet.wav")
    assert result.success, result.error
    norm = pre.calls[0][3]
    assert pre.calls[1][3] == norm                           # normalize_loudness returned its input
    assert pre.calls[2][3] == (norm, [])                     # remove_silence too
    assert tr.paths == [norm] and os.path.basename(norm) == "quiet_norm.wav"


def test_resident_copy_is_dropped_when_the_file_changes(ref, tmp_path):
    """Stage outputs are cached on the device by path; a foreign stage that rewrites the file (the vocal separator writes its own
    files, but nothing stops a custom one from overwriting) or a wiped temp_dir (media_handler.py:98-102) must win."""
    pcm = speech_noise(5.0, 3).numpy()
    _write_wav(tmp_path / "x.wav", pcm)
    pre = _preprocessor(tmp_path / "t")
    p1 = pre.reduce_stationary_noise(str(tmp_path / "x.wav"))
    assert pre._lookup(p1) is not None
    other = speech_noise(5.0, 4).numpy()
    _write_wav(p1, other)                                    # rewritten behind our back
    st = os.stat(p1)
    os.utime(p1, ns=(st.st_atime_ns, st.st_mtime_ns + 1_000_000))
    p2 = pre.normalize_audio(p1)
    assert np.array_equal(_read_wav(p2)[0], O.peak_normalize(other))
    os.remove(p2)
    with pytest.raises(Exception) as ei:                     # the file is gone: the reference would fail to open it
        pre.normalize_loudness(p2)
    assert type(ei.value).__name__ == "AudioProcessingError"
    pre.close()


def test_multichannel_and_other_rate_inputs(ref, tmp_path):
    """preprocessor.py:260-261 (normalize_loudness averages the channels) and :227-233 (normalize_audio resamples / downmixes)."""
    left, right = speech_noise(4.0, 21).numpy(), speech_noise(4.0, 22).numpy()
    inter = np.stack([left, right], axis=1).reshape(-1)
    _write_wav(tmp_path / "st.wav", inter, channels=2)
    pre = _preprocessor(tmp_path / "t")
    out, sr, ch = _read_wav(pre.normalize_loudness(str(tmp_path / "st.wav")))
    want, _, applied = O.normalize_loudness(inter, SR, channels=2)
    d = np.abs(out.astype(np.int32) - want.astype(np.int32))
    assert applied and (sr, ch) == (SR, 1) and out.size == left.size and d.max() <= 1 and np.mean(d != 0) <= 5e-3
    _write_wav(tmp_path / "st48.wav", np.stack([speech_noise(2.0, 5, sr=48000).numpy()] * 2, axis=1).reshape(-1), sr=48000, channels=2)
    out2, sr2, ch2 = _read_wav(pre.normalize_audio(str(tmp_path / "st48.wav")))
    assert (sr2, ch2) == (SR, 1) and out2.size == 2 * SR
    # 8-bit input is outside the path: the reference's error type, not a wrong answer
    with wave.open(str(tmp_path / "w8.wav"), "wb") as wf:
        wf.setnchannels(1); wf.setsampwidth(1); wf.setframerate(SR); wf.writeframes(bytes(8000))
    with pytest.raises(Exception) as ei:
        pre.normalize_audio(str(tmp_path / "w8.wav"))
    assert type(ei.value).__name__ == "AudioProcessingError" and "Audio normalization failed" in str(ei.value)
    pre.close()

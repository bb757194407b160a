"""CPU tests of the oracle (oracle/reference_path.py): golden fixtures, independent cross-checks, reference semantics."""
import os
import warnings

import numpy as np
import pytest
import torch

from modular_audio_pipeline_b200.synth import speech_noise
from oracle import reference_path as O

SR = 16000
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["clip_a", "clip_b"])
def test_oracle_reproduces_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    r = O.front_end(g["pcm"], SR, bool(g["stationary"]), int(g["n_mels"]), 0)
    assert np.array_equal(r["denoised_i16"], g["denoised_i16"])
    assert np.array_equal(r["norm_i16"], g["norm_i16"])
    assert np.array_equal(r["loud_i16"], g["loud_i16"])
    assert abs(r["lufs"] - float(g["lufs"])) < 1e-9
    assert np.abs(r["log_mel"] - g["log_mel"]).max() < 1e-6
    assert tuple(r["noise_clip"]) == tuple(g["noise_clip"])


@pytest.mark.parametrize("name", ["clip_a", "clip_b"])
def test_golden_independent_implementations(name):
    """The fixtures also hold outputs of implementations that share no code with the oracle."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    assert np.abs(g["hf_log_mel"] - g["oracle_log_mel_30s"]).max() <= 1e-6          # transformers.WhisperFeatureExtractor
    assert abs(float(g["ta_lufs"]) - float(g["quiet_lufs"])) <= 1e-3                 # torchaudio.functional.loudness


def test_frame_stats_vectorised_equals_reference_loop():
    pcm = speech_noise(20, 3).numpy().astype(np.float32)
    e1, z1 = O.frame_stats_loop(pcm, SR)
    e2, z2 = O.frame_stats(pcm, SR)
    assert np.array_equal(e1, e2) and np.array_equal(z1, z2)


def test_detect_noise_segments_matches_literal_loop():
    pcm = speech_noise(30, 5).numpy().astype(np.float32)
    # literal transcription of preprocessor.py:128-145
    e, z = O.frame_stats_loop(pcm, SR)
    nf = (e < np.percentile(e, 20)) & (z > np.percentile(z, 50) * 0.5)
    segs, in_noise, start = [], False, 0
    for i, v in enumerate(nf):
        if v and not in_noise:
            start, in_noise = i, True
        elif not v and in_noise:
            if (i - start) * 160 / SR >= 0.1:
                segs.append((start * 160, i * 160))
            in_noise = False
    assert segs == O.detect_noise_segments(pcm, SR)
    assert len(segs) > 0


def test_mel_filters_match_transformers():
    from transformers.audio_utils import mel_filter_bank
    for n_mels in (80, 128):
        hf = mel_filter_bank(201, n_mels, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T
        assert np.abs(O.mel_filters(n_mels) - hf.astype(np.float32)).max() < 1e-6


def test_log_mel_matches_whisper_feature_extractor():
    from transformers import WhisperFeatureExtractor
    x = speech_noise(30, 2).numpy().astype(np.float32) / 32768.0
    for n_mels in (80, 128):
        hf = WhisperFeatureExtractor(feature_size=n_mels)(x, sampling_rate=SR, return_tensors="np")["input_features"][0]
        assert np.abs(O.log_mel_spectrogram(x, n_mels, 0) - hf).max() <= 1e-6


def test_peak_normalize_matches_audioop():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        audioop = pytest.importorskip("audioop")
    pcm = speech_noise(3, 1).numpy()
    peak = audioop.max(pcm.tobytes(), 2)
    want = np.frombuffer(audioop.mul(pcm.tobytes(), 2, O.peak_normalize_factor(peak)), dtype=np.int16)
    assert np.array_equal(O.peak_normalize(pcm), want)
    assert audioop.max(np.array([-32768], dtype=np.int16).tobytes(), 2) == 32768


def test_loudness_matches_torchaudio_on_quiet_signal():
    import torchaudio
    x = speech_noise(20, 4).numpy().astype(np.float32) / 32768.0 * np.float32(0.5)
    ta = torchaudio.functional.loudness(torch.from_numpy(x)[None], SR).item()
    assert abs(O.integrated_loudness(x, SR) - ta) <= 1e-3


def test_k_weighting_coefficients_at_16k():
    (b1, a1), (b2, a2) = O.k_weighting_coeffs(16000)
    assert np.allclose(b1, [1.444233908713844, -1.839542987741787, 0.685432300820589], atol=1e-12)
    assert np.allclose(a1, [1.0, -1.1071392365035588, 0.3972624582962044], atol=1e-12)
    assert np.allclose(b2, [0.9852425301500196, -1.9704850603000392, 0.9852425301500196], atol=1e-12)
    assert np.allclose(a2, [1.0, -1.9703753578723906, 0.9705947627276879], atol=1e-12)


def test_chunk_geometry():
    assert O.chunk_table(960000) == [(0, 600000, 660000), (600000, 360000, 660000)]
    assert O.chunk_table(480000) == [(0, 480000, 540000)]
    assert O._stft(np.zeros(660000)).shape == (513, 2579)
    assert O._stft(np.zeros(540000)).shape == (513, 2110)
    vf, vt = O.smoothing_taps(16000)
    assert len(vf) == 33 and len(vt) == 7 and abs(np.outer(vf, vt).sum() - 68.0) < 1e-9


def test_loudness_glue_semantics():
    silent = np.zeros(SR, dtype=np.int16)
    out, lufs, applied = O.normalize_loudness(silent, SR)
    assert not applied and not np.isfinite(lufs) and np.array_equal(out, silent)
    with pytest.raises(ValueError):
        O.integrated_loudness(np.zeros(100, dtype=np.float32), SR)
    assert list(O.to_int16_trunc(np.array([1.7, -1.7, 40000.0], dtype=np.float32))) == [1, -1, -25536]


def test_front_door_restatement_matches_stdlib_audioop():
    """ratecv / tomono / the normalize_audio chain (preprocessor.py:227-233) against the REAL audioop pydub calls."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        audioop = pytest.importorskip("audioop")
    rng = np.random.default_rng(0)
    for ch in (1, 2):
        for a, b in [(48000, 16000), (44100, 16000), (22050, 16000), (8000, 16000), (32000, 16000), (11025, 16000)]:
            for n in (1, 2, 5, 1000, 48001):
                x = rng.integers(-32768, 32768, size=n * ch).astype(np.int16)
                want = np.frombuffer(audioop.ratecv(x.tobytes(), 2, ch, a, b, None)[0], dtype=np.int16)
                assert np.array_equal(want, O.ratecv(x, ch, a, b)), (ch, a, b, n)
    x = rng.integers(-20000, 20000, size=2 * 30011).astype(np.int16)
    assert np.array_equal(np.frombuffer(audioop.tomono(x.tobytes(), 2, 0.5, 0.5), dtype=np.int16), O.tomono(x, 2))
    # the whole chain as pydub runs it: max -> mul -> ratecv -> tomono
    peak = audioop.max(x.tobytes(), 2)
    f = O.peak_normalize_factor(peak)
    y = audioop.mul(x.tobytes(), 2, f)
    y = audioop.ratecv(y, 2, 2, 48000, 16000, None)[0]
    y = np.frombuffer(audioop.tomono(y, 2, 0.5, 0.5), dtype=np.int16)
    assert np.array_equal(y, O.front_door(x, 2, 48000, 16000))


def test_property_ratecv_closed_form_any_rate_pair():
    """hypothesis: the closed-form ratecv (output m reads input ceil(m in / out)) equals audioop.ratecv for arbitrary rate
    pairs -- up- and down-sampling, coprime rates -- lengths and channel counts; audioop.mul's floor for arbitrary factors."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        audioop = pytest.importorskip("audioop")
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(ch=st.sampled_from([1, 2]), a=st.integers(4000, 96000), b=st.sampled_from([8000, 16000, 22050, 44100, 48000, 7919]),
           n=st.integers(1, 3000), seed=st.integers(0, 2 ** 31 - 1))
    def check_ratecv(ch, a, b, n, seed):
        x = np.random.default_rng(seed).integers(-32768, 32768, size=n * ch).astype(np.int16)
        want = np.frombuffer(audioop.ratecv(x.tobytes(), 2, ch, a, b, None)[0], dtype=np.int16)
        assert np.array_equal(want, O.ratecv(x, ch, a, b))

    @settings(max_examples=150, deadline=None)
    @given(peak=st.integers(1, 32767), n=st.integers(1, 2000), seed=st.integers(0, 2 ** 31 - 1))
    def check_mul(peak, n, seed):
        x = np.random.default_rng(seed).integers(-peak, peak + 1, size=n).astype(np.int16)
        x[0] = peak
        want = np.frombuffer(audioop.mul(x.tobytes(), 2, O.peak_normalize_factor(audioop.max(x.tobytes(), 2))), dtype=np.int16)
        assert np.array_equal(want, O.peak_normalize(x))

    check_ratecv()
    check_mul()

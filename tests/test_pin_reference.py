"""PIN HOOK: the oracle against the REAL libraries the reference delegates to, whenever they can be imported.

The reference (``/root/reference``) computes nothing on this path itself: it calls noisereduce ~=3.0.3, pyloudnorm ~=0.1.1,
pydub ~=0.25.1 (over stdlib audioop), openai-whisper >=20250625 and faster-whisper ==1.2.1 (``environment.yml:39-49``).  None of
them is installable in the build container (no index, not in /opt/wheelhouse), so ``oracle/reference_path.py`` restates their
published algorithms and says "parity unpinned" where nothing independent backs it.  Every test here skips when its library is
missing and, when it is present -- on a developer machine, or if ``baseline/_ref`` ever holds an install (it is put on
``sys.path`` below) -- diffs the restatement against the real thing on the committed golden PCM, stage by stage, at the tolerances
a float64 re-implementation of the same formulas should meet.  ``tests/golden/make_golden.py --from-libraries`` regenerates the
golden vectors from the libraries in the same situation.  A failure here means the ORACLE is wrong, not the kernels.
"""
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_INSTALL = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF_INSTALL) and _REF_INSTALL not in sys.path:
    sys.path.append(_REF_INSTALL)

from oracle import reference_path as O  # noqa: E402

SR = 16000


def _real(name):
    """The real library or a skip: other tests register empty stand-in modules under these names (tests/test_dropin.py)."""
    mod = pytest.importorskip(name)
    if getattr(mod, "__file__", None) is None:
        pytest.skip(f"{name}: only a stand-in module is registered")
    return mod


def _golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return g["pcm"], g


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(x, dtype=np.float64) - ref
    return 10 * np.log10(np.sum(ref ** 2) / max(np.sum(err ** 2), 1e-300))


def test_pin_status_is_reported():
    """Always runs: states which of the reference's libraries this machine can pin against (none in the build container)."""
    have = {}
    for mod in ("noisereduce", "pyloudnorm", "pydub", "whisper", "faster_whisper", "librosa"):
        try:
            have[mod] = getattr(__import__(mod), "__file__", None) is not None
        except Exception:
            have[mod] = False
    print("pin status:", have)
    assert set(have) == {"noisereduce", "pyloudnorm", "pydub", "whisper", "faster_whisper", "librosa"}


@pytest.mark.parametrize("stationary", [True, False])
@pytest.mark.parametrize("name", ["clip_a", "clip_b"])
def test_noisereduce(name, stationary):
    nr = _real("noisereduce")
    pcm, _ = _golden(name)
    audio = pcm.astype(np.float32)                                        # preprocessor.py:171: unscaled int16 values
    s, e = O.choose_noise_clip(audio, SR)
    want = nr.reduce_noise(y=audio, sr=SR, y_noise=audio[s:e], prop_decrease=0.8, stationary=stationary)   # :200-206
    got = O.reduce_noise(audio, SR, audio[s:e], 0.8, stationary)
    assert want.shape == got.shape and snr_db(want, got) >= 120.0
    assert np.array_equal(O.to_int16_trunc(want), O.to_int16_trunc(got))
    # more than one 600 000-sample chunk: the chunk / padding geometry is where a restatement would go wrong by 100 dB
    long = np.tile(audio, 1 + 700000 // len(audio))[:700001]
    want = nr.reduce_noise(y=long, sr=SR, y_noise=audio[s:e], prop_decrease=0.8, stationary=stationary)
    assert snr_db(want, O.reduce_noise(long, SR, audio[s:e], 0.8, stationary)) >= 120.0


def test_noisereduce_short_noise_clip_8k():
    nr = _real("noisereduce")
    from modular_audio_pipeline_b200.synth import speech_noise
    audio = speech_noise(12.0, 31, sr=8000).numpy().astype(np.float32)
    noise = speech_noise(1.0, 77, sr=8000, sigma=0.02).numpy().astype(np.float32)[:880]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = nr.reduce_noise(y=audio, sr=8000, y_noise=noise, prop_decrease=0.8, stationary=True)
        got = O.reduce_noise(audio, 8000, noise, 0.8, True)
    assert snr_db(want, got) >= 120.0


@pytest.mark.parametrize("name", ["clip_a", "clip_b"])
def test_pyloudnorm(name):
    pyln = _real("pyloudnorm")
    _, g = _golden(name)
    samples = g["norm_i16"].astype(np.float32) / 32768.0                  # preprocessor.py:257
    for rate in (SR, 48000, 8000):
        want = pyln.Meter(rate).integrated_loudness(samples)              # :263-264
        assert abs(want - O.integrated_loudness(samples, rate)) <= 1e-9
    loud = pyln.Meter(SR).integrated_loudness(samples)
    normalized = pyln.normalize.loudness(samples, loud, -16.0)            # :271
    peak = np.abs(normalized).max()
    if peak > 1.0:
        normalized = normalized / peak
    want16 = np.clip(normalized * 32768, -32768, 32767).astype(np.int16)  # :279
    got16, _, applied = O.normalize_loudness(g["norm_i16"], SR)
    assert applied and np.array_equal(want16, got16)


@pytest.mark.parametrize("name", ["clip_a", "clip_b"])
def test_pydub_normalize_and_silence(name):
    pydub = _real("pydub")
    from pydub.effects import normalize
    from pydub.silence import detect_nonsilent, detect_silence
    _, g = _golden(name)
    den = g["denoised_i16"]
    seg = pydub.AudioSegment(data=den.tobytes(), sample_width=2, frame_rate=SR, channels=1)
    norm = np.frombuffer(normalize(seg)._data, dtype=np.int16)            # preprocessor.py:227-233
    assert np.array_equal(norm, O.peak_normalize(den))
    loud = g["loud_i16"]
    seg = pydub.AudioSegment(data=loud.tobytes(), sample_width=2, frame_rate=SR, channels=1)
    thr = seg.dBFS - 40.0
    assert detect_silence(seg, min_silence_len=250, silence_thresh=thr) == O.detect_silence_ms(loud, SR, 250, 40.0)
    assert detect_nonsilent(seg, min_silence_len=250, silence_thresh=thr) == O.detect_nonsilent_ms(loud, SR, 250, 40.0)
    # the reference's own loop, :330-358
    ranges = detect_nonsilent(seg, min_silence_len=250, silence_thresh=thr)
    cleaned, pos, maps = pydub.AudioSegment.empty(), 0, []
    for start_ms, end_ms in ranges:
        s, e = max(0, start_ms - 100), min(len(seg), end_ms + 100)
        chunk = seg[s:e]
        dur = len(chunk)
        maps.append((pos / 1000.0, (pos + dur) / 1000.0, s / 1000.0, e / 1000.0))
        if len(cleaned) == 0:
            cleaned = chunk
        else:
            cf = min(20, dur // 4)
            cleaned = cleaned.append(chunk, crossfade=cf)
            pos -= cf
        pos += dur
    want = O.remove_silence(loud, SR)
    if not ranges:
        assert want is None
    else:
        assert np.array_equal(np.frombuffer(cleaned._data, dtype=np.int16), want[0]) and maps == want[1]


def test_pydub_front_door():
    pydub = _real("pydub")
    from pydub.effects import normalize
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(2 * 48000 * 3) * 6000).clip(-32768, 32767).astype(np.int16)
    seg = pydub.AudioSegment(data=x.tobytes(), sample_width=2, frame_rate=48000, channels=2)
    out = normalize(seg).set_frame_rate(SR).set_channels(1).set_sample_width(2)      # preprocessor.py:227-233
    assert np.array_equal(np.frombuffer(out._data, dtype=np.int16), O.front_door(x, 2, 48000, SR))


@pytest.mark.parametrize("n_mels,padding", [(80, 480000), (128, 0), (80, 160)])
def test_whisper_log_mel(n_mels, padding):
    whisper = _real("whisper")
    import torch
    _, g = _golden("clip_a")
    audio = g["loud_i16"].astype(np.float32) / 32768.0
    want = whisper.audio.log_mel_spectrogram(torch.from_numpy(audio), n_mels, padding=padding).numpy()    # transcriber.py:229-238
    got = O.log_mel_spectrogram(audio, n_mels, padding)
    assert want.shape == got.shape and np.abs(want - got).max() <= 1e-6


def test_faster_whisper_feature_extractor():
    fw = _real("faster_whisper")
    from faster_whisper.feature_extractor import FeatureExtractor
    _, g = _golden("clip_b")
    audio = g["loud_i16"].astype(np.float32) / 32768.0
    fx = FeatureExtractor(feature_size=128)
    want = fx(audio, padding=160)                                         # transcriber.py:421-427
    got = O.log_mel_spectrogram(audio, 128, 160)
    assert want.shape == got.shape and np.abs(want - got).max() <= 1e-4
    assert (fx.sampling_rate, fx.hop_length, fx.nb_max_frames, fx.time_per_frame) == (16000, 160, 3000, 0.01)
    assert fw is not None

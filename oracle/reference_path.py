"""CPU oracle for the denoise -> peak-normalise -> BS.1770 loudness -> Whisper log-mel hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``modular_audio_pipeline_b200/`` imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg do,
and there only as the checker / CPU baseline.

PARITY UNPINNED.  The reference (``/root/reference``, AndersonGACFilho/Modular-Audio-Pipeline) delegates all
arithmetic on this path to pip dependencies that are neither vendored nor installable offline:
noisereduce ~=3.0.3, pyloudnorm ~=0.1.1, pydub ~=0.25.1 (over stdlib ``audioop``), openai-whisper >=20250625
(``environment.yml:39-49``), and it ships no tests, goldens or fixtures (``tests/__init__.py`` is empty).
This module therefore *restates* those libraries' published algorithms from the same numpy / scipy / torch
primitives they call, anchored on the reference's own call sites:

  S0  noise-clip auto-detect        audio_pipeline/preprocessor.py:92-147, :174-197   (reference's own code)
  S1-4 spectral-gate denoise        preprocessor.py:200-206 -> noisereduce.reduce_noise(stationary=True, prop_decrease=0.8)
  S5  peak normalise                preprocessor.py:227-233 -> pydub.effects.normalize over audioop.max / audioop.mul
  S6-7 loudness normalise           preprocessor.py:256-279 -> pyloudnorm.Meter.integrated_loudness / normalize.loudness
  S8-9 Whisper log-mel              transcriber.py:229-238  -> whisper.audio.log_mel_spectrogram

Independent cross-checks run in ``tests/test_oracle.py``: log-mel against ``transformers.WhisperFeatureExtractor``,
the mel matrix against ``transformers.audio_utils.mel_filter_bank``, loudness against
``torchaudio.functional.loudness``, peak normalise against stdlib ``audioop``.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import scipy.signal as sps

# --------------------------------------------------------------------------------------------------------------
# S0  noise-clip auto-detect -- preprocessor.py:92-147 (verbatim semantics) and clip choice :174-197
# --------------------------------------------------------------------------------------------------------------


def frame_stats_loop(audio: np.ndarray, sr: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-frame RMS energy and zero-crossing rate, literal loop of preprocessor.py:105-122."""
    frame_length = int(sr * 0.025)
    hop_length = int(sr * 0.010)
    num_frames = (len(audio) - frame_length) // hop_length + 1
    energies = np.zeros(max(num_frames, 0))
    zcrs = np.zeros(max(num_frames, 0))
    for i in range(num_frames):
        frame = audio[i * hop_length: i * hop_length + frame_length]
        energies[i] = np.sqrt(np.mean(frame ** 2))
        zcrs[i] = np.sum(np.abs(np.diff(np.signbit(frame)))) / frame_length
    return energies, zcrs


def frame_stats(audio: np.ndarray, sr: int, batch: int = 16384) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised equal of :func:`frame_stats_loop` (bit-identical; asserted in tests/test_oracle.py).

    ``np.mean`` over the contiguous last axis of a float32 matrix runs numpy's pairwise summation once per row
    with the same n=frame_length as the reference's per-frame call, so the float32 rounding is the same.
    """
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    frame_length = int(sr * 0.025)
    hop_length = int(sr * 0.010)
    num_frames = (len(audio) - frame_length) // hop_length + 1
    if num_frames <= 0:
        return np.zeros(0), np.zeros(0)
    win = np.lib.stride_tricks.sliding_window_view(audio, frame_length)[::hop_length][:num_frames]
    energies = np.zeros(num_frames)
    zcrs = np.zeros(num_frames)
    for s in range(0, num_frames, batch):
        fr = np.ascontiguousarray(win[s:s + batch])
        energies[s:s + batch] = np.sqrt(np.mean(fr ** 2, axis=1))
        sb = np.signbit(fr)
        zcrs[s:s + batch] = np.sum(sb[:, 1:] != sb[:, :-1], axis=1) / frame_length
    return energies, zcrs


def segments_from_stats(energies: np.ndarray, zcrs: np.ndarray, hop_length: int, sr: int) -> List[Tuple[int, int]]:
    """preprocessor.py:124-147 on precomputed per-frame statistics."""
    if len(energies) == 0:
        return []
    energy_threshold = np.percentile(energies, 20)
    zcr_threshold = np.percentile(zcrs, 50)
    noise_frames = (energies < energy_threshold) & (zcrs > zcr_threshold * 0.5)
    segs: List[Tuple[int, int]] = []
    # run-length decode of the :135-145 loop: a run counts only when a False frame closes it
    d = np.diff(np.concatenate(([0], noise_frames.astype(np.int8), [0])))
    starts = np.nonzero(d == 1)[0]
    ends = np.nonzero(d == -1)[0]
    for s, e in zip(starts, ends):
        if e >= len(noise_frames):           # still open at the end: dropped by the reference (no flush)
            continue
        if (e - s) * hop_length / sr >= 0.1:
            segs.append((int(s * hop_length), int(e * hop_length)))
    return segs


def detect_noise_segments(audio: np.ndarray, sr: int, fast: bool = True) -> List[Tuple[int, int]]:
    """preprocessor.py:92-147.  ``audio`` is float32 holding unscaled int16 values (:171)."""
    energies, zcrs = (frame_stats if fast else frame_stats_loop)(audio, sr)
    return segments_from_stats(energies, zcrs, int(sr * 0.010), sr)


def choose_noise_clip(audio: np.ndarray, sr: int, auto_detect: bool = True,
                      noise_sample_duration_s: float = 0.5) -> Tuple[int, int]:
    """Clip choice of preprocessor.py:179-197 as a (start, end) sample range."""
    if auto_detect:
        segs = detect_noise_segments(audio, sr)
        if segs:
            longest = max(segs, key=lambda x: x[1] - x[0])      # first maximum on ties, like :185
            return longest
    n = int(sr * noise_sample_duration_s)
    return (0, min(n, len(audio)))


# --------------------------------------------------------------------------------------------------------------
# S1-S4  spectral gating -- noisereduce 3.0.x SpectralGateStationary / SpectralGateNonStationary
# --------------------------------------------------------------------------------------------------------------

N_FFT = 1024
HOP = 256
CHUNK = 600000
PAD = 30000


def _stft(v: np.ndarray) -> np.ndarray:
    return sps.stft(v, nfft=N_FFT, nperseg=N_FFT, noverlap=N_FFT - HOP, padded=False)[2]


def _istft(Z: np.ndarray) -> np.ndarray:
    return sps.istft(Z, nfft=N_FFT, nperseg=N_FFT, noverlap=N_FFT - HOP)[1]


def amp_to_db(A: np.ndarray, top_db: float = 80.0) -> np.ndarray:
    """noisereduce ``_amp_to_db``: 20 log10(|x|+eps64), floored at (row max over time) - top_db."""
    D = 20.0 * np.log10(np.abs(A) + np.finfo(np.float64).eps)
    return np.maximum(D, np.max(D, axis=-1, keepdims=True) - top_db)


def smoothing_taps(sr: int, freq_mask_smooth_hz: float = 500.0, time_mask_smooth_ms: float = 50.0):
    """The two 1-D triangles whose normalised outer product is noisereduce's smoothing filter."""
    nf = int(freq_mask_smooth_hz / (sr / (N_FFT / 2)))
    nt = int(time_mask_smooth_ms / ((HOP / sr) * 1000))
    vf = np.concatenate([np.linspace(0, 1, nf + 1, endpoint=False), np.linspace(1, 0, nf + 2)])[1:-1]
    vt = np.concatenate([np.linspace(0, 1, nt + 1, endpoint=False), np.linspace(1, 0, nt + 2)])[1:-1]
    return vf, vt


def smoothing_filter(sr: int) -> np.ndarray:
    vf, vt = smoothing_taps(sr)
    K = np.outer(vf, vt)
    return K / np.sum(K)


def stationary_threshold(y_noise: np.ndarray, n_std: float = 1.5) -> np.ndarray:
    """SpectralGateStationary.__init__: per-bin mean + 1.5 * (population) std of the noise clip's dB STFT."""
    yn = np.asarray(y_noise, dtype=np.float64)[:CHUNK]
    Dn = amp_to_db(np.abs(_stft(yn)))
    return np.mean(Dn, axis=1) + n_std * np.std(Dn, axis=1)


def _filter_stationary(c: np.ndarray, thresh: np.ndarray, K: np.ndarray, prop: float) -> np.ndarray:
    S = _stft(c)
    M = (amp_to_db(np.abs(S)) > thresh[:, None]) * prop + (1.0 - prop)
    M = sps.fftconvolve(M, K, mode="same")
    d = _istft(S * M)
    out = np.zeros(len(c))
    out[:len(d)] = d
    return out


def nonstationary_b(sr: int, time_constant_s: float = 2.0) -> float:
    t = time_constant_s * sr / HOP
    return (math.sqrt(1 + 4 * t * t) - 1) / (2 * t * t)


def _filter_nonstationary(c: np.ndarray, K: np.ndarray, prop: float, sr: int,
                          thresh_n_mult: float = 2.0, sigmoid_slope: float = 10.0) -> np.ndarray:
    S = _stft(c)
    A = np.abs(S)
    b = nonstationary_b(sr)
    Sm = sps.filtfilt([b], [1.0, b - 1.0], A, axis=-1, padtype=None)
    with np.errstate(divide="ignore", invalid="ignore"):
        R = (A - Sm) / Sm
    M = 1.0 / (1.0 + np.exp(-(R - thresh_n_mult) * sigmoid_slope))
    M = sps.fftconvolve(M, K, mode="same")
    M = M * prop + (1.0 - prop)
    d = _istft(S * M)
    out = np.zeros(len(c))
    out[:len(d)] = d
    return out


def chunk_table(n: int) -> List[Tuple[int, int, int]]:
    """(start, kept_len, padded_len) for every denoise chunk of an n-sample recording (get_traces geometry)."""
    if n > CHUNK:
        return [(s, min(s + CHUNK, n) - s, CHUNK + 2 * PAD) for s in range(0, n, CHUNK)]
    return [(0, n, n + 2 * PAD)]


def reduce_noise(y: np.ndarray, sr: int, y_noise: Optional[np.ndarray] = None, prop_decrease: float = 0.8,
                 stationary: bool = True) -> np.ndarray:
    """``noisereduce.reduce_noise`` with the arguments of preprocessor.py:200-206; returns float32 (dtype of y)."""
    y = np.asarray(y)
    n = len(y)
    K = smoothing_filter(sr)
    thresh = stationary_threshold(y if y_noise is None else y_noise) if stationary else None
    result = np.zeros(n, dtype=np.float64)
    for (s, keep, L) in chunk_table(n):
        c = np.zeros(L, dtype=np.float64)
        lo = max(s - PAD, 0)
        hi = min(s + (L - 2 * PAD) + PAD, n)
        c[lo - (s - PAD): hi - (s - PAD)] = y[lo:hi]
        f = _filter_stationary(c, thresh, K, prop_decrease) if stationary else \
            _filter_nonstationary(c, K, prop_decrease, sr)
        result[s:s + keep] = f[PAD:PAD + keep]
    return result.astype(y.dtype if y.dtype.kind == "f" else np.float32)


def stationary_tie_samples(y: np.ndarray, sr: int, y_noise: np.ndarray, band_db: float = 1e-4) -> np.ndarray:
    """Boolean mask over the samples of y that a HARD-threshold tie can reach.  The stationary gate is `dB > thresh`, a step
    function: a cell whose dB value lies within `band_db` of its threshold is decided by the last bits of the FFT, which differ
    between any two implementations (scipy's float32 pocketfft in the real library, the float64 here, the float32 kernels).  A
    flipped cell moves the smoothed mask of its (2 nf + 1) x (2 nt + 1) neighbourhood, i.e. the output of the frames t - nt ..
    t + nt.  Test infrastructure for the parity tests: those samples are compared at a looser bound, everything else at the
    north-star 80 dB."""
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    thresh = stationary_threshold(y_noise)
    vf, vt = smoothing_taps(sr)
    nt = len(vt) // 2
    tied = np.zeros(n, dtype=bool)
    for (s, keep, L) in chunk_table(n):
        c = np.zeros(L)
        lo = max(s - PAD, 0)
        hi = min(s + (L - 2 * PAD) + PAD, n)
        c[lo - (s - PAD): hi - (s - PAD)] = y[lo:hi]
        D = amp_to_db(np.abs(_stft(c)))
        frames = np.nonzero((np.abs(D - thresh[:, None]) <= band_db).any(axis=0))[0]
        for t in frames:                                               # frame t is centred on padded sample t * HOP
            a = (t - nt) * HOP - N_FFT // 2 - PAD + s
            b = (t + nt) * HOP + N_FFT // 2 - PAD + s
            tied[max(a, s): max(min(b, s + keep), s)] = True
    return tied


def to_int16_trunc(x: np.ndarray) -> np.ndarray:
    """``reduced.astype(np.int16)`` of preprocessor.py:211: truncation toward zero, low 16 bits on overflow."""
    return np.asarray(x, dtype=np.float32).astype(np.int64).astype(np.int16)


def denoise_int16(pcm: np.ndarray, sr: int, stationary: bool = True, auto_detect: bool = True,
                  noise_sample_duration_s: float = 0.5, prop_decrease: float = 0.8):
    """reduce_stationary_noise, WAV-free: int16 in -> (float32 denoised, int16 as written at :211, clip range)."""
    audio = np.asarray(pcm, dtype=np.int16).astype(np.float32)          # :171 (unscaled)
    s, e = choose_noise_clip(audio, sr, auto_detect, noise_sample_duration_s)
    red = reduce_noise(audio, sr, audio[s:e], prop_decrease, stationary)
    return red, to_int16_trunc(red), (s, e)


# --------------------------------------------------------------------------------------------------------------
# S5  peak normalise -- pydub.effects.normalize(seg, headroom=0.1) over audioop.max / audioop.mul
# --------------------------------------------------------------------------------------------------------------


def peak_normalize_factor(peak: int, headroom: float = 0.1) -> float:
    target_peak = 32768.0 * (10 ** (-headroom / 20))
    needed_boost = 20 * math.log10(target_peak / peak)          # ratio_to_db
    return 10 ** (float(needed_boost) / 20)                      # db_to_float in apply_gain


def peak_normalize(pcm: np.ndarray, headroom: float = 0.1) -> np.ndarray:
    pcm = np.asarray(pcm, dtype=np.int16)
    peak = int(np.max(np.abs(pcm.astype(np.int32)))) if len(pcm) else 0     # audioop.max: 32768 for -32768
    if peak == 0:
        return pcm.copy()
    f = peak_normalize_factor(peak, headroom)
    v = np.floor(np.clip(pcm.astype(np.float64) * f, -32768.0, 32767.0))    # audioop.mul: floor(fbound(...))
    return v.astype(np.int16)


# --------------------------------------------------------------------------------------------------------------
# S6-S7  loudness -- pyloudnorm 0.1.1 Meter(rate) "K-weighting", block 0.400 s, + glue preprocessor.py:256-279
# --------------------------------------------------------------------------------------------------------------


def k_weighting_coeffs(rate: float):
    """pyloudnorm IIRfilter designs (RBJ forms evaluated at ``rate``): high shelf then high pass."""
    def design(G, Q, fc, kind):
        A = 10 ** (G / 40.0)
        w0 = 2.0 * np.pi * (fc / rate)
        alpha = np.sin(w0) / (2.0 * Q)
        if kind == "high_shelf":
            b0 = A * ((A + 1) + (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha)
            b1 = -2 * A * ((A - 1) + (A + 1) * np.cos(w0))
            b2 = A * ((A + 1) + (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha)
            a0 = (A + 1) - (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha
            a1 = 2 * ((A - 1) - (A + 1) * np.cos(w0))
            a2 = (A + 1) - (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha
        else:
            b0 = (1 + np.cos(w0)) / 2
            b1 = -(1 + np.cos(w0))
            b2 = (1 + np.cos(w0)) / 2
            a0 = 1 + alpha
            a1 = -2 * np.cos(w0)
            a2 = 1 - alpha
        return np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0
    return design(4.0, 1 / np.sqrt(2), 1500.0, "high_shelf"), design(0.0, 0.5, 38.0, "high_pass")


def block_energies(x: np.ndarray, rate: int) -> np.ndarray:
    """K-weight float32 mono samples and return the 400 ms / 75 % overlap block mean squares z[j]."""
    x = np.asarray(x, dtype=np.float32)
    if len(x) < 0.4 * rate:
        raise ValueError("Audio must have length greater than the block size.")
    d = x.copy()
    for (b, a) in k_weighting_coeffs(rate):
        d[:] = sps.lfilter(b, a, d)                    # double math, float32 storage between the stages
    T_g, step = 0.4, 0.25
    T = len(x) / rate
    nb = int(np.round(((T - T_g) / (T_g * step))) + 1)
    z = np.zeros(nb)
    for j in range(nb):
        l = int(T_g * (j * step) * rate)
        u = int(T_g * (j * step + 1) * rate)
        z[j] = (1.0 / (T_g * rate)) * float(np.sum(np.square(d[l:u])))
    return z


def gate_loudness(z: np.ndarray) -> float:
    """Absolute (-70 LUFS) then relative (-10 LU) gating of pyloudnorm's integrated_loudness (mono, G=1)."""
    with np.errstate(divide="ignore"):
        l = -0.691 + 10.0 * np.log10(z)
    J = l >= -70.0
    with np.errstate(divide="ignore", invalid="ignore"):
        z_avg = np.mean(z[J]) if np.any(J) else np.nan
        gamma_r = -0.691 + 10.0 * np.log10(z_avg) - 10.0
    J = (l > gamma_r) & (l > -70.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        z_avg = np.nan_to_num(np.mean(z[J])) if np.any(J) else 0.0
        return float(-0.691 + 10.0 * np.log10(z_avg))


def integrated_loudness(x: np.ndarray, rate: int) -> float:
    return gate_loudness(block_energies(x, rate))


def normalize_loudness(pcm: np.ndarray, rate: int, target_lufs: float = -16.0, channels: int = 1):
    """preprocessor.py:256-279, WAV-free.  Returns (int16 out, LUFS, applied) -- ``applied`` False = pass-through.
    ``channels`` > 1: ``pcm`` is interleaved and the reference meters / writes the float32 channel mean (:260-261)."""
    samples = np.asarray(pcm, dtype=np.int16).astype(np.float32) / np.float32(32768.0)
    if channels > 1:
        samples = samples.reshape(-1, channels).mean(axis=1)
    loudness = integrated_loudness(samples, rate)
    if not np.isfinite(loudness) or loudness < -70:
        return np.asarray(pcm, dtype=np.int16).copy(), loudness, False
    gain = np.float32(10.0 ** ((target_lufs - loudness) / 20.0))   # numpy 1.26: f64 scalar * f32 array stays f32
    normalized = gain * samples
    peak = np.abs(normalized).max()
    if peak > 1.0:
        normalized = normalized / peak
    out = np.clip(normalized * np.float32(32768), -32768, 32767).astype(np.int16)
    return out, loudness, True


# --------------------------------------------------------------------------------------------------------------
# S8-S9  Whisper log-mel -- whisper/audio.py log_mel_spectrogram (N_FFT 400, HOP 160, Slaney mel, fmax 8 kHz)
# --------------------------------------------------------------------------------------------------------------

W_SR, W_NFFT, W_HOP = 16000, 400, 160


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filters(n_mels: int) -> np.ndarray:
    """``librosa.filters.mel(sr=16000, n_fft=400, n_mels)`` (what whisper ships in mel_filters.npz), float32."""
    fftfreqs = np.fft.rfftfreq(W_NFFT, 1.0 / W_SR)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(W_SR / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    W = np.zeros((n_mels, len(fftfreqs)))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        W[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    W *= enorm[:, None]
    return W.astype(np.float32)


def log_mel_unclamped(audio: np.ndarray, n_mels: int = 80, padding: int = 0) -> np.ndarray:
    """log10(max(mel, 1e-10)) before the global max-8 clamp (used for sharded parity checks)."""
    import torch
    a = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    if padding > 0:
        a = torch.nn.functional.pad(a, (0, padding))
    window = torch.hann_window(W_NFFT)
    stft = torch.stft(a, W_NFFT, W_HOP, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    filters = torch.from_numpy(mel_filters(n_mels))
    mel_spec = filters @ magnitudes
    return torch.clamp(mel_spec, min=1e-10).log10().numpy()


def log_mel_spectrogram(audio: np.ndarray, n_mels: int = 80, padding: int = 0) -> np.ndarray:
    """whisper.audio.log_mel_spectrogram; ``audio`` float32 in [-1, 1).  Returns float32 [n_mels, (N+padding)//160]."""
    log_spec = log_mel_unclamped(audio, n_mels, padding)
    log_spec = np.maximum(log_spec, log_spec.max() - np.float32(8.0))
    return ((log_spec + np.float32(4.0)) / np.float32(4.0)).astype(np.float32)


# --------------------------------------------------------------------------------------------------------------
# The whole hot path as the fused GPU entry computes it (pipeline.py:331-347 order, then the transcriber front-end)
# --------------------------------------------------------------------------------------------------------------


def front_end(pcm: np.ndarray, sr: int = 16000, stationary: bool = True, n_mels: int = 80, padding: int = 0,
              target_lufs: float = -16.0, auto_detect: bool = True):
    """denoise -> int16 -> peak normalise -> loudness normalise -> int16 -> log-mel.  Returns a dict of every stage."""
    red, den16, clip = denoise_int16(pcm, sr, stationary, auto_detect)
    norm16 = peak_normalize(den16)
    loud16, lufs, applied = normalize_loudness(norm16, sr, target_lufs)
    mel = log_mel_spectrogram(loud16.astype(np.float32) / np.float32(32768.0), n_mels, padding)
    return {"denoised_f32": red, "denoised_i16": den16, "noise_clip": clip, "norm_i16": norm16,
            "loud_i16": loud16, "lufs": lufs, "applied": applied, "log_mel": mel}



# ------------------------------------------------------------------------------------------------------------------
# Front door: what normalize_audio does to a WAV that is not yet 16-bit mono at the target
# rate -- preprocessor.py:227-233: normalize(seg).set_frame_rate(sr).set_channels(1).set_sample_width(2), i.e. pydub over
# stdlib audioop, in this order: audioop.max / audioop.mul on the interleaved samples, audioop.ratecv(data, 2, channels,
# in_rate, out_rate, None) (weights 1, 0), audioop.tomono(data, 2, 0.5, 0.5).  Restated from CPython's Modules/audioop.c:
#   ratecv  rates divided by their gcd; d starts at -outrate; every input frame adds outrate; while d >= 0 an output frame
#           (int)((prev * d + cur * (outrate - d)) / outrate) is emitted (double arithmetic on the samples scaled to 32 bit,
#           truncation toward zero, high 16 bits kept) and d -= inrate.  Closed form: output m reads input n = ceil(m in / out)
#           with d = n out - m in, so there are floor((N - 1) out / in) + 1 outputs.
#   tomono  floor(fbound(l * 0.5 + r * 0.5)).
# ------------------------------------------------------------------------------------------------------------------
def ratecv(pcm: np.ndarray, channels: int, in_rate: int, out_rate: int) -> np.ndarray:
    """audioop.ratecv(pcm.tobytes(), 2, channels, in_rate, out_rate, None)[0] for int16 interleaved input."""
    import math
    x = np.asarray(pcm, dtype=np.int16).reshape(-1, channels)
    g = math.gcd(in_rate, out_rate)
    inr, outr = in_rate // g, out_rate // g
    n_in = x.shape[0]
    if n_in == 0:
        return np.zeros(0, dtype=np.int16)
    n_out = ((n_in - 1) * outr) // inr + 1
    m = np.arange(n_out, dtype=np.int64)
    n = -((-m * inr) // outr)                                   # ceil(m inr / outr)
    d = (n * outr - m * inr).astype(np.float64)[:, None]
    cur = x[n].astype(np.int64) << 16
    prev = np.where((n > 0)[:, None], x[np.maximum(n - 1, 0)].astype(np.int64) << 16, 0)
    val = (prev.astype(np.float64) * d + cur.astype(np.float64) * (float(outr) - d)) / float(outr)
    out = np.trunc(val).astype(np.int64) >> 16                   # (int) cast, then the high 16 bits
    return out.astype(np.int16).reshape(-1)


def tomono(pcm: np.ndarray, channels: int) -> np.ndarray:
    """pydub set_channels(1): audioop.tomono(data, 2, 0.5, 0.5) for stereo (floor of the mean); mono passes through."""
    x = np.asarray(pcm, dtype=np.int16).reshape(-1, channels)
    if channels == 1:
        return x[:, 0].copy()
    if channels != 2:
        raise ValueError("only mono / stereo input is restated")
    v = x[:, 0].astype(np.float64) * 0.5 + x[:, 1].astype(np.float64) * 0.5
    return np.floor(np.clip(v, -32768.0, 32767.0)).astype(np.int16)


def front_door(pcm: np.ndarray, channels: int, in_rate: int, out_rate: int, normalize: bool = True) -> np.ndarray:
    """normalize_audio on arbitrary 16-bit input: peak normalise (interleaved), ratecv, tomono -> int16 mono at out_rate."""
    x = np.asarray(pcm, dtype=np.int16).reshape(-1)
    if normalize:
        x = peak_normalize(x)
    if in_rate != out_rate:
        x = ratecv(x, channels, in_rate, out_rate)
    return tomono(x, channels)


# --------------------------------------------------------------------------------------------------------------
# Next-row 1: remove_silence -- preprocessor.py:290-375 over pydub.silence / AudioSegment.append, restated on the
# REAL stdlib audioop primitives pydub calls (audioop.rms / mul / add), per-millisecond loop exactly like pydub.
# --------------------------------------------------------------------------------------------------------------


def _audioop():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import audioop
    return audioop


def _seg_len_ms(n: int, sr: int) -> int:
    """pydub ``len(seg)``: round(1000 * (frame_count / frame_rate)) -- Python's round-half-even on that float."""
    return int(round(1000 * (float(n) / sr)))


def _seg_slice(pcm: np.ndarray, sr: int, a_ms, b_ms) -> np.ndarray:
    """pydub ``AudioSegment.__getitem__(slice)``: both ends clipped to len(seg) ms, positions converted with
    int(ms * (frame_rate / 1000.0)); a slice that runs past the last frame (len(seg) rounds UP by a fraction of a millisecond)
    is padded with silence to the expected length, unless more than 2 ms of frames are missing (TooManyMissingFrames)."""
    r = sr / 1000.0
    n_ms = _seg_len_ms(len(pcm), sr)
    a_ms, b_ms = min(a_ms, n_ms), min(b_ms, n_ms)
    f0, f1 = int(a_ms * r), int(b_ms * r)
    data = pcm[f0:f1]
    missing = max(f1 - f0, 0) - len(data)
    if missing:
        if missing > 2 * r:
            raise ValueError("TooManyMissingFrames: you should never be filling in more than 2 ms with silence here")
        data = np.concatenate([data, np.zeros(missing, dtype=np.int16)])
    return data


def detect_silence_ms(pcm: np.ndarray, sr: int, min_silence_len: int = 250, offset_db: float = 40.0):
    """pydub.silence.detect_silence(seg, min_silence_len, seg.dBFS - offset_db, seek_step=1)."""
    audioop = _audioop()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    seg_len = _seg_len_ms(len(pcm), sr)
    if seg_len < min_silence_len:
        return []
    rms = audioop.rms(pcm.tobytes(), 2)
    dbfs = 20 * math.log10(rms / 32768.0) if rms > 0 else -float("inf")
    thresh = (10 ** ((dbfs - offset_db) / 20)) * 32768.0
    starts = []
    for i in range(0, seg_len - min_silence_len + 1):
        sl = _seg_slice(pcm, sr, i, i + min_silence_len)
        if audioop.rms(sl.tobytes(), 2) <= thresh:
            starts.append(i)
    if not starts:
        return []
    ranges = []
    prev_i = starts.pop(0)
    cur = prev_i
    for s in starts:
        if s != prev_i + 1 and s > prev_i + min_silence_len:
            ranges.append([cur, prev_i + min_silence_len])
            cur = s
        prev_i = s
    ranges.append([cur, prev_i + min_silence_len])
    return ranges


def detect_nonsilent_ms(pcm, sr, min_silence_len=250, offset_db=40.0):
    silent = detect_silence_ms(pcm, sr, min_silence_len, offset_db)
    len_seg = _seg_len_ms(len(pcm), sr)
    if not silent:
        return [[0, len_seg]]
    if silent[0][0] == 0 and silent[0][1] == len_seg:
        return []
    prev_end, out, end_i = 0, [], 0
    for start_i, end_i in silent:
        out.append([prev_end, start_i])
        prev_end = end_i
    if end_i != len_seg:
        out.append([prev_end, len_seg])
    if out[0] == [0, 0]:
        out.pop(0)
    return out


def _fade(chunk: np.ndarray, from_gain: float, to_gain: float) -> np.ndarray:
    """AudioSegment.fade(start=0, end=inf) for durations <= 100 ms: per-frame audioop.mul with a linear gain ramp."""
    audioop = _audioop()
    from_power = 10 ** (from_gain / 20)
    delta = 10 ** (to_gain / 20) - from_power
    n = len(chunk)
    step = delta / n
    out = np.empty(n, dtype=np.int16)
    for i in range(n):
        out[i] = np.frombuffer(audioop.mul(chunk[i:i + 1].tobytes(), 2, from_power + step * i), dtype=np.int16)[0]
    return out


def remove_silence(pcm: np.ndarray, sr: int, min_silence_len=250, offset_db=40.0, margin=100):
    """preprocessor.py:311-358 -> (int16 cleaned, [(processed_start, processed_end, original_start, original_end)]) or None."""
    audioop = _audioop()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    ranges = detect_nonsilent_ms(pcm, sr, min_silence_len, offset_db)
    if not ranges:
        return None
    n_ms = _seg_len_ms(len(pcm), sr)
    cleaned = np.zeros(0, dtype=np.int16)
    maps, pos = [], 0
    for start_ms, end_ms in ranges:
        s, e = max(0, start_ms - margin), min(n_ms, end_ms + margin)
        chunk = _seg_slice(pcm, sr, s, e)
        dur = _seg_len_ms(len(chunk), sr)
        maps.append((pos / 1000.0, (pos + dur) / 1000.0, s / 1000.0, e / 1000.0))
        if _seg_len_ms(len(cleaned), sr) == 0:              # `if len(cleaned) == 0` is a length in ms
            cleaned = chunk.copy()
        else:
            cf = min(20, dur // 4)
            ncf = int(cf * (sr / 1000.0))
            if cf > _seg_len_ms(len(cleaned), sr):          # AudioSegment.append -> preprocessor.py:374-375 AudioProcessingError
                raise ValueError(f"Crossfade is longer than the original AudioSegment ({cf}ms > {_seg_len_ms(len(cleaned), sr)}ms)")
            if ncf > 0:
                xf = _fade(cleaned[-ncf:], 0, -120)
                fin = _fade(chunk[:ncf], -120, 0)
                xf = np.frombuffer(audioop.add(xf.tobytes(), fin.tobytes(), 2), dtype=np.int16)
                cleaned = np.concatenate([cleaned[:-ncf], xf, chunk[ncf:]])
            else:
                cleaned = np.concatenate([cleaned, chunk])
            pos -= cf
        pos += dur
    return cleaned, maps

"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical stage inputs.

Tolerances are BASELINE.json's: log-mel max-abs <= 1e-4, integrated loudness within 0.01 LU, denoised waveform
SNR >= 80 dB vs the oracle's float output (and its int16 form <= 1 LSB on <= 0.1 % of samples); integer stages
(noise-clip indices, peak normalise) bit-exact.
"""
import numpy as np
import pytest
import torch

from modular_audio_pipeline_b200.synth import speech_noise
from oracle import reference_path as O

pytestmark = pytest.mark.gpu
SR = 16000


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(x, dtype=np.float64) - ref
    return 10 * np.log10(np.sum(ref ** 2) / max(np.sum(err ** 2), 1e-300))


def gen(seconds, seed, **kw):
    return speech_noise(seconds, seed, **kw)


# ---- S0 ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seconds,seed", [(30, 3), (60, 0), (97.3, 5)])
def test_frame_stats_bit_exact(frontend, seconds, seed):
    pcm = gen(seconds, seed)
    plan = frontend.plan([pcm.numel()], n_mels=0)
    e, z = plan.frame_stats(plan.pack([pcm]))
    eo, zo = O.frame_stats(pcm.numpy().astype(np.float32), SR)
    assert np.array_equal(e.cpu().numpy().astype(np.float64), eo)
    assert np.array_equal(z.cpu().numpy() / 400.0, zo)


@pytest.mark.parametrize("seconds,seed", [(30, 3), (60, 0), (97.3, 5), (12, 9), (700.3, 6), (1800.0, 7), (3600.0, 1)])
def test_noise_clip_indices(frontend, seconds, seed):
    # recordings above 32768 frames (328 s) take the grid-wide co-operative selection kernel
    pcm = gen(seconds, seed)
    plan = frontend.plan([pcm.numel()], n_mels=0)
    _, _, clip = plan.denoise(plan.pack([pcm]))
    want = O.choose_noise_clip(pcm.numpy().astype(np.float32), SR)
    assert tuple(clip.cpu().numpy()[0]) == want


def test_noise_clip_fallback_when_nothing_detected(frontend):
    # a pure tone has no low-energy frames with high ZCR runs -> reference falls back to the first 0.5 s (:189-192)
    t = np.arange(5 * SR) / SR
    pcm = torch.from_numpy((8000 * np.sin(2 * np.pi * 440 * t)).astype(np.int16))
    plan = frontend.plan([pcm.numel()], n_mels=0)
    _, _, clip = plan.denoise(plan.pack([pcm]))
    want = O.choose_noise_clip(pcm.numpy().astype(np.float32), SR)
    assert tuple(clip.cpu().numpy()[0]) == want


# ---- S1-S4 --------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seconds,seed", [(20, 1), (37.5, 2), (60, 0), (100.2, 4)])
@pytest.mark.parametrize("stationary", [True, False])
def test_denoise_parity(frontend, seconds, seed, stationary):
    pcm = gen(seconds, seed, meeting=not stationary)
    plan = frontend.plan([pcm.numel()], stationary=stationary, n_mels=0)
    out16, outf, clip = plan.denoise(plan.pack([pcm]), want_float=True)
    red, den16, _ = O.denoise_int16(pcm.numpy(), SR, stationary=stationary)
    got = outf[0].cpu().numpy()
    s = snr_db(red, got)
    assert s >= 80.0, f"SNR {s:.1f} dB"
    g16 = out16[0].cpu().numpy().astype(np.int32)
    diff = np.abs(g16 - den16.astype(np.int32))
    assert diff.max() <= 1
    assert np.mean(diff != 0) <= 1e-3


@pytest.mark.parametrize("n", [600001, 599999, 4000, 1500, 1024])
@pytest.mark.parametrize("stationary", [True, False])
def test_denoise_parity_edge_lengths(frontend, n, stationary):
    """One sample either side of the 600 000-sample chunk boundary (a second chunk of ONE sample), and recordings of a few
    frames only (shorter than the S0 analysis needs: the leading-samples fallback clip)."""
    pcm = gen(38.0, 3, meeting=not stationary)[:n].contiguous()
    plan = frontend.plan([n], stationary=stationary, n_mels=0)
    out16, outf, clip = plan.denoise(plan.pack([pcm]), want_float=True)
    red, den16, ref_clip = O.denoise_int16(pcm.numpy(), SR, stationary=stationary)
    assert outf[0].numel() == n
    s = snr_db(red, outf[0].cpu().numpy())
    assert s >= 80.0, f"SNR {s:.1f} dB"
    diff = np.abs(out16[0].cpu().numpy().astype(np.int32) - den16.astype(np.int32))
    assert diff.max() <= 1 and np.mean(diff != 0) <= 2e-3
    if stationary:
        assert tuple(clip.cpu().numpy()[0]) == tuple(ref_clip)


def test_all_zero_recording_through_the_fused_entry(frontend):
    """Digital silence: the stationary gate returns zeros, the loudness stage passes through (-inf LUFS, :267-269), the
    log-mel is the constant the reference produces for zeros."""
    z = torch.zeros(2 * SR, dtype=torch.int16)
    plan = frontend.plan([z.numel()], stationary=True, n_mels=80)
    res = plan.process(plan.pack([z]))
    assert not np.isfinite(res.lufs.item())
    assert int(res.pcm[0].abs().max()) == 0
    want = O.log_mel_spectrogram(np.zeros(2 * SR, dtype=np.float32), 80, 0)
    assert np.abs(res.log_mel[0].cpu().numpy() - want).max() <= 1e-4


def test_denoise_fixed_clip_and_external_noise(frontend):
    pcm = gen(25, 7)
    x = pcm.numpy()
    # auto_detect off -> first 0.5 s (preprocessor.py:193-197)
    plan = frontend.plan([pcm.numel()], auto_detect_noise=False, n_mels=0)
    _, outf, clip = plan.denoise(plan.pack([pcm]), want_float=True)
    red, _, rng = O.denoise_int16(x, SR, auto_detect=False)
    assert tuple(clip.cpu().numpy()[0]) == rng == (0, 8000)
    assert snr_db(red, outf[0].cpu().numpy()) >= 80.0
    # explicit noise sample (noise_sample_path, :174-178)
    noise = gen(1.3, 99, sigma=0.02)
    plan2 = frontend.plan([pcm.numel()], external_noise=True, n_mels=0)
    plan2.set_noise_clips([0], [noise.numel()])
    _, outf2, _ = plan2.denoise(plan2.pack([pcm]), noise_pcm=noise.cuda(), want_float=True)
    red2 = O.reduce_noise(x.astype(np.float32), SR, noise.numpy().astype(np.float32), 0.8, True)
    assert snr_db(red2, outf2[0].cpu().numpy()) >= 80.0


def test_denoise_top_db_row_clamp(frontend):
    # a bin whose chunk maximum is > 80 dB above the noise threshold opens the whole row (amp_to_db top_db clamp)
    rng = np.random.default_rng(0)
    n = 20 * SR
    t = np.arange(n) / SR
    x = rng.normal(0, 1.2, n)
    x[5 * SR: 6 * SR] += 30000 * np.sin(2 * np.pi * 1000 * t[5 * SR: 6 * SR])
    pcm = torch.from_numpy(np.clip(np.round(x), -32768, 32767).astype(np.int16))
    plan = frontend.plan([n], auto_detect_noise=False, n_mels=0)
    _, outf, _ = plan.denoise(plan.pack([pcm]), want_float=True)
    red, _, _ = O.denoise_int16(pcm.numpy(), SR, auto_detect=False)
    assert snr_db(red, outf[0].cpu().numpy()) >= 80.0


# ---- S5 -----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1])
def test_peak_normalize_bit_exact(frontend, seed):
    pcm = gen(10, seed)
    plan = frontend.plan([pcm.numel()], n_mels=0)
    out = plan.peak_normalize(plan.pack([pcm]))[0].cpu().numpy()
    assert np.array_equal(out, O.peak_normalize(pcm.numpy()))
    try:                                                    # pydub.effects.normalize = audioop.max + audioop.mul (stdlib C code)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            import audioop
        b = pcm.numpy().tobytes()
        ref = np.frombuffer(audioop.mul(b, 2, O.peak_normalize_factor(audioop.max(b, 2))), dtype=np.int16)
        assert np.array_equal(out, ref)
    except ImportError:
        pass
    silent = torch.zeros(8000, dtype=torch.int16)
    plan = frontend.plan([8000], n_mels=0)
    assert np.array_equal(plan.peak_normalize(plan.pack([silent]))[0].cpu().numpy(), silent.numpy())


# ---- S6-S7 ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seconds,seed,scale", [(30, 0, 1.0), (60.05, 1, 0.2), (10, 2, 4.0), (0.45, 3, 1.0)])
def test_loudness_parity(frontend, seconds, seed, scale):
    pcm = gen(seconds, seed)
    pcm = torch.clamp((pcm.float() * scale).round(), -32768, 32767).to(torch.int16)
    plan = frontend.plan([pcm.numel()], n_mels=0)
    out, lufs = plan.loudness(plan.pack([pcm]))
    want16, want_lufs, applied = O.normalize_loudness(pcm.numpy(), SR)
    assert abs(lufs.item() - want_lufs) <= 0.01
    assert abs(lufs.item() - want_lufs) <= 1e-4           # what FP64 recurrences actually deliver
    d = np.abs(out[0].cpu().numpy().astype(np.int32) - want16.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) <= 1e-3


@pytest.mark.parametrize("seed,click", [(0, 32767), (1, 30011), (2, 23456), (3, 17777), (4, 29999)])
def test_limiter_division_fused_equals_staged_and_oracle(frontend, seed, click):
    """Limiter active (gain * peak > 1): the fused path divides with a per-thread reciprocal + per-sample correction
    (maf_common.cuh div_by), the standalone stage with __fdiv_rn, the oracle with numpy float32 division -- the int16
    results must be identical between the two CUDA paths and within the LSB-flip allowance of the oracle."""
    pcm = gen(20, seed)
    pcm = (pcm.float() * 0.15).round().to(torch.int16)
    pcm[SR * 3 + 7] = click; pcm[SR * 11 + 5] = -click          # clicks fix the peak, the speech fixes the loudness
    plan = frontend.plan([pcm.numel()], stationary=True, n_mels=80)
    res = plan.process(plan.pack([pcm]))
    den16, _, _ = plan.denoise(plan.pack([pcm]))
    norm16 = plan.peak_normalize(plan.pack([den16[0]]))
    staged, lufs = plan.loudness(plan.pack([norm16[0]]))
    norm_np = norm16[0].cpu().numpy()
    want16, want_lufs, applied = O.normalize_loudness(norm_np, SR)
    peak = np.abs(norm_np.astype(np.int32)).max() / 32768.0
    assert applied and 10 ** ((-16.0 - want_lufs) / 20) * peak > 1.05        # the limiter really is active
    assert torch.equal(res.pcm[0], staged[0])
    d = np.abs(staged[0].cpu().numpy().astype(np.int32) - want16.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) <= 1e-3


def test_loudness_silence_passthrough_and_too_short(frontend):
    from modular_audio_pipeline_b200 import MafError
    silent = torch.zeros(3 * SR, dtype=torch.int16)
    plan = frontend.plan([silent.numel()], n_mels=0)
    out, lufs = plan.loudness(plan.pack([silent]))
    assert not np.isfinite(lufs.item())                    # -inf, reference returns the input unchanged (:267-269)
    assert np.array_equal(out[0].cpu().numpy(), silent.numpy())
    short = gen(0.3, 0)
    plan = frontend.plan([short.numel()], n_mels=0)
    with pytest.raises(MafError):                          # pyloudnorm: "Audio must have length greater than the block size"
        plan.loudness(plan.pack([short]))


# ---- S8-S9 ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("seconds,padding", [(30, 0), (11.37, 160), (60, 480000), (1.0, 0)])
def test_logmel_parity(frontend, n_mels, seconds, padding):
    pcm = gen(seconds, 11)
    plan = frontend.plan([pcm.numel()], n_mels=n_mels, mel_padding=padding)
    mel = plan.log_mel(plan.pack([pcm]))[0].cpu().numpy()
    want = O.log_mel_spectrogram(pcm.numpy().astype(np.float32) / 32768.0, n_mels, padding)
    assert mel.shape == want.shape
    assert np.abs(mel - want).max() <= 1e-4


# ---- fused path -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("stationary,n_mels,padding", [(True, 80, 480000), (False, 128, 0)])
def test_fused_front_end(frontend, stationary, n_mels, padding):
    """The fused entry must equal the stage entry points chained on the GPU bit for bit (each stage is parity-tested
    against the oracle above), and stay close to the oracle's own chain."""
    pcm = gen(60, 0, meeting=not stationary)
    n = pcm.numel()
    plan = frontend.plan([n], stationary=stationary, n_mels=n_mels, mel_padding=padding)
    res = plan.process(plan.pack([pcm]))
    den16, _, clip = plan.denoise(plan.pack([pcm]))
    norm16 = plan.peak_normalize(plan.pack([den16[0]]))
    loud16, lufs = plan.loudness(plan.pack([norm16[0]]))
    mel = plan.log_mel(plan.pack([loud16[0]]))
    assert torch.equal(res.pcm[0], loud16[0])
    assert torch.equal(res.log_mel[0], mel[0])
    assert res.lufs.item() == lufs.item()
    # chained oracle: one flipped LSB after the denoise is amplified by the peak-normalise factor (~10x here), so the
    # end-to-end comparison is bounded by that amplification, not by 1 LSB ("Hard parts")
    ref = O.front_end(pcm.numpy(), SR, stationary, n_mels, padding)
    assert abs(res.lufs.item() - ref["lufs"]) <= 0.01
    amp = 32768.0 / max(1, np.abs(ref["denoised_i16"].astype(np.int32)).max())
    got16 = res.pcm[0].cpu().numpy()
    d = np.abs(got16.astype(np.int32) - ref["loud_i16"].astype(np.int32))
    assert d.max() <= 2 * amp + 2 and np.mean(d != 0) <= 2e-3
    want = O.log_mel_spectrogram(got16.astype(np.float32) / 32768.0, n_mels, padding)
    assert np.abs(res.log_mel[0].cpu().numpy() - want).max() <= 1e-4
    if stationary:
        assert tuple(res.noise_clip.cpu().numpy()[0]) == ref["noise_clip"]


@pytest.mark.parametrize("name", ["clip_a", "clip_b"])
def test_golden_fixtures(frontend, name):
    """Committed golden vectors (tests/golden/make_golden.py): oracle outputs plus the independent HF log-mel."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    pcm = torch.from_numpy(g["pcm"])
    stationary, n_mels = bool(g["stationary"]), int(g["n_mels"])
    plan = frontend.plan([pcm.numel()], stationary=stationary, n_mels=n_mels)
    den16, denf, clip = plan.denoise(plan.pack([pcm]), want_float=True)
    assert snr_db(g["denoised_f32"], denf[0].cpu().numpy()) >= 80.0
    if stationary:
        assert tuple(clip.cpu().numpy()[0]) == tuple(g["noise_clip"])
    norm = plan.peak_normalize(plan.pack([torch.from_numpy(g["denoised_i16"])]))[0].cpu().numpy()
    assert np.array_equal(norm, g["norm_i16"])
    loud, lufs = plan.loudness(plan.pack([torch.from_numpy(g["norm_i16"])]))
    assert abs(lufs.item() - float(g["lufs"])) <= 0.01
    d = np.abs(loud[0].cpu().numpy().astype(np.int32) - g["loud_i16"].astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) <= 1e-3
    mel = plan.log_mel(plan.pack([torch.from_numpy(g["loud_i16"])]))[0].cpu().numpy()
    assert np.abs(mel - g["log_mel"]).max() <= 1e-4
    # independent implementation: transformers.WhisperFeatureExtractor pads to 30 s before the STFT
    n = pcm.numel()
    plan30 = frontend.plan([n], n_mels=n_mels, mel_padding=480000 - n)
    mel30 = plan30.log_mel(plan30.pack([torch.from_numpy(g["loud_i16"])]))[0].cpu().numpy()
    assert np.abs(mel30 - g["hf_log_mel"]).max() <= 1e-4


def test_batch_equals_single(frontend):
    """Ragged batch: every item must come out bit-identical to running it alone (no cross-item leakage)."""
    clips = [gen(s, 20 + i) for i, s in enumerate([5.0, 30.0, 41.7, 8.25])]
    plan = frontend.plan([c.numel() for c in clips], n_mels=80)
    res = plan.process(plan.pack(clips))
    torch.cuda.synchronize()
    for i, c in enumerate(clips):
        p1 = frontend.plan([c.numel()], n_mels=80)
        r1 = p1.process(p1.pack([c]))
        assert torch.equal(res.pcm[i], r1.pcm[0])
        assert torch.equal(res.log_mel[i], r1.log_mel[0])
        assert res.lufs[i].item() == r1.lufs[0].item()


def test_process_host_matches_device(frontend):
    pcm = gen(20, 4)
    plan = frontend.plan([pcm.numel()], n_mels=80)
    res = plan.process(plan.pack([pcm]))
    hin = torch.zeros(plan.total_pcm, dtype=torch.int16).pin_memory()
    hin[:pcm.numel()] = pcm
    hout = torch.empty(plan.total_out, dtype=torch.int16).pin_memory()
    hmel = torch.empty(plan.total_mel, dtype=torch.float32).pin_memory()
    hl = torch.empty(1, dtype=torch.float64).pin_memory()
    plan.process_host(hin, hout, hmel, hl)
    assert torch.equal(plan.split_out(hout)[0], res.pcm[0].cpu())
    assert torch.equal(plan.split_mel(hmel)[0], res.log_mel[0].cpu())
    assert hl.item() == res.lufs.item()


def test_host_pipeline_overlap_matches_device(frontend):
    """HostPipeline (double-buffered H2D / kernels / D2H) must return exactly what the device entry returns."""
    from modular_audio_pipeline_b200 import HostPipeline
    clips = [gen(12, 30 + i) for i in range(3)]
    lengths = [c.numel() for c in clips]
    pipe = HostPipeline(frontend, lengths, depth=2, n_mels=80)
    want = []
    for rep in range(3):                                   # three submissions through two slots
        batch = [torch.roll(c, rep * 1000) for c in clips]
        slot = pipe.next_slot()
        pipe.plans[slot].pack(batch, out=pipe.h_in[slot])
        assert pipe.submit() == slot
        plan = frontend.plan(lengths, n_mels=80)
        want.append((slot, plan.process(plan.pack(batch))))
        if rep >= 1:                                       # check the previous submission while this one is in flight
            pslot, pres = want[rep - 1]
            out, mel, lufs = pipe.results(pslot)
            for i in range(3):
                assert torch.equal(out[i], pres.pcm[i].cpu()) and torch.equal(mel[i], pres.log_mel[i].cpu())
            assert torch.equal(lufs, pres.lufs.cpu())
    pipe.drain()


def test_host_pipeline_device_handoff(frontend):
    """mel_to_host=False: PCM and LUFS come back in pinned host memory, the log-mel stays a CUDA tensor."""
    from modular_audio_pipeline_b200 import HostPipeline
    clips = [gen(10, 40), gen(7.5, 41)]
    lengths = [c.numel() for c in clips]
    pipe = HostPipeline(frontend, lengths, depth=3, mel_to_host=False, n_mels=128, mel_padding=160)
    plan = frontend.plan(lengths, n_mels=128, mel_padding=160)
    want = plan.process(plan.pack(clips))
    slot = pipe.next_slot()
    pipe.plans[slot].pack(clips, out=pipe.h_in[slot])
    pipe.submit()
    out, mel, lufs = pipe.results(slot)
    assert pipe.d2h_bytes == 2 * plan.total_out + 8 * len(clips)
    for i in range(2):
        assert mel[i].is_cuda and torch.equal(mel[i], want.log_mel[i])
        assert not out[i].is_cuda and torch.equal(out[i], want.pcm[i].cpu())
    assert torch.equal(lufs, want.lufs.cpu())


# ---- other sample rates: the kernels' generic paths (time taps nt = 1 / 9, frequency taps nf = 32 / 5, loudness
# ---- segmentation for 800- / 4800-sample sub-blocks, serial frame statistics) ---------------------------------
@pytest.mark.parametrize("sr,seconds", [(8000, 40.0), (48000, 20.0), (22050, 15.0)])
@pytest.mark.parametrize("stationary", [True, False])
def test_denoise_parity_other_rates(frontend, sr, seconds, stationary):
    # 8 kHz stationary: the detected noise clip (960 samples here) is shorter than n_fft, so scipy shrinks nperseg to the clip
    # length and the hop to nperseg - 768; noise_pass_kernel follows (maf_common.cuh clip_nperseg)
    pcm = gen(seconds, 11, sr=sr, meeting=not stationary)
    plan = frontend.plan([pcm.numel()], sample_rate=sr, stationary=stationary, n_mels=0)
    out16, outf, clip = plan.denoise(plan.pack([pcm]), want_float=True)
    red, den16, seg = O.denoise_int16(pcm.numpy(), sr, stationary=stationary)
    if stationary:
        assert tuple(clip.cpu().numpy()[0]) == tuple(seg)
    got = outf[0].cpu().numpy()
    diff = np.abs(out16[0].cpu().numpy().astype(np.int32) - den16.astype(np.int32))
    if stationary:
        # the stationary gate is a hard `dB > thresh`: a cell within 5e-5 dB of its threshold (twice the measured
        # float32 error of the thresholds, 2e-5 dB) is decided by the last bits of the FFT (float64 in the oracle, float32 here
        # and in scipy's pocketfft); the frames such a tie can reach (a few % of the signal) are compared at 40 dB, every other
        # sample at the north-star 80 dB
        audio = pcm.numpy().astype(np.float32)
        tied = O.stationary_tie_samples(audio, sr, audio[seg[0]:seg[1]], band_db=5e-5)
        assert tied.mean() <= 0.05
        assert snr_db(red[~tied], got[~tied]) >= 80.0
        assert snr_db(red, got) >= 40.0
        diff = diff[~tied]
    else:
        s = snr_db(red, got)
        assert s >= 80.0, f"SNR {s:.1f} dB"
    assert diff.max() <= 1 and np.mean(diff != 0) <= 1e-3


@pytest.mark.parametrize("sr,seconds", [(8000, 40.0), (48000, 20.0), (32000, 12.3)])
def test_loudness_parity_other_rates(frontend, sr, seconds):
    pcm = gen(seconds, 12, sr=sr)
    plan = frontend.plan([pcm.numel()], sample_rate=sr, n_mels=0)
    out, lufs = plan.loudness(plan.pack([pcm]))
    want16, want_lufs, applied = O.normalize_loudness(pcm.numpy(), sr)
    assert abs(lufs.item() - want_lufs) <= 1e-4
    # a LUFS difference of 5e-6 LU is a gain difference of 6e-7: with samples of a few thousand LSB that flips the
    # truncation of ~0.2 % of them by one LSB
    d = np.abs(out[0].cpu().numpy().astype(np.int32) - want16.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) <= 5e-3


# ---- front door of normalize_audio: peak normalise -> ratecv -> tomono, bit-exact vs the oracle -----
@pytest.mark.parametrize("channels", [1, 2])
@pytest.mark.parametrize("in_rate,out_rate", [(48000, 16000), (44100, 16000), (22050, 16000), (8000, 16000), (16000, 16000)])
def test_front_door_bit_exact(frontend, channels, in_rate, out_rate):
    rng = np.random.default_rng(channels * 100003 + in_rate)
    n = 3 * in_rate + 17
    x = (rng.standard_normal(n * channels) * 6000).clip(-32768, 32767).astype(np.int16)
    want = O.front_door(x, channels, in_rate, out_rate)
    got = frontend.front_door(torch.from_numpy(x).to(frontend.device), channels, in_rate, out_rate)
    assert got.shape[0] == want.shape[0]
    assert np.array_equal(got.cpu().numpy(), want)
    raw = frontend.front_door(torch.from_numpy(x).to(frontend.device), channels, in_rate, out_rate, normalize=False)
    assert np.array_equal(raw.cpu().numpy(), O.front_door(x, channels, in_rate, out_rate, normalize=False))
    # and directly against the C implementation pydub calls (stdlib audioop, still shipped with Python 3.12): max / mul / ratecv /
    # tomono in pydub's order -- this stage's parity IS pinned to the reference's own arithmetic
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            import audioop
    except ImportError:
        return
    b = x.tobytes()
    peak = audioop.max(b, 2)
    b = audioop.mul(b, 2, O.peak_normalize_factor(peak)) if peak else b
    if in_rate != out_rate:
        b, _ = audioop.ratecv(b, 2, channels, in_rate, out_rate, None)
    if channels == 2:
        b = audioop.tomono(b, 2, 0.5, 0.5)
    assert np.array_equal(got.cpu().numpy(), np.frombuffer(b, dtype=np.int16))


# ---- full size: BASELINE config 2 (1 h, non-stationary, 128-mel, 30 s pad), the workload bench.py times ----------------
@pytest.mark.timeout(900)
def test_full_size_cfg2_against_oracle_and_properties(frontend):
    """The headline workload itself: all 96 chunks / 7776 tiles of a 1 h recording.  Stage-wise parity of the denoise and
    the loudness against the oracle (which needs ~20 s of CPU for the hour), the log-mel against the oracle on the GPU's own
    final PCM, plus size-independent properties: halving the input lowers the loudness by 20 log10(2) LU, and the
    fused entry equals the stages chained."""
    pcm = gen(3600.0, 1, meeting=True)
    n = pcm.numel()
    plan = frontend.plan([n], stationary=False, n_mels=128, mel_padding=480000)
    packed = plan.pack([pcm])
    out16, outf, _ = plan.denoise(packed, want_float=True)
    red, den16, _ = O.denoise_int16(pcm.numpy(), SR, stationary=False)
    got = outf[0].cpu().numpy()
    diff = np.abs(out16[0].cpu().numpy().astype(np.int32) - den16.astype(np.int32))
    s = snr_db(red, got)
    assert s >= 80.0, f"SNR {s:.1f} dB"
    assert diff.max() <= 1 and np.mean(diff != 0) <= 1e-3
    # fused entry on the same input
    res = plan.process(packed)
    final = res.pcm[0].cpu().numpy()
    # loudness of the GPU's own peak-normalised signal vs the oracle on the same samples
    norm = O.peak_normalize(out16[0].cpu().numpy())
    want16, want_lufs, applied = O.normalize_loudness(norm, SR)
    assert applied and abs(res.lufs.item() - want_lufs) <= 1e-4
    d = np.abs(final.astype(np.int32) - want16.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) <= 5e-3
    # log-mel of the final PCM
    want_mel = O.log_mel_spectrogram(final.astype(np.float32) / 32768.0, 128, 480000)
    assert res.log_mel[0].shape == want_mel.shape == (128, (n + 480000) // 160)
    assert np.abs(res.log_mel[0].cpu().numpy() - want_mel).max() <= 1e-4
    # property: half the amplitude -> loudness lower by 20 log10(2) (up to the int16 rounding of the halved samples)
    half = (pcm.to(torch.int32) // 2).to(torch.int16)
    p2 = frontend.plan([n], n_mels=0)
    _, l1 = p2.loudness(p2.pack([pcm]))
    _, l2 = p2.loudness(p2.pack([half]))
    assert abs((l1.item() - l2.item()) - 20 * np.log10(2.0)) <= 2e-3


@pytest.mark.timeout(900)
def test_full_size_cfg3_batch_equals_singles(frontend):
    """BASELINE config 3 at full size: 1024 x 30 s clips in ONE plan (stationary, 80-mel).  Every item of the batch must
    equal, bit for bit, the same clip processed alone (items are independent: no statistic may leak between files), and a
    few of them are checked against the oracle."""
    base = [gen(30.0, 100 + i) for i in range(16)]
    clips = [base[i % 16] for i in range(1024)]
    lengths = [c.numel() for c in clips]
    plan = frontend.plan(lengths, stationary=True, n_mels=80, mel_padding=0)
    res = plan.process(plan.pack(clips))
    singles = []
    for i in range(16):
        p1 = frontend.plan([lengths[i]], stationary=True, n_mels=80, mel_padding=0)
        r1 = p1.process(p1.pack([base[i]]))
        singles.append((r1.pcm[0].clone(), r1.log_mel[0].clone(), r1.lufs.clone(), r1.noise_clip.clone()))
        p1.close()
    for i in range(1024):
        pcm1, mel1, lufs1, clip1 = singles[i % 16]
        assert torch.equal(res.pcm[i], pcm1) and torch.equal(res.log_mel[i], mel1), f"item {i}"
        assert res.lufs[i].item() == lufs1.item() and torch.equal(res.noise_clip[i], clip1[0])
    for i in (0, 5, 11):
        ref = O.front_end(base[i].numpy(), SR, True, 80, 0)
        assert tuple(res.noise_clip[i].cpu().numpy()) == tuple(ref["noise_clip"])
        assert abs(res.lufs[i].item() - ref["lufs"]) <= 0.01


# ---- noise clips shorter than n_fft (config.py:198 allows 8 kHz: 100 ms = 800 samples) --------------------------------------
@pytest.mark.parametrize("clip_len", [769, 800, 880, 1001, 1023, 1024, 1025])
def test_short_noise_clip_follows_scipy_shrunken_nperseg(frontend, clip_len):
    """scipy.signal.stft(clip, nperseg=1024, noverlap=768) with len(clip) < 1024 uses nperseg = len(clip), hop = nperseg - 768
    and an nperseg-point periodic Hann window (oracle: the real scipy call); the thresholds, hence the gate, must follow."""
    import warnings
    sr = 8000
    pcm = gen(12.0, 31, sr=sr)
    noise = gen(1.0, 77, sr=sr, sigma=0.02)[:clip_len].contiguous()
    plan = frontend.plan([pcm.numel()], sample_rate=sr, external_noise=True, n_mels=0)
    plan.set_noise_clips([0], [clip_len])
    _, outf, clip = plan.denoise(plan.pack([pcm]), noise_pcm=noise.to(frontend.device), want_float=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                    # scipy: "nperseg = 1024 is greater than input length"
        red = O.reduce_noise(pcm.numpy().astype(np.float32), sr, noise.numpy().astype(np.float32), 0.8, True)
    assert snr_db(red, outf[0].cpu().numpy()) >= 80.0
    assert tuple(clip.cpu().numpy()[0]) == (0, clip_len)   # an external clip reports its own range (never uninitialised memory)


def test_noise_clip_not_longer_than_noverlap_is_an_error(frontend):
    """scipy raises "noverlap must be less than nperseg" for clips of <= 768 samples; the reference wraps that into
    AudioProcessingError (preprocessor.py:207-208).  The plan refuses such clips instead of producing numbers."""
    from modular_audio_pipeline_b200 import MafError
    pcm = gen(5.0, 1)
    plan = frontend.plan([pcm.numel()], external_noise=True, n_mels=0)
    with pytest.raises(MafError) as ei:
        plan.set_noise_clips([0], [768])
    assert ei.value.code == -1 and "768" in ei.value.message
    def denoise_of(plan, n):
        return plan.denoise(plan.pack([pcm[:n].contiguous()]))
    with pytest.raises(MafError):                          # the fallback clip would be 0.04 s = 640 samples
        denoise_of(frontend.plan([pcm.numel()], noise_sample_duration_s=0.04, n_mels=0), pcm.numel())
    short = frontend.plan([700], n_mels=0)                 # a 700-sample recording: its only possible clip is itself
    with pytest.raises(MafError):
        denoise_of(short, 700)
    assert short.peak_normalize(short.pack([pcm[:700].contiguous()]))[0].numel() == 700    # the other stages of that plan work
    with pytest.raises(MafError) as ei:                    # below 8 kHz (config.py:198 allows 8000 upwards) the plan is refused outright
        frontend.plan([60000], sample_rate=6000, n_mels=0)
    assert ei.value.code == -3


# ---- non-stationary gate on an all-zero frequency row: documented deviation (silence instead of noisereduce's NaN) ---------
def test_nonstationary_all_zero_chunk_guard(frontend):
    """noisereduce divides by the smoothed magnitude: a chunk that is digital silence gives 0 / 0 = NaN for the WHOLE chunk
    (SURVEY.md A.3).  The kernels take the relative excess as 0 there -- mask sigmoid(-20), output exactly 0 -- and every
    other chunk must still match the oracle."""
    pcm = gen(100.0, 4, meeting=True)
    n = pcm.numel()
    pcm[600000 - 30000: 2 * 600000 + 30000] = 0            # chunk 1 and both its 30 000-sample paddings: every row of its STFT is zero
    plan = frontend.plan([n], stationary=False, n_mels=80)
    out16, outf, _ = plan.denoise(plan.pack([pcm]), want_float=True)
    got = outf[0].cpu().numpy()
    assert np.isfinite(got).all()
    assert not got[600000: 1200000].any() and not out16[0][600000: 1200000].any()
    with np.errstate(all="ignore"):
        red, _, _ = O.denoise_int16(pcm.numpy(), SR, stationary=False)
    assert np.isnan(red[600000: 1200000]).all()            # the library behaviour the guard replaces
    for a, b in ((0, 600000), (1200000, n)):
        assert snr_db(red[a:b], got[a:b]) >= 80.0
    res = plan.process(plan.pack([pcm]))                   # and the fused entry stays finite end to end
    assert np.isfinite(res.lufs.item()) and torch.isfinite(res.log_mel[0]).all()


# ---- error paths of the tensor-level entry points ----------------------------------------------------------------------------
def test_front_door_and_mel_windows_error_paths(frontend):
    from modular_audio_pipeline_b200 import MafError
    x = torch.zeros(4800, dtype=torch.int16, device=frontend.device)
    with pytest.raises(MafError) as ei:
        frontend.front_door(x.float(), 1, 48000, 16000)                       # wrong dtype
    assert ei.value.code == -1 and "int16" in ei.value.message
    with pytest.raises(MafError):
        frontend.front_door(x.cpu(), 1, 48000, 16000)                         # not on the device
    with pytest.raises(MafError) as ei:
        frontend.front_door(x, 3, 48000, 16000)                               # three channels
    assert ei.value.code == -3
    with pytest.raises(MafError):
        frontend.front_door(x[:4799], 2, 48000, 16000)                        # odd number of interleaved stereo samples
    with pytest.raises(MafError) as ei:
        frontend.front_door(x, 1, 0, 16000)                                   # the library's own argument check
    assert ei.value.code == -1 and "rate" in ei.value.message
    mel = torch.zeros(80, 100, device=frontend.device)
    seeks = torch.zeros(2, dtype=torch.int32, device=frontend.device)
    out = torch.empty(2, 80, 3000, dtype=torch.float16, device=frontend.device)
    with pytest.raises(MafError):
        frontend.mel_windows(mel, seeks, torch.zeros(3, dtype=torch.int32, device=frontend.device), 3000, out)   # lens != seeks
    with pytest.raises(MafError):
        frontend.mel_windows(mel.double(), seeks, None, 3000, out)
    with pytest.raises(MafError):
        frontend.mel_windows(mel, seeks.long(), None, 3000, out)
    with pytest.raises(MafError):
        frontend.mel_windows(mel, seeks, None, 3000, out[:1])                 # output of the wrong size
    with pytest.raises(MafError) as ei:
        frontend.mel_windows(mel, seeks, None, 0, out[:0].reshape(2, 80, 0))  # the library's own shape check
    assert ei.value.code == -1
    assert frontend.mel_windows(mel, seeks, None, 3000, out).shape == (2, 80, 3000)


def test_plan_rejects_oversized_batches(frontend):
    from modular_audio_pipeline_b200 import MafError
    with pytest.raises(MafError) as ei:
        frontend.plan([16000] * 65536, n_mels=0)
    assert ei.value.code == -3 and "65535" in ei.value.message


# ---- bounds check without compute-sanitizer (closed on the GPU pool): red zones between the workspace buffers ---------------
@pytest.mark.parametrize("stationary,n_mels,padding,lengths", [
    (True, 80, 0, [16000 * 7 + 13, 600000, 600001, 1234567]),          # ragged batch, single- and multi-chunk items
    (False, 128, 480000, [2400000 + 777]),                             # 5 chunks, 30 s Whisper pad
    (True, 0, 0, [48000, 30001]),                                      # no log-mel: the standalone gain kernel
])
def test_workspace_red_zones(frontend, monkeypatch, stationary, n_mels, padding, lengths):
    """MAF_DEBUG_REDZONE puts 4 KB of unused bytes after EVERY workspace buffer.  The workspace, and guard bands around the
    caller's input / output buffers, are filled with 0xFF before the step: afterwards every gap and guard band must be
    untouched (no out-of-bounds write), and PCM, log-mel and LUFS must equal the run on the packed layout bit for bit (no
    result depends on bytes outside a buffer: 0xFF reads back as NaN / -1, which would not go unnoticed)."""
    clips = [gen(n / SR, 40 + i, meeting=not stationary)[:n].contiguous() for i, n in enumerate(lengths)]
    kw = dict(stationary=stationary, n_mels=n_mels, mel_padding=padding)
    plain = frontend.plan(lengths, **kw)
    want = plain.process(plain.pack(clips))
    monkeypatch.setenv("MAF_DEBUG_REDZONE", "4096")
    plan = frontend.plan(lengths, **kw)
    monkeypatch.delenv("MAF_DEBUG_REDZONE")
    lay = plan.workspace_layout()
    assert len(lay) > 20 and plan.workspace_bytes >= plain.workspace_bytes + 4096 * (len(lay) - 1)
    G = 4096                                                           # guard band (bytes) on both sides of the caller's buffers
    def guarded(n_elems, dtype):
        item = torch.empty((), dtype=dtype).element_size()
        raw = torch.full((2 * G + max(n_elems, 8) * item,), 255, dtype=torch.uint8, device=frontend.device)
        return raw, raw[G: G + max(n_elems, 8) * item].view(dtype)
    ws_raw = torch.full((plan.workspace_bytes,), 255, dtype=torch.uint8, device=frontend.device)
    plan._ws = ws_raw
    in_raw, pcm = guarded(plan.total_pcm, torch.int16)
    pcm.zero_()
    plan.pack(clips, out=pcm)
    out_raw, out = guarded(plan.total_out, torch.int16)
    mel_raw, mel = guarded(plan.total_mel, torch.float32)
    got = plan.process(pcm, out_pcm=out, mel=mel if n_mels else None)
    torch.cuda.synchronize()
    # 1. results identical to the packed layout
    for a, b in zip(got.pcm, want.pcm):
        assert torch.equal(a, b)
    assert torch.equal(got.lufs, want.lufs)
    if n_mels:
        for a, b in zip(got.log_mel, want.log_mel):
            assert torch.equal(a, b)
    # 2. every gap between (and after) the workspace buffers still holds the fill pattern
    ends = sorted((o + sz) for o, sz in lay)
    starts = sorted(o for o, _ in lay) + [plan.workspace_bytes]
    touched = []
    for e in ends:
        nxt = min(s for s in starts if s >= e)
        if nxt > e and not bool((ws_raw[e:nxt] == 255).all()):
            touched.append((e, nxt))
    assert not touched, f"workspace red zones written: {touched[:4]}"
    # 3. and so do the guard bands around the caller's buffers
    for name, raw, n_bytes in (("pcm in", in_raw, max(plan.total_pcm, 8) * 2), ("pcm out", out_raw, max(plan.total_out, 8) * 2),
                               ("mel", mel_raw, max(plan.total_mel, 8) * 4)):
        assert bool((raw[:G] == 255).all()) and bool((raw[G + n_bytes:] == 255).all()), f"guard band of {name} written"


# ---- gating of a LONG recording: the co-operative grid kernel (>= 16 384 blocks) against numpy and against the one-CTA kernel ----
@pytest.mark.parametrize("n_tiles", [9000, 16387, 36000, 288003])
def test_gate_long_recordings(frontend, n_tiles):
    """pyloudnorm's two gates over the 400 ms blocks (= 4 consecutive 100 ms sub-blocks) of a recording of up to 8 h; the
    grid kernel that takes over at 16 384 blocks must give the LUFS of a float64 numpy evaluation (and the gain derived from it)."""
    import ctypes as C
    from modular_audio_pipeline_b200._capi import check
    rng = np.random.default_rng(n_tiles)
    sub = rng.random(n_tiles) * 1e-3 * (rng.random(n_tiles) > 0.3)          # 30 % silent sub-blocks: both gates bite
    sub[: n_tiles // 7] *= 1e-9                                             # a stretch below the absolute gate
    nb = n_tiles - 3
    inv_block = 1.0 / (0.4 * SR)
    z = inv_block * (sub[:-3] + sub[1:-2] + sub[2:-1] + sub[3:])
    zabs = 10.0 ** ((-70.0 + 0.691) / 10.0)
    keep = z >= zabs
    zrel = z[keep].mean() * 0.1
    keep2 = (z > zrel) & (z > zabs)
    want = -0.691 + 10.0 * np.log10(z[keep2].mean())
    d_sub = torch.from_numpy(sub).to(frontend.device)
    pk = torch.full((1,), 20000, dtype=torch.int32, device=frontend.device)
    gate = torch.zeros(32, dtype=torch.uint8, device=frontend.device)
    st = torch.cuda.current_stream().cuda_stream
    check(frontend._lib.maf_gate(frontend._h, d_sub.data_ptr(), n_tiles, nb, SR, C.c_float(-16.0), pk.data_ptr(), gate.data_ptr(), st))
    torch.cuda.synchronize()
    lufs = float(gate[:8].view(torch.float64).item())
    gain = float(gate[8:12].view(torch.float32).item())
    assert abs(lufs - want) <= 1e-9
    assert gain == np.float32(10.0 ** ((-16.0 - lufs) / 20.0))
    # twice the same input: bit-identical (fixed partition, fixed summation order)
    gate2 = torch.zeros(32, dtype=torch.uint8, device=frontend.device)
    check(frontend._lib.maf_gate(frontend._h, d_sub.data_ptr(), n_tiles, nb, SR, C.c_float(-16.0), pk.data_ptr(), gate2.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(gate, gate2)

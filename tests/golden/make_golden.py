"""Regenerate tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py), or -- where
the reference's libraries are installed -- from the libraries themselves (python tests/golden/make_golden.py --from-libraries).

The reference ships no golden vectors and its numerical dependencies (noisereduce, pyloudnorm, pydub, whisper) are
not installable offline, so these fixtures pin the ORACLE (oracle/reference_path.py), which restates those
libraries, plus two independent implementations available in this image:
  * transformers.WhisperFeatureExtractor  -> `hf_log_mel`   (independent of the oracle's torch.stft route)
  * torchaudio.functional.loudness        -> `ta_lufs`      (independent BS.1770 meter)
Inputs are 3 s synthetic speech+noise clips (modular_audio_pipeline_b200.synth.speech_noise).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from modular_audio_pipeline_b200.synth import speech_noise  # noqa: E402
from oracle import reference_path as O  # noqa: E402


def front_end_from_libraries(pcm, sr, stationary, n_mels):
    """The same stage chain through the REAL libraries the reference calls (preprocessor.py:171-211, :227-233, :256-279,
    transcriber.py:229-238) -- only possible where noisereduce / pyloudnorm / pydub / whisper are installed (they are not in
    the build container; tests/test_pin_reference.py is the matching check)."""
    import noisereduce as nr
    import pydub
    import pyloudnorm as pyln
    import whisper
    from pydub.effects import normalize
    audio = pcm.astype(np.float32)
    clip = O.choose_noise_clip(audio, sr)                                  # the reference's own code (:92-147, :179-197)
    red = nr.reduce_noise(y=audio, sr=sr, y_noise=audio[clip[0]:clip[1]], prop_decrease=0.8, stationary=stationary)
    den16 = red.astype(np.int16)
    seg = pydub.AudioSegment(data=den16.tobytes(), sample_width=2, frame_rate=sr, channels=1)
    norm16 = np.frombuffer(normalize(seg)._data, dtype=np.int16)
    samples = norm16.astype(np.float32) / 32768.0
    lufs = pyln.Meter(sr).integrated_loudness(samples)
    y = pyln.normalize.loudness(samples, lufs, -16.0)
    peak = np.abs(y).max()
    if peak > 1.0:
        y = y / peak
    loud16 = np.clip(y * 32768, -32768, 32767).astype(np.int16)
    mel = whisper.audio.log_mel_spectrogram(torch.from_numpy(loud16.astype(np.float32) / 32768.0), n_mels).numpy()
    return {"denoised_f32": red, "denoised_i16": den16, "noise_clip": clip, "norm_i16": norm16, "loud_i16": loud16, "lufs": lufs,
            "applied": True, "log_mel": mel}


def main():
    import torchaudio
    from transformers import WhisperFeatureExtractor
    here = os.path.dirname(os.path.abspath(__file__))
    from_libs = "--from-libraries" in sys.argv
    for name, seconds, seed, stationary, n_mels in [("clip_a", 3.0, 0, True, 80), ("clip_b", 3.0, 1, False, 128)]:
        pcm = speech_noise(seconds, seed, meeting=not stationary).numpy()
        r = front_end_from_libraries(pcm, 16000, stationary, n_mels) if from_libs else O.front_end(pcm, 16000, stationary, n_mels, 0)
        x = r["loud_i16"].astype(np.float32) / 32768.0
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        hf = fe(x, sampling_rate=16000, return_tensors="np", padding="max_length")["input_features"][0]
        padded = O.log_mel_spectrogram(x, n_mels, 480000 - len(x))           # HF pads to 30 s before the STFT
        # torchaudio's lfilter clamps to +-1, so the meter cross-check runs on a quarter-scale copy (K-weighted peak < 1)
        quiet = (r["norm_i16"].astype(np.float32) / 32768.0) * np.float32(0.25)
        ta = torchaudio.functional.loudness(torch.from_numpy(quiet)[None], 16000).item()
        quiet_lufs = O.integrated_loudness(quiet, 16000)
        np.savez_compressed(os.path.join(here, name + ".npz"), pcm=pcm, stationary=stationary, n_mels=n_mels,
                            denoised_i16=r["denoised_i16"], denoised_f32=r["denoised_f32"], noise_clip=np.array(r["noise_clip"]),
                            norm_i16=r["norm_i16"], loud_i16=r["loud_i16"], lufs=r["lufs"], log_mel=r["log_mel"],
                            hf_log_mel=hf.astype(np.float32), oracle_log_mel_30s=padded, ta_lufs=ta, quiet_lufs=quiet_lufs)
        print(name, "lufs", r["lufs"], "quiet", quiet_lufs, "ta", ta, "hf-vs-oracle", np.abs(hf - padded).max())


if __name__ == "__main__":
    main()

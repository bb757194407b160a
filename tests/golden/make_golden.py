"""Regenerate tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

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


def main():
    import torchaudio
    from transformers import WhisperFeatureExtractor
    here = os.path.dirname(os.path.abspath(__file__))
    for name, seconds, seed, stationary, n_mels in [("clip_a", 3.0, 0, True, 80), ("clip_b", 3.0, 1, False, 128)]:
        pcm = speech_noise(seconds, seed, meeting=not stationary).numpy()
        r = O.front_end(pcm, 16000, stationary, n_mels, 0)
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

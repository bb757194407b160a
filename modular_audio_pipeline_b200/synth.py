"""Synthetic "speech + noise" PCM used by tests and bench.py (generator G(T, seed)).

A harmonic voiced source (f0 = 120 + 30 sin(2 pi 0.3 t) Hz, 11 harmonics ~ 1/k) times a 0.7 Hz syllabic gate
(on ~55 % of the time) times a 3.1 Hz envelope, amplitude 0.25, plus white Gaussian noise, scaled by 0.5 and
rounded to int16.  ``meeting=True`` lets the noise sigma wander 0.005..0.03 with a 90 s period (config 2).
"""
from __future__ import annotations

import math

import torch


def speech_noise(seconds: float, seed: int = 0, sr: int = 16000, sigma: float = 0.01, meeting: bool = False,
                 device: str | torch.device = "cpu") -> torch.Tensor:
    """Return int16 mono PCM of ``seconds`` at ``sr`` as a 1-D tensor on ``device``."""
    n = int(round(seconds * sr))
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234567 + seed)
    out = torch.empty(n, dtype=torch.int16, device=dev)
    block = 1 << 22
    two_pi = 2.0 * math.pi
    ph0 = (seed * 0.61803398875) % 1.0
    for s in range(0, n, block):
        e = min(s + block, n)
        t = torch.arange(s, e, device=dev, dtype=torch.float64) / sr
        phase = two_pi * (120.0 * t - (30.0 / (two_pi * 0.3)) * torch.cos(two_pi * 0.3 * t) + ph0)
        voiced = torch.zeros_like(t)
        for k in range(1, 12):
            voiced += torch.sin(k * phase) / k
        gate = torch.clamp((torch.sin(two_pi * (0.7 * t + ph0)) + 0.156) * 8.0, 0.0, 1.0)
        env = 0.6 + 0.4 * torch.sin(two_pi * 3.1 * t)
        sig = 0.25 * voiced * gate * env / 1.8
        if meeting:
            sg = 0.0175 + 0.0125 * torch.sin(two_pi * t / 90.0)
        else:
            sg = sigma
        noise = torch.randn(e - s, generator=gen, device=dev, dtype=torch.float64) * sg
        x = 0.5 * (sig + noise)
        out[s:e] = torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16)
    return out


def music_bed_stereo(seconds: float, seed: int = 0, sr: int = 48000, device: str | torch.device = "cpu") -> torch.Tensor:
    """BASELINE config 5 input: interleaved stereo int16 at ``sr`` -- the speech generator in the left
    channel, a "music bed" of three detuned sawtooth chords (amplitude 0.2) plus white noise (sigma 0.05) in both."""
    n = int(round(seconds * sr))
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(7654321 + seed)
    t = torch.arange(n, device=dev, dtype=torch.float64) / sr
    bed = torch.zeros_like(t)
    for root in (110.0, 164.81, 220.5):                         # three chords, each voice slightly detuned
        for det in (0.997, 1.0, 1.004):
            ph = (root * det * t) % 1.0
            bed += (2.0 * ph - 1.0)
    bed = 0.2 * bed / 9.0
    speech = speech_noise(seconds, seed, sr=sr, device=dev).to(torch.float64) / 32768.0
    out = torch.empty(n, 2, dtype=torch.int16, device=dev)
    for c in range(2):
        noise = torch.randn(n, generator=gen, device=dev, dtype=torch.float64) * 0.05
        x = bed + noise + (speech if c == 0 else 0.0)
        out[:, c] = torch.clamp(torch.round(0.5 * x * 32768.0), -32768, 32767).to(torch.int16)
    return out.reshape(-1)

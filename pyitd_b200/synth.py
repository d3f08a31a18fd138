"""Synthetic workloads for the five BASELINE.json configs.

There is no network and the reference ships no data besides one 8000-sample vector, so every
benchmark and every full-size test runs on these generators.  numpy versions are used where the
definition is tied to ``numpy.random.default_rng`` (configs 1 and 4); torch versions generate
on whichever device they are given so that multi-GiB batches never cross PCIe.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def config1_chirp(n: int = 65536, seed: int = 0) -> np.ndarray:
    """Config 1: one chirp + noise signal, float64 (the reference's own CPU-sized case)."""
    t = np.arange(n) / n
    rng = np.random.default_rng(seed)
    return np.sin(2 * np.pi * (5 * t + 200 * t * t)) + 0.1 * rng.standard_normal(n)


def eeg_like(n_channels: int, n: int = 65536, seed: int = 1234, device="cpu",
             dtype=torch.float64, first_channel: int = 0, total_channels: int | None = None,
             chunk: int = 512) -> torch.Tensor:
    """Configs 2 and 5: pink noise + 10 Hz (phase-shifted per channel) + 50 Hz + white noise,
    sampled at 256 Hz.  ``first_channel``/``total_channels`` let a shard generate its slice of a
    larger batch with the same per-channel phase it would have had unsharded."""
    total = n_channels if total_channels is None else total_channels
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty((n_channels, n), dtype=dtype, device=device)
    t = torch.arange(n, dtype=torch.float64, device=device) / 256.0
    f = torch.fft.rfftfreq(n, d=1.0 / 256.0).to(device=device, dtype=torch.float64)
    f[0] = f[1]
    shaping = 1.0 / torch.sqrt(f)
    for c0 in range(0, n_channels, chunk):
        c1 = min(c0 + chunk, n_channels)
        w = torch.randn((c1 - c0, n), generator=gen, dtype=torch.float64, device=device)
        pink = torch.fft.irfft(torch.fft.rfft(w, dim=1) * shaping, n=n, dim=1)
        pink = pink / pink.std(dim=1, keepdim=True)
        ch = torch.arange(first_channel + c0, first_channel + c1, dtype=torch.float64,
                          device=device).unsqueeze(1)
        x = pink + 0.5 * torch.sin(2 * math.pi * 10.0 * t + 2 * math.pi * ch / total)
        x = x + 0.2 * torch.sin(2 * math.pi * 50.0 * t)
        x = x + 0.05 * torch.randn((c1 - c0, n), generator=gen, dtype=torch.float64, device=device)
        out[c0:c1] = x.to(dtype)
    return out


def long_signal(n: int = 1 << 28, seed: int = 3, device="cpu", chunk: int = 1 << 24) -> torch.Tensor:
    """Config 3: one long float32 signal, four harmonics + white noise."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty(n, dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        e = min(s + chunk, n)
        t = torch.arange(s, e, dtype=torch.float64, device=device) / n
        x = torch.zeros(e - s, dtype=torch.float64, device=device)
        for k in (3, 17, 257, 4099):
            x += torch.sin(2 * math.pi * k * t) / math.sqrt(k)
        x += 0.25 * torch.randn(e - s, generator=gen, dtype=torch.float64, device=device)
        out[s:e] = x.to(torch.float32)
    return out


def audio_frames(seconds: float = 600.0, fs: int = 48000, frame: int = 8192, seed: int = 7) -> np.ndarray:
    """Config 4: harmonic tone with vibrato and tremolo + noise, float32, cut into
    non-overlapping frames (the ragged tail is dropped) -> (3515, 8192) at the defaults."""
    n = int(round(seconds * fs))
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    f0 = 110.0 * (1 + 0.01 * np.sin(2 * np.pi * 5 * t))
    phi = 2 * np.pi * np.cumsum(f0) / fs
    x = np.zeros(n)
    for h in range(1, 9):
        x += np.sin(h * phi) / h
    x *= 0.5 + 0.5 * np.sin(2 * np.pi * 0.5 * t) ** 2
    x += 0.01 * rng.standard_normal(n)
    nframes = n // frame
    return x[: nframes * frame].astype(np.float32).reshape(nframes, frame)

"""Synthetic clips for benchmarking and parity tests (there is no dataset in the container).

`synthetic_clip` follows SURVEY.md 8d: per channel a sum of low-frequency 2-D sinusoids whose phases drift
linearly with the frame index, one moving soft-edged disc and 2 % uniform noise, clipped to [0, 1] and
quantised to k/255 (what torchvision ToTensor yields for 8-bit PNG frames, reference model.py:64-65).
Frames are returned as uint8 [N,3,H,W]; frame i has the normalised index i/N (reference model.py:37).
"""
import math

import torch


def synthetic_clip(n_frames, height, width, seed=1234, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_waves = 5
    fy = torch.rand(3, n_waves, generator=g) * 3.0 + 0.5
    fx = torch.rand(3, n_waves, generator=g) * 3.0 + 0.5
    ph = torch.rand(3, n_waves, generator=g) * 2 * math.pi
    drift = (torch.rand(3, n_waves, generator=g) - 0.5) * 1.5
    amp = torch.rand(3, n_waves, generator=g) * 0.12 + 0.03
    base = torch.rand(3, generator=g) * 0.3 + 0.35
    noise_seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g))
    dev = torch.device(device)
    fy, fx, ph, drift, amp, base = [t.to(dev) for t in (fy, fx, ph, drift, amp, base)]
    ys = torch.linspace(0, 1, height, device=dev).view(1, 1, height, 1)
    xs = torch.linspace(0, 1, width, device=dev).view(1, 1, 1, width)
    ng = torch.Generator(device=dev).manual_seed(noise_seed)
    frames = torch.empty(n_frames, 3, height, width, dtype=torch.uint8, device=dev)
    for i in range(n_frames):
        t = i / max(n_frames, 1)
        arg = 2 * math.pi * (fy.view(3, n_waves, 1, 1) * ys + fx.view(3, n_waves, 1, 1) * xs) \
            + ph.view(3, n_waves, 1, 1) + 2 * math.pi * drift.view(3, n_waves, 1, 1) * t
        img = base.view(3, 1, 1) + (amp.view(3, n_waves, 1, 1) * torch.sin(arg)).sum(1)
        cy, cx = 0.5 + 0.3 * math.sin(2 * math.pi * t), 0.5 + 0.35 * math.cos(2 * math.pi * t)
        r2 = (ys[0] - cy) ** 2 + ((xs[0] - cx) * width / height) ** 2
        disc = torch.sigmoid((0.15 ** 2 - r2) * 400.0)
        img = img * (1 - 0.6 * disc) + 0.6 * disc * torch.tensor([0.9, 0.4, 0.2], device=dev).view(3, 1, 1)
        img = img + (torch.rand(3, height, width, generator=ng, device=dev) - 0.5) * 0.04
        frames[i] = (img.clamp(0, 1) * 255.0).round().to(torch.uint8)
    return frames


def frame_index(i, n_frames):
    """Normalised frame index fed to the positional encoding (reference model.py:37)."""
    return float(i) / n_frames

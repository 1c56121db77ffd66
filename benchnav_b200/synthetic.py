"""Synthetic benchmark inputs: fBm terrain -> Horn slopes -> terrain classes -> per-class slip
model -> Normal(mean, std) slip prediction; the ``expected_value`` risk map is its mean.

Vectorised stand-ins for the reference's offline generators (terrain_properties.py:237-351, 398-443;
slip_model.py:81-113; scripts/generate_terrain_dataset.py:31-34) -- used only to make inputs of the right
shape and statistics for bench.py and the large-size tests; deterministic given ``seed``.
"""

from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch


def fbm_heights(g: int, resolution: float, rng: np.random.Generator, roughness: float = 0.75, gain: float = 10.0
                ) -> np.ndarray:
    """Spectral fractional-Brownian surface: amplitude ~ f^-(H+1), random phases, real inverse FFT."""
    n = g + 2  # one cell of padding for the slope stencil
    fy = np.fft.fftfreq(n)[:, None] * n
    fx = np.fft.rfftfreq(n)[None, :] * n
    f = np.sqrt(fx * fx + fy * fy)
    amp = np.where(f > 0, np.power(np.maximum(f, 1e-9), -(roughness + 1.0)), 0.0)
    phase = rng.uniform(0.0, 2.0 * np.pi, size=amp.shape)
    surf = np.fft.irfft2(amp * np.exp(1j * phase), s=(n, n))
    surf = surf / (np.abs(surf).max() + 1e-12) * gain * resolution * 0.6
    return (surf - surf.min()).astype(np.float32)


def horn_slopes_deg(heights_padded: np.ndarray, resolution: float) -> np.ndarray:
    h = heights_padded
    gx = ((h[:-2, 2:] + 2 * h[1:-1, 2:] + h[2:, 2:]) - (h[:-2, :-2] + 2 * h[1:-1, :-2] + h[2:, :-2])) / (8 * resolution)
    gy = ((h[2:, :-2] + 2 * h[2:, 1:-1] + h[2:, 2:]) - (h[:-2, :-2] + 2 * h[:-2, 1:-1] + h[:-2, 2:])) / (8 * resolution)
    return np.degrees(np.arctan(np.sqrt(gx * gx + gy * gy))).astype(np.float32)


def class_map(g: int, rng: np.random.Generator, n_classes: int = 4) -> np.ndarray:
    """Low-frequency noise field thresholded into ``n_classes`` occupied terrain classes."""
    coarse = rng.standard_normal((max(4, g // 16), max(4, g // 16)))
    reps = int(np.ceil(g / coarse.shape[0]))
    field = np.kron(coarse, np.ones((reps, reps)))[:g, :g]
    k = np.ones(5) / 5.0
    for _ in range(3):  # cheap smoothing
        field = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, field)
        field = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, field)
    edges = np.quantile(field, np.linspace(0, 1, n_classes + 1)[1:-1])
    return np.digitize(field, edges).astype(np.int64)


def slip_distribution(slopes_deg: np.ndarray, classes: np.ndarray, rng: np.random.Generator, n_total: int = 10
                      ) -> Tuple[np.ndarray, np.ndarray]:
    """Per-class latent slip model: mean = clamp(s*1e-3*|phi|^n + o, 0, 1), std heteroscedastic."""
    sens = rng.uniform(1.0, 9.0, n_total)
    nonl = rng.uniform(1.4, 2.0, n_total)
    offs = rng.uniform(0.0, 0.1, n_total)
    noise = rng.uniform(0.1, 0.2, n_total)
    occupied = rng.choice(n_total, size=int(classes.max()) + 1, replace=False)
    c = occupied[classes]
    phi = np.abs(slopes_deg)
    mean = np.clip(sens[c] * 1e-3 * np.power(phi, nonl[c]) + offs[c], 0.0, 1.0)
    std = noise[c] * (0.2 + 0.02 * phi)
    return mean.astype(np.float32), std.astype(np.float32)


def make_terrain(grid_size: int, resolution: float = 0.5, seed: int = 0) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    hp = fbm_heights(grid_size, resolution, rng)
    # rescale the relief so that the slope statistics do not depend on the grid size: 95th percentile = 20 deg
    p95 = np.percentile(np.tan(np.radians(horn_slopes_deg(hp, resolution))), 95)
    hp = (hp * (np.tan(np.radians(20.0)) / max(p95, 1e-9))).astype(np.float32)
    slopes = horn_slopes_deg(hp, resolution)
    classes = class_map(grid_size, rng)
    mean, std = slip_distribution(slopes, classes, rng)
    return {"heights": torch.from_numpy(hp[1:-1, 1:-1].copy()), "slopes": torch.from_numpy(slopes),
            "t_classes": torch.from_numpy(classes), "slip_mean": torch.from_numpy(mean),
            "slip_std": torch.from_numpy(std)}


def benchmark_problem(grid_size: int, resolution: float = 0.5, seed: int = 0):
    """Terrain + problem of SURVEY 8d: start (8, 8, pi/4), goal (0.375 G r, 0.375 G r), threshold 0.3.

    Returns (risk [G,G] fp32 on CPU, start state [3], goal [2], stuck_threshold).
    """
    terr = make_terrain(grid_size, resolution, seed)
    risk = terr["slip_mean"].clone()
    # keep the start cell drivable so that rollouts actually move (the reference env samples start cells likewise)
    c = int(8.0 / resolution)
    if c < grid_size:
        risk[max(0, c - 1):c + 2, max(0, c - 1):c + 2] = torch.clamp(risk[max(0, c - 1):c + 2, max(0, c - 1):c + 2], max=0.2)
    lim = grid_size * resolution
    start = torch.tensor([min(8.0, 0.25 * lim), min(8.0, 0.25 * lim), float(np.pi / 4)], dtype=torch.float32)
    goal = torch.tensor([0.375 * lim, 0.375 * lim], dtype=torch.float32)
    return risk, start, goal, 0.3

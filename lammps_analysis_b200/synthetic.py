"""Seeded synthetic trajectories for the BASELINE configs.

Everything is generated with ``numpy.random.default_rng(seed)`` on the host (small
configs) or ``torch.Generator`` Philox on the device (C4 / C5), in the MDSuite store
layout ``(n_atoms, n_frames, 3)`` float32, atom-major.
"""
from __future__ import annotations

import numpy as np


def rocksalt_lattice(n_side: int, box: float) -> np.ndarray:
    """(n_side^3, 3) simple-cubic sites; species alternate by parity (rock salt)."""
    g = np.arange(n_side)
    ijk = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    return ijk, (ijk + 0.5) * (box / n_side)


def nacl_trajectory(n_atoms: int, n_frames: int, box: float, seed: int, sigma_step: float = 0.05,
                    v_sigma: float = 1.0, ou_gamma: float = 0.05):
    """NaCl-like system: perturbed rock-salt lattice + per-frame Gaussian random walk, wrapped
    into [0, L); velocities from a discrete Ornstein-Uhlenbeck process (analytic VACF
    v_sigma^2 * (1 - gamma)^m).

    Returns dict(species -> dict(Positions, Velocities)) float32 (A, T, 3), and box (3,).
    """
    n_side = round(n_atoms ** (1 / 3))
    if n_side**3 != n_atoms:
        raise ValueError("n_atoms must be a cube")
    rng = np.random.default_rng(seed)
    ijk, sites = rocksalt_lattice(n_side, box)
    parity = ijk.sum(axis=1) % 2
    out = {}
    for name, par in (("Na", 0), ("Cl", 1)):
        base = sites[parity == par]
        A = base.shape[0]
        steps = rng.normal(0.0, sigma_step, size=(A, n_frames, 3))
        steps[:, 0, :] = rng.normal(0.0, 0.3, size=(A, 3))
        pos = base[:, None, :] + np.cumsum(steps, axis=1)
        pos = np.mod(pos, box)
        pos32 = pos.astype(np.float32)
        pos32[pos32 >= np.float32(box)] = 0.0
        vel = np.empty((A, n_frames, 3))
        vel[:, 0] = rng.normal(0.0, v_sigma, size=(A, 3))
        noise = rng.normal(0.0, v_sigma * np.sqrt(1 - (1 - ou_gamma) ** 2), size=(A, n_frames, 3))
        for t in range(1, n_frames):
            vel[:, t] = (1 - ou_gamma) * vel[:, t - 1] + noise[:, t]
        out[name] = {"Positions": pos32, "Velocities": vel.astype(np.float32)}
    return out, np.array([box, box, box], dtype=np.float64)


def device_fluid(n_atoms: int, n_frames: int, box: float, seed: int, device, sigma_step=0.05,
                 frames_chunk: int = 64):
    """Jittered-lattice fluid with a random walk, generated on the device, wrapped into
    [0, L).  Returns a CUDA float32 tensor (n_atoms, n_frames, 3)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_side = int(np.ceil(n_atoms ** (1 / 3)))
    idx = torch.arange(n_atoms, device=device)
    ijk = torch.stack([idx // (n_side * n_side), (idx // n_side) % n_side, idx % n_side], dim=1)
    a = box / n_side
    cur = (ijk.to(torch.float32) + 0.5) * a
    cur = cur + (torch.rand(n_atoms, 3, device=device, generator=gen) - 0.5) * (0.6 * a)
    out = torch.empty(n_atoms, n_frames, 3, dtype=torch.float32, device=device)
    for t0 in range(0, n_frames, frames_chunk):
        t1 = min(n_frames, t0 + frames_chunk)
        steps = torch.randn(n_atoms, t1 - t0, 3, device=device, generator=gen) * sigma_step
        if t0 == 0:
            steps[:, 0] = 0
        walk = cur[:, None, :] + torch.cumsum(steps, dim=1)
        cur = walk[:, -1].clone()
        w = torch.remainder(walk, box)
        w[w >= box] = 0.0
        out[:, t0:t1] = w
    return out

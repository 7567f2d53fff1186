"""Synthetic terrain for the positionability configs, generated on the DEVICE (the
fixtures of BASELINE configs[2..4]).

The reference builds its maps on the host with numpy (`perlinnumpy2d.generate_fractal_noise_2d`
driven by `maps.ground`, maps.py:190-202,289-297): fine at 1 Mi points, a minute of host time for
the 50 M-point map of configs[3].  This module evaluates the same gradient noise with torch
tensor ops on the GPU: the only host work is drawing the random gradient angles — (res + 1)^2
numbers per octave from numpy's global generator, in the reference's order, so a fixed seed names
the same terrain — and the two coordinate axes.  The arithmetic is float64 mul / add / sub /
remainder in the reference's operation order (nothing is contracted across torch ops), then one cast
to float32 per layer like the reference's `.astype("float32")`.

Written for this repo (no code shared with tests/terrain.py, which restates the numpy generator
for the CPU suite and pins this one).
"""
import numpy as np


def _gradient_noise(shape, res, angles, device, torch):
    """One octave: Perlin gradient noise on a (shape[0], shape[1]) lattice with res[0] x res[1] cells."""
    n0, n1 = int(shape[0]), int(shape[1])
    r0, r1 = int(res[0]), int(res[1])
    d0, d1 = n0 // r0, n1 // r1                      # lattice points per cell
    f64 = torch.float64
    i = torch.arange(n0, device=device, dtype=f64)
    j = torch.arange(n1, device=device, dtype=f64)
    u = torch.remainder(i * (r0 / n0), 1.0)[:, None]   # position inside the cell, per axis
    v = torch.remainder(j * (r1 / n1), 1.0)[None, :]
    ci = torch.arange(n0, device=device) // d0          # cell index per axis
    cj = torch.arange(n1, device=device) // d1
    a = torch.from_numpy(angles).to(device)
    gx, gy = torch.cos(a), torch.sin(a)                 # (r0 + 1, r1 + 1) gradient table

    def corner(di, dj, du, dv):
        ii, jj = (ci + di)[:, None], (cj + dj)[None, :]
        return (u - du) * gx[ii, jj] + (v - dv) * gy[ii, jj]

    n00, n10, n01, n11 = corner(0, 0, 0.0, 0.0), corner(1, 0, 1.0, 0.0), corner(0, 1, 0.0, 1.0), corner(1, 1, 1.0, 1.0)
    fu = u * u * u * (u * (u * 6 - 15) + 10)
    fv = v * v * v * (v * (v * 6 - 15) + 10)
    lo = n00 * (1 - fu) + fu * n10
    hi = n01 * (1 - fu) + fu * n11
    return np.sqrt(2) * ((1 - fv) * lo + fv * hi)


def fractal_noise(shape, res, octaves, persistence, lacunarity, device, torch):
    """Sum of octaves; draws the gradient angles of every octave from numpy's GLOBAL generator, one
    (res + 1)^2 block per octave, like the reference (seed it first)."""
    total = torch.zeros((int(shape[0]), int(shape[1])), dtype=torch.float64, device=device)
    frequency, amplitude = 1, 1.0
    for _ in range(octaves):
        r = (frequency * res[0], frequency * res[1])
        angles = 2 * np.pi * np.random.rand(r[0] + 1, r[1] + 1)
        total += amplitude * _gradient_noise(shape, r, angles, device, torch)
        frequency *= lacunarity
        amplitude *= persistence
    return total


def perlin_terrain(n=1024, seed=42, x=(-2000.0, 2000.0), y=(-6000.0, 2000.0), device="cuda"):
    """(ny * nx, 3) float32 tensor on `device`: the lattice + two fractal layers of maps.py:289-297
    (300 mm of 5 octaves at res (8, 4), 30 mm of 3 octaves at res (32, 16)), y-major / x fastest.
    n: points per side, or (ny, nx); each a multiple of 128."""
    import torch
    ny, nx = (n, n) if np.isscalar(n) else (int(n[0]), int(n[1]))
    xs = torch.from_numpy(np.linspace(x[0], x[1], nx).astype(np.float32)).to(device)
    ys = torch.from_numpy(np.linspace(y[0], y[1], ny).astype(np.float32)).to(device)
    out = torch.empty((ny, nx, 3), dtype=torch.float32, device=device)
    out[:, :, 0] = xs[None, :]
    out[:, :, 1] = ys[:, None]
    np.random.seed(seed=seed)
    z = (fractal_noise((ny, nx), (8, 4), 5, 0.35, 2, device, torch) * 300).to(torch.float32)
    z += (fractal_noise((ny, nx), (32, 16), 3, 0.2, 2, device, torch) * 30).to(torch.float32)
    out[:, :, 2] = z
    return out.reshape(-1, 3)


def body_lattice(terrain, nx, ny, nz, z_above=350.0):
    """nx x ny x nz pose lattice over the map's xy extent and z in [zmin, zmax + z_above]
    (before.py:26-35), x-major / z fastest; torch tensor on the terrain's device."""
    import torch
    lo = terrain.min(dim=0).values.cpu().numpy()
    hi = terrain.max(dim=0).values.cpu().numpy()
    ax = [torch.from_numpy(np.linspace(lo[k], hi[k] + (z_above if k == 2 else 0.0), m, dtype=np.float32)).to(terrain.device)
          for k, m in enumerate((nx, ny, nz))]
    X, Y, Z = torch.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    return torch.stack([X, Y, Z], -1).reshape(-1, 3).contiguous()

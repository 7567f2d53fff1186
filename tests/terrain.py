"""Synthetic terrain / body-pose fixtures for the positionability configs.

`fractal_noise_2d` restates the reference's `perlinnumpy2d.generate_fractal_noise_2d`
(perlinnumpy2d.py:8-96) — same numpy operations, same use of the global numpy RNG, so a fixed seed
gives the reference's terrain bit for bit (checked in tests/test_terrain.py where the reference
tree is present).  `perlin_terrain` is the lattice + noise part of `maps.ground`
(maps.py:190-202,289-297); `body_lattice` the pose grid of before.py:24-61.
"""
import numpy as np


def _fade(t):
    return t * t * t * (t * (t * 6 - 15) + 10)


def perlin_noise_2d(shape, res):
    delta = (res[0] / shape[0], res[1] / shape[1])
    d = (shape[0] // res[0], shape[1] // res[1])
    grid = np.mgrid[0:res[0]:delta[0], 0:res[1]:delta[1]].transpose(1, 2, 0) % 1
    angles = 2 * np.pi * np.random.rand(res[0] + 1, res[1] + 1)
    gradients = np.dstack((np.cos(angles), np.sin(angles)))
    gradients = gradients.repeat(d[0], 0).repeat(d[1], 1)
    g00 = gradients[:-d[0], :-d[1]]
    g10 = gradients[d[0]:, :-d[1]]
    g01 = gradients[:-d[0], d[1]:]
    g11 = gradients[d[0]:, d[1]:]
    n00 = np.sum(np.dstack((grid[:, :, 0], grid[:, :, 1])) * g00, 2)
    n10 = np.sum(np.dstack((grid[:, :, 0] - 1, grid[:, :, 1])) * g10, 2)
    n01 = np.sum(np.dstack((grid[:, :, 0], grid[:, :, 1] - 1)) * g01, 2)
    n11 = np.sum(np.dstack((grid[:, :, 0] - 1, grid[:, :, 1] - 1)) * g11, 2)
    t = _fade(grid)
    n0 = n00 * (1 - t[:, :, 0]) + t[:, :, 0] * n10
    n1 = n01 * (1 - t[:, :, 0]) + t[:, :, 0] * n11
    return np.sqrt(2) * ((1 - t[:, :, 1]) * n0 + t[:, :, 1] * n1)


def fractal_noise_2d(shape, res, octaves=1, persistence=0.5, lacunarity=2):
    noise = np.zeros(shape)
    frequency, amplitude = 1, 1
    for _ in range(octaves):
        noise += amplitude * perlin_noise_2d(shape, (frequency * res[0], frequency * res[1]))
        frequency *= lacunarity
        amplitude *= persistence
    return noise


def perlin_terrain(n=1024, seed=42, x=(-2000.0, 2000.0), y=(-6000.0, 2000.0)):
    """Lattice with the two fractal layers of maps.py:289-297.  n: points per side, or (ny, nx);
    each a multiple of 128 (perlinnumpy2d.py:84-85)."""
    ny, nx = (n, n) if np.isscalar(n) else (int(n[0]), int(n[1]))
    xs, ys = np.linspace(x[0], x[1], nx), np.linspace(y[0], y[1], ny)
    X, Y, Z = np.meshgrid(xs, ys, 0)
    ground = np.concatenate([X.reshape(-1, 1), Y.reshape(-1, 1), Z.reshape(-1, 1)], axis=1).astype("float32")
    np.random.seed(seed=seed)
    ground[:, 2] += (fractal_noise_2d((ny, nx), (8, 4), 5, 0.35, 2) * 300).reshape(-1).astype("float32")
    ground[:, 2] += (fractal_noise_2d((ny, nx), (32, 16), 3, 0.2, 2) * 30).reshape(-1).astype("float32")
    return np.ascontiguousarray(ground, np.float32)


def body_lattice(terrain, nx, ny, nz, z_above=350.0):
    """nx x ny x nz pose lattice over the map's xy extent and z in [zmin, zmax + 350]
    (before.py:26-35), x-major / z fastest."""
    lo = terrain.min(axis=0)
    hi = terrain.max(axis=0)
    xs = np.linspace(lo[0], hi[0], nx, dtype=np.float32)
    ys = np.linspace(lo[1], hi[1], ny, dtype=np.float32)
    zs = np.linspace(lo[2], hi[2] + z_above, nz, dtype=np.float32)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    return np.ascontiguousarray(np.stack([X, Y, Z], -1).reshape(-1, 3), np.float32)


def sine_terrain(n=49, extent=600.0, amp=60.0):
    gx, gy = np.meshgrid(np.linspace(-extent, extent, n, dtype=np.float32),
                         np.linspace(-extent, extent, n, dtype=np.float32), indexing="ij")
    gz = (amp * np.sin(gx / 170) * np.cos(gy / 230)).astype(np.float32)
    return np.ascontiguousarray(np.stack([gx, gy, gz], -1).reshape(-1, 3), np.float32)

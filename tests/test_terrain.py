"""CPU-only: the fixture generator restates the reference's Perlin terrain exactly."""
import os
import sys

import numpy as np
import pytest

from tests import terrain


def test_fractal_noise_matches_reference_generator():
    ref_dir = "/root/reference"
    if not os.path.exists(os.path.join(ref_dir, "perlinnumpy2d.py")):
        pytest.skip("reference tree not present")
    sys.path.insert(0, ref_dir)
    try:
        import perlinnumpy2d as ref
    finally:
        sys.path.remove(ref_dir)
    np.random.seed(42)
    want = ref.generate_fractal_noise_2d(shape=(256, 256), res=(8, 4), octaves=5, persistence=0.35, lacunarity=2)
    np.random.seed(42)
    got = terrain.fractal_noise_2d((256, 256), (8, 4), 5, 0.35, 2)
    assert np.array_equal(got, want)


def test_perlin_terrain_shape_and_determinism():
    t1 = terrain.perlin_terrain(128)
    t2 = terrain.perlin_terrain(128)
    assert t1.shape == (128 * 128, 3) and t1.dtype == np.float32
    assert np.array_equal(t1, t2)
    assert t1[:, 0].min() == -2000 and t1[:, 1].max() == 2000
    assert 100 < np.ptp(t1[:, 2]) < 1500
    b = terrain.body_lattice(t1, 4, 5, 6)
    assert b.shape == (120, 3) and b[1, 2] > b[0, 2]  # z fastest


def test_device_generator_matches_the_numpy_restatement(lrm):
    """lrm_b200.fixtures.perlin_terrain (torch tensor ops; here on the CPU device) returns the bytes
    of the numpy generator for the same seed, on square and ragged shapes; body_lattice too."""
    torch = pytest.importorskip("torch")
    from importlib import import_module
    fx = import_module("lrm_b200.fixtures")
    for shape in (128, (256, 128), (384, 512)):
        want = terrain.perlin_terrain(shape, seed=7)
        got = fx.perlin_terrain(shape, seed=7, device="cpu").numpy()
        assert np.array_equal(want, got), shape
    t = terrain.perlin_terrain(128)
    assert np.array_equal(terrain.body_lattice(t, 5, 6, 7), fx.body_lattice(torch.from_numpy(t), 5, 6, 7).numpy())


@pytest.mark.gpu
def test_device_generator_on_the_gpu(lrm):
    """The same on cuda:0 at BASELINE configs[2]'s size (1024 x 1024): float64 mul / add / remainder
    are exact IEEE operations on both sides, so the terrain is the reference's bit for bit."""
    torch = pytest.importorskip("torch")
    from importlib import import_module
    fx = import_module("lrm_b200.fixtures")
    want = terrain.perlin_terrain(1024)
    got = fx.perlin_terrain(1024, device="cuda").cpu().numpy()
    assert np.array_equal(want, got)

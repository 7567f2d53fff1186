"""File-protocol driver (legged-robot-movability-cuda_b200/driver/lrm_cuda.cpp): the reference's `cuda`
executable (several_leg.cpp:124-223) over the C ABI — same input / output files."""
import os
import subprocess

import numpy as np
import pytest

from tests import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "legged-robot-movability-cuda_b200")


def _driver():
    import lrm_loader
    lrm_loader.build()
    exe = os.path.join(PKG, "lrm_cuda")
    assert os.path.exists(exe)
    return exe


def _write_inputs(d, pts):
    for k, c in enumerate("xyz"):
        np.ascontiguousarray(pts[:, k], np.float32).tofile(os.path.join(d, f"dist_input_t{c}.bin"))


def test_driver_fails_loudly_without_gpu(tmp_path):
    """No CPU mode: without a device the driver reports the library's CUDA error and exits 1
    (CUDA_CHECK_ERROR convention, cross_compiled.cu:12-20); a missing input file is an error too."""
    import torch
    exe = _driver()
    r = subprocess.run([exe, "--dir", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "Error opening file" in r.stderr
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _write_inputs(str(tmp_path), np.zeros((8, 3), np.float32))
    r = subprocess.run([exe, "--dir", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA error" in r.stderr
    assert not os.path.exists(tmp_path / "out_reachability.bin")


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_driver_round_trip(tmp_path, oracle, fused):
    """before.py-style query grid in, out_reachability.bin / out_dist_x{x,y,z}.bin out, checked
    against the oracle (same bars as the library tests)."""
    exe = _driver()
    g = [np.linspace(lo, hi, n, dtype=np.float32) for lo, hi, n in ((-100, 600, 71), (-400, 400, 41), (-500, 200, 71))]
    pts = np.stack(np.meshgrid(*g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    _write_inputs(str(tmp_path), pts)
    r = subprocess.run([exe, "--dir", str(tmp_path)] + (["--fused"] if fused else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "ns per point" in r.stdout
    reach = np.fromfile(tmp_path / "out_reachability.bin", np.uint8)
    vec = np.stack([np.fromfile(tmp_path / f"out_dist_x{c}.bin", np.float32) for c in "xyz"], 1)
    assert len(reach) == len(pts) and vec.shape == pts.shape
    leg = oracle.get_leg(1, 0.0)
    fr = parity.flag_report(pts, reach, oracle.reach(pts, leg, threads=8), lambda p: oracle.reach(p, leg, threads=8))
    dr = parity.dist_report(pts, vec, oracle.dist(pts, leg, threads=8)[0], lambda p: oracle.dist(p, leg, threads=8)[0])
    assert fr["unexplained"] == 0 and dr["unexplained"] == 0, (fr, dr)

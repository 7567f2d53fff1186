"""CPU-only: the C-ABI library builds, loads, exports every symbol include/lrm_c.h declares, its
pure-host entry points agree with the oracle, and compute calls fail loudly (no CPU fallback)
when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lrm_c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lrm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = _declared_symbols()
    for s in ("lrm_reach", "lrm_dist", "lrm_reach_dist", "lrm_reach_dist_soa", "lrm_positionability",
              "lrm_forward_kine", "lrm_make_lattice", "lrm_last_error", "lrm_default_leg"):
        assert s in syms


def test_library_exports_every_declared_symbol(lrm):
    L = lrm.lib()
    for s in _declared_symbols():
        assert hasattr(L, s), f"liblrm_b200.so does not export {s}"
    assert L.lrm_abi_version() == 1


def test_no_torch_or_cxx_types_in_abi():
    text = open(os.path.join(ROOT, "include", "lrm_c.h")).read()
    assert "torch" not in text and "std::" not in text and "template" not in text


def test_leg_struct_layout(lrm):
    assert ctypes.sizeof(lrm.LegDimensions) == 56  # sizeof(LegDimensions), SURVEY §2


def test_default_legs_match_oracle(lrm, port):
    for robot in (0, 1):
        for az in (0.0, 0.7853982, 3.0):
            leg = lrm.get_leg(robot, az)
            assert np.array_equal(leg.as_array().view(np.uint32), port.get_leg(robot, az).view(np.uint32))
    with pytest.raises(lrm.LrmError):
        lrm.get_leg(2, 0.0)


def test_full_struct_orientations_match_oracle(lrm, port, golden):
    q = lrm.full_struct_orientations()
    assert q.shape == (45, 4)
    assert np.array_equal(q.view(np.uint32), port.full_struct_orientations().view(np.uint32))
    assert np.array_equal(q.view(np.uint32), golden["full_struct_quats"].view(np.uint32))
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-6)


def test_rpy_to_quat_matches_oracle(lrm, port):
    """RPYtoQuat (octree_util.cu.h:164-172), incl. the SURVEY appendix D known answers."""
    for r, p, y in ((0, 0, 0), (-0.7853982, 0, 0), (-0.7853982, -0.3926991, -0.3926991), (0.1, -0.2, 2.5)):
        assert np.array_equal(lrm.rpy_to_quat(r, p, y).view(np.uint32),
                              np.asarray(port.rpy_to_quat(r, p, y), np.float32).view(np.uint32))
    assert np.array_equal(lrm.rpy_to_quat(0, 0, 0), np.array([-1, 0, 0, 0], np.float32))
    assert np.allclose(lrm.rpy_to_quat(-0.7853982, 0, 0), [-0.92388, -0.38268, 0, 0], atol=1e-5)
    q = lrm.yaw_orientations(16)
    assert q.shape == (16, 4) and np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-6)


def test_argument_validation(lrm):
    L = lrm.lib()
    leg = lrm.get_M2_leg()
    assert L.lrm_reach(None, 4, ctypes.byref(leg), None, None, 0, None, None) == -1
    assert b"NULL" in L.lrm_last_error()
    assert L.lrm_reach(None, 0, None, None, None, 0, None, None) == -1
    lo = (ctypes.c_float * 3)(0, 0, 0)
    dims = (ctypes.c_uint32 * 3)(2, 2, 2)
    assert L.lrm_make_lattice(None, lo, lo, dims, 0, 9, None) == -1  # beyond the lattice


def test_compute_fails_loudly_without_gpu(lrm):
    """The product has no CPU path: on a machine without CUDA a compute call must raise."""
    if lrm.lib().lrm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    pts = np.zeros((8, 3), np.float32)
    with pytest.raises(lrm.LrmError) as e:
        lrm.reachability(pts, lrm.get_M2_leg())
    assert "lrm error -2" in str(e.value)
    with pytest.raises(lrm.LrmError):
        lrm.distance(pts, lrm.get_M2_leg())
    with pytest.raises(lrm.LrmError):
        lrm.positionability(pts, pts, [lrm.get_M2_leg()] * 4)


def test_lattice_host_formula(lrm):
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (5, 4, 3))
    pts = lrm.lattice_host(lo, step, dims)
    assert pts.shape == (60, 3)
    assert np.array_equal(pts[0], [-100, -400, -500])
    assert np.array_equal(pts[1], [-100, -400, np.float32(-500) + np.float32(1) * step[2]])  # z fastest
    assert np.array_equal(pts[3][:2], [-100, np.float32(-400) + step[1]])
    assert np.allclose(pts[-1], [600, 400, 200], atol=1e-3)
    sub = lrm.lattice_host(lo, step, dims, first=17, count=9)
    assert np.array_equal(sub, pts[17:26])


def test_product_never_references_the_oracle():
    """Guard the rule that only tests/, smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "legged-robot-movability-cuda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_port" not in text and "libref_oracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
    inc = open(os.path.join(ROOT, "include", "lrm_c.h")).read()
    assert "oracle_port" not in inc


def test_compat_header_compiles(lrm, tmp_path):
    """A reference-style call site (bench.cpp:120-152 shape) builds with plain g++ against
    include/lrm_compat.hpp + liblrm_b200.so; without a GPU it must die like CUDA_CHECK_ERROR."""
    import subprocess
    exe = tmp_path / "compat_example"
    libdir = os.path.dirname(lrm.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "compat_example.cpp"), "-L", libdir, "-llrm_b200",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if lrm.lib().lrm_device_count() == 0:
        assert r.returncode != 0 and "CUDA error in" in r.stderr
    else:
        assert r.returncode == 0 and "reachable" in r.stdout

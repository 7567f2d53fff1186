"""Pins against the reference's OWN GPU code: the multi-leg pipeline and the octree
predicate are __device__-only in the reference, so the CPU restatement (oracle_port.c) cannot be
checked against a compiled CPU twin.  Instead the unmodified reference sources are compiled for
sm_100 without the fast-math family (oracle/Makefile `refgpu`, oracle/_ref/libref_gpu*_precise.so,
shipped prebuilt to the GPU box) and

  * -m gpu: run there in a fresh subprocess (tools/refgpu_dump.py) next to the product and the
    restatement — all three must agree exactly;
  * CPU suite: the same outputs, captured on a B200 by tests/golden/make_refgpu_golden.py and
    committed as tests/golden/refgpu_*.npz, pin the restatement without a GPU.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import pin_scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")


def _key(a):
    return {tuple(r) for r in np.ascontiguousarray(a, np.float32).view(np.uint32).reshape(-1, 3).tolist()}


def _dump(what, tmp_path, lib):
    if not os.path.exists(os.path.join(REF, lib)):
        pytest.skip(f"oracle/_ref/{lib} not built (make -C oracle refgpu needs /root/reference)")
    out = str(tmp_path / f"refgpu_{what}.npz")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "refgpu_dump.py"), what, out], check=True,
                   cwd=ROOT, timeout=600)
    return np.load(out)


def _load_golden(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not captured yet (tests/golden/make_refgpu_golden.py on a GPU box)")
    return np.load(path)


# ---- robot_full_struct (several_leg.cu:796-877) ---------------------------------------------
def _check_full(ref, port, lrm=None):
    quats = port.full_struct_orientations()
    for name, (terr, bodies, legs) in pin_scenes.full_struct_scenes().items():
        want = _key(ref[name + "_standable_xyz"])
        assert len(want) > 1000, name          # the scene must exercise the predicate
        got_port = port.standability(bodies, terr, [l for l in legs], quats, pre_cull=True,
                                     threads=os.cpu_count() or 1)
        kp = _key(bodies[got_port != 0])
        assert kp == want, (name, "restatement vs reference GPU", len(kp ^ want))
        if lrm is not None:
            llegs = [lrm.LegDimensions.from_array(l) for l in legs]
            got = lrm.positionability(bodies, terr, llegs, quats, pre_cull=True)
            kg = _key(bodies[got != 0])
            assert kg == want, (name, "lrm_positionability vs reference GPU", len(kg ^ want))
            # first-orientation index too (the reference only returns the set)
            assert np.array_equal(got, got_port), name


def test_restatement_matches_reference_gpu_full_struct_golden(port):
    """op_standability == the reference's robot_full_struct (precise build) run on a B200."""
    _check_full(_load_golden("refgpu_full_struct.npz"), port)


@pytest.mark.gpu
def test_full_struct_matches_reference_gpu_live(lrm, port, tmp_path):
    """robot_full_struct of the reference, on this GPU, == lrm_positionability(pre_cull) ==
    op_standability: symmetric difference 0 on both scenes."""
    ref = _dump("full", tmp_path, "libref_gpu_several_precise.so")
    _check_full(ref, port, lrm)
    gold = os.path.join(GOLD, "refgpu_full_struct.npz")
    if os.path.exists(gold):
        g = np.load(gold)
        for name in pin_scenes.full_struct_scenes():
            assert _key(g[name + "_standable_xyz"]) == _key(ref[name + "_standable_xyz"]), name


# ---- validity_child (several_leg_octree.cu:19-151) ------------------------------------------
def _check_oct(ref, port, lrm=None):
    foot = pin_scenes.oct_footholds()
    seen = set()
    for name, (box, pv, leg) in pin_scenes.oct_cases().items():
        want_f, want_b = ref[name + "_flags"], ref[name + "_boxes"]
        pf, pb = port.validity_children(box, pv, foot, leg)
        assert np.array_equal(pb, want_b), (name, "child boxes")
        assert np.array_equal(pf, want_f), (name, "restatement vs validity_child", pf.T.tolist(), want_f.T.tolist())
        if lrm is not None:
            gf, gb = lrm.oct_children(foot, lrm.LegDimensions.from_array(leg), box, pv)
            assert np.array_equal(gb, want_b), (name, "child boxes (product)")
            assert np.array_equal(gf, want_f), (name, "lrm_oct_children vs validity_child", gf.T.tolist(),
                                                want_f.T.tolist())
        seen |= {tuple(r) for r in want_f.tolist()}
    # the cases must reach every outcome of the write-back (:134-150) that a live child can have
    assert {(0, 0, 1, 0), (0, 0, 1, 1), (1, 0, 1, 1), (1, 1, 1, 0)} <= seen, seen


def test_restatement_matches_reference_gpu_validity_child_golden(port):
    """op_validity_children == the reference's validity_child kernel (precise build, launched as
    one thread: the only shape in which its __syncthreads() is not in divergent code) run on a
    B200, per child and per flag, on every case of pin_scenes.oct_cases()."""
    _check_oct(_load_golden("refgpu_validity_child.npz"), port)


@pytest.mark.gpu
def test_validity_child_matches_reference_gpu_live(lrm, port, tmp_path):
    """The reference's validity_child on this GPU == lrm_oct_children == op_validity_children."""
    _check_oct(_dump("oct", tmp_path, "libref_gpu_precise.so"), port, lrm)

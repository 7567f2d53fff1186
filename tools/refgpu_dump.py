"""TEST INFRASTRUCTURE: run the reference's OWN GPU code (compiled unmodified for sm_100 without the
fast-math family, oracle/_ref/libref_gpu*_precise.so) on the fixed pin scenes and write what it
returns to an .npz.  Needs a GPU; no torch, no product library: the reference's thrust pipeline
must own a fresh CUDA context (tools/vs_refgpu.py).

    python tools/refgpu_dump.py full out.npz   # robot_full_struct (several_leg.cu:796-877)
    python tools/refgpu_dump.py oct  out.npz   # validity_child    (several_leg_octree.cu:19-151)

tests/test_refgpu_pin.py runs this in a subprocess on the GPU box and compares the product and the
CPU restatement with it; tests/golden/make_refgpu_golden.py stores the same outputs as committed
fixtures for the CPU suite.
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import pin_scenes  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
sz, vp = ctypes.c_size_t, ctypes.c_void_p


def quiet_call(fn, *a):
    """The reference prints a line per orientation / node."""
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        return fn(*a)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


def dump_full(out):
    s = ctypes.CDLL(os.path.join(REF, "libref_gpu_several_precise.so"))
    s.refgpu_full_struct.restype = ctypes.c_double
    s.refgpu_full_struct.argtypes = [vp, sz, vp, sz, vp, sz, vp, sz, ctypes.POINTER(sz)]
    res = {}
    for name, (terr, bodies, legs) in pin_scenes.full_struct_scenes().items():
        la = np.ascontiguousarray(legs, np.float32)
        xyz = np.empty((len(bodies), 3), np.float32)
        cnt = sz(0)
        ms = quiet_call(s.refgpu_full_struct, bodies.ctypes.data, len(bodies), terr.ctypes.data, len(terr),
                        la.ctypes.data, 4, xyz.ctypes.data, len(bodies), ctypes.byref(cnt))
        res[name + "_standable_xyz"] = xyz[:cnt.value].copy()
        res[name + "_wall_ms"] = np.float64(ms)
        print(name, "standable", cnt.value, "of", len(bodies), f"{ms:.0f} ms", flush=True)
    np.savez_compressed(out, **res)


def dump_oct(out):
    g = ctypes.CDLL(os.path.join(REF, "libref_gpu_precise.so"))
    g.refgpu_validity_child.restype = ctypes.c_int
    g.refgpu_validity_child.argtypes = [vp, ctypes.c_int, vp, sz, vp, vp, vp]
    res = {}
    foot = pin_scenes.oct_footholds()
    for name, (box, pv, leg) in pin_scenes.oct_cases().items():
        flags = np.zeros((8, 4), np.uint8)
        boxes = np.zeros((8, 6), np.float32)
        la = np.ascontiguousarray(leg, np.float32)
        b = np.ascontiguousarray(box, np.float32)
        t0 = time.perf_counter()
        rc = quiet_call(g.refgpu_validity_child, b.ctypes.data, int(pv), foot.ctypes.data, len(foot),
                        la.ctypes.data, flags.ctypes.data, boxes.ctypes.data)
        assert rc == 0, (name, rc)
        res[f"{name}_flags"] = flags
        res[f"{name}_boxes"] = boxes
        print(name, flags.T.tolist(), f"{time.perf_counter() - t0:.2f} s", flush=True)
        np.savez_compressed(out, **res)   # after every case: a slow case still leaves the earlier ones


if __name__ == "__main__":
    {"full": dump_full, "oct": dump_oct}[sys.argv[1]](sys.argv[2])

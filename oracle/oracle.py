"""TEST INFRASTRUCTURE ONLY — ctypes access to the CPU oracles.

  port  : oracle/liboracle_port.so   plain-C restatement (oracle_port.c), builds anywhere with gcc
  ref   : oracle/_ref/libref_oracle.so  the reference's own sources compiled where they lie under
          /root/reference (oracle/Makefile, ref_shim.cu); built in the container, shipped prebuilt

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module.  Nothing under legged-robot-movability-cuda_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "liboracle_port.so")
REF_LIB = os.path.join(HERE, "_ref", "libref_oracle.so")
REFERENCE_TREE = "/root/reference"

_vp, _sz, _ci, _cf = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float


def build(ref=True, quiet=True):
    """Compile the C restatement, and the reference shim when the reference tree is present."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.run(["make", "-C", HERE, "port"], check=True, stdout=out)
    if ref and os.path.isdir(REFERENCE_TREE):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=out)


def _as_f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        assert a.ndim == 2 and a.shape[1] == cols
    return a


def leg_array(leg):
    """Accept a 14-float array or any ctypes structure with the LegDimensions layout."""
    if isinstance(leg, np.ndarray):
        a = np.ascontiguousarray(leg, dtype=np.float32)
    else:
        a = np.frombuffer(bytes(leg), dtype=np.float32).copy()
    assert a.size == 14
    return a


class _Oracle:
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.L = ctypes.CDLL(self.path)

    def _f(self, name):
        return getattr(self.L, self.prefix + name)


class PortOracle(_Oracle):
    """oracle_port.c"""
    prefix = "op_"
    path = PORT_LIB
    kind = "port"

    def __init__(self):
        if not os.path.exists(self.path):
            build(ref=False)
        super().__init__()
        L = self.L
        L.op_get_leg.argtypes = [_ci, _cf, _vp]
        L.op_reach.argtypes = [_vp, _sz, _vp, _vp, _vp, _ci]
        L.op_dist.argtypes = [_vp, _sz, _vp, _vp, _vp, _vp, _ci]
        L.op_find_region.argtypes = [_cf, _cf, _vp]
        L.op_insert_circles.argtypes = [_cf, _cf, _vp, _vp]
        L.op_insert_intersec.argtypes = [_vp, _vp]
        L.op_qt_rotate_p.argtypes = [_vp, _vp, _vp]
        L.op_qt_multiply_p.argtypes = [_vp, _vp, _vp]
        L.op_quat_from_vect_angle_p.argtypes = [_vp, _cf, _vp]
        L.op_rpy_to_quat_p.argtypes = [_cf, _cf, _cf, _vp]
        L.op_rotate_leg_data_p.argtypes = [_vp, _vp, _vp]
        L.op_full_struct_orientations.argtypes = [_vp]
        L.op_quaternion_from_angle_index.argtypes = [ctypes.c_uint, _vp]
        L.op_standability.argtypes = [_vp, _sz, _vp, _sz, _vp, _ci, _vp, _ci, _ci, _vp, _ci]
        L.op_apply_recurs.argtypes = [_vp, _sz, _vp, _vp, _ci, _vp]
        L.op_apply_oct.argtypes = [_vp, _sz, _vp, _ci, _vp, _sz]
        L.op_apply_oct.restype = _sz
        L.op_validity_children.argtypes = [_vp, _ci, _vp, _sz, _vp, _vp, _vp]
        L.op_create_child_box.argtypes = [_vp, ctypes.c_uint, _vp, _vp, _vp]
        L.op_create_child_box.restype = ctypes.c_uint

    def get_leg(self, robot, azimuth=0.0):
        out = np.zeros(14, np.float32)
        self.L.op_get_leg(robot, azimuth, out.ctypes.data)
        return out

    def reach(self, pts, leg, quat=(1, 0, 0, 0), threads=1):
        pts = _as_f32(pts, 3)
        leg = leg_array(leg)
        q = _as_f32(quat)
        out = np.zeros(len(pts), np.uint8)
        self.L.op_reach(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, out.ctypes.data, threads)
        return out

    def dist(self, pts, leg, quat=(1, 0, 0, 0), threads=1):
        pts = _as_f32(pts, 3)
        leg = leg_array(leg)
        q = _as_f32(quat)
        out = np.zeros_like(pts)
        fl = np.zeros(len(pts), np.uint8)
        self.L.op_dist(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, out.ctypes.data,
                       fl.ctypes.data, threads)
        return out, fl

    def find_region(self, x, y, leg):
        leg = leg_array(leg)
        return self.L.op_find_region(x, y, leg.ctypes.data)

    def insert_circles(self, x, y, leg):
        leg = leg_array(leg)
        out = np.zeros(16, np.float32)
        n = self.L.op_insert_circles(x, y, leg.ctypes.data, out.ctypes.data)
        return out.reshape(4, 4)[:n]

    def insert_intersec(self, leg):
        leg = leg_array(leg)
        out = np.zeros(20, np.float32)
        n = self.L.op_insert_intersec(leg.ctypes.data, out.ctypes.data)
        return out.reshape(10, 2)[:n]

    def rpy_to_quat(self, r, p, y):
        out = np.zeros(4, np.float32)
        self.L.op_rpy_to_quat_p(r, p, y, out.ctypes.data)
        return out

    def rotate_leg_data(self, quat, leg):
        leg = leg_array(leg)
        q = _as_f32(quat)
        out = np.zeros(14, np.float32)
        self.L.op_rotate_leg_data_p(q.ctypes.data, leg.ctypes.data, out.ctypes.data)
        return out

    def qt_rotate(self, quat, v):
        q, v = _as_f32(quat), _as_f32(v)
        out = np.zeros(3, np.float32)
        self.L.op_qt_rotate_p(q.ctypes.data, v.ctypes.data, out.ctypes.data)
        return out

    def full_struct_orientations(self):
        out = np.zeros((45, 4), np.float32)
        n = self.L.op_full_struct_orientations(out.ctypes.data)
        assert n == 45
        return out

    def quaternion_from_angle_index(self, idx):
        out = np.zeros(4, np.float32)
        self.L.op_quaternion_from_angle_index(idx, out.ctypes.data)
        return out

    def standability(self, bodies, targets, legs, quats, pre_cull=False, threads=1):
        bodies, targets = _as_f32(bodies, 3), _as_f32(targets, 3)
        legs = np.ascontiguousarray(np.stack([leg_array(l) for l in legs]), np.float32)
        quats = _as_f32(quats).reshape(-1, 4)
        out = np.zeros(len(bodies), np.uint8)
        self.L.op_standability(bodies.ctypes.data, len(bodies), targets.ctypes.data, len(targets),
                               legs.ctypes.data, len(legs), quats.ctypes.data, len(quats),
                               1 if pre_cull else 0, out.ctypes.data, threads)
        return out


    def validity_children(self, parent_box6, parent_validity, footholds, leg):
        """validity_child (several_leg_octree.cu:19-151) on the 8 children of one parent box:
        (flags 8 x [validity, leaf, raw, onEdge], boxes 8 x 6)."""
        f = _as_f32(footholds, 3)
        leg = leg_array(leg)
        box = _as_f32(parent_box6)
        flags = np.zeros((8, 4), np.uint8)
        boxes = np.zeros((8, 6), np.float32)
        self.L.op_validity_children(box.ctypes.data, int(parent_validity), f.ctypes.data, len(f), leg.ctypes.data,
                                    flags.ctypes.data, boxes.ctypes.data)
        return flags, boxes

    def apply_oct(self, footholds, leg, max_depth=1, cap=1 << 20):
        """Sequential restatement of apply_oct (several_leg_octree.cu:391-488): centres of the
        valid leaf / raw nodes after `max_depth` refinement passes."""
        f = _as_f32(footholds, 3)
        leg = leg_array(leg)
        out = np.zeros((cap, 3), np.float32)
        n = self.L.op_apply_oct(f.ctypes.data, len(f), leg.ctypes.data, max_depth, out.ctypes.data, cap)
        assert n <= cap
        return out[:n].copy()

    def apply_recurs(self, pts, leg, max_depth=1, quat=(1, 0, 0, 0), fill=-1.0):
        """Sequential restatement of apply_recurs: (leaf depth, 0, 0) per point; points outside the
        root box keep (fill, 0, 0)."""
        pts = _as_f32(pts, 3)
        leg = leg_array(leg)
        q = _as_f32(quat)
        out = np.zeros_like(pts)
        out[:, 0] = fill
        self.L.op_apply_recurs(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, max_depth,
                               out.ctypes.data)
        return out

    def create_child_box(self, parent6, child, small3=(0, 0, 0)):
        p = _as_f32(parent6)
        s = np.ascontiguousarray(small3, np.uint8)
        c = np.zeros(6, np.float32)
        missing = ctypes.c_int(0)
        r = self.L.op_create_child_box(p.ctypes.data, child, s.ctypes.data, c.ctypes.data, ctypes.byref(missing))
        return r, c, missing.value


class RefOracle(_Oracle):
    """The reference's own host code (ref_shim.cu)."""
    prefix = "ref_"
    path = REF_LIB
    kind = "reference"

    def __init__(self):
        super().__init__()
        L = self.L
        L.ref_get_leg.argtypes = [_ci, _cf, _vp]
        L.ref_reach.argtypes = [_vp, _sz, _vp, _vp, _vp, _ci]
        L.ref_dist.argtypes = [_vp, _sz, _vp, _vp, _vp, _vp, _ci]
        L.ref_apply_reach_cpu.argtypes = [_vp, _sz, _vp, _vp]
        L.ref_apply_reach_cpu.restype = ctypes.c_double
        L.ref_apply_dist_cpu.argtypes = [_vp, _sz, _vp, _vp]
        L.ref_apply_dist_cpu.restype = ctypes.c_double
        L.ref_find_region.argtypes = [_cf, _cf, _vp]
        L.ref_insert_circles.argtypes = [_cf, _cf, _vp, _vp]
        L.ref_insert_intersec.argtypes = [_vp, _vp]
        L.ref_qt_rotate.argtypes = [_vp, _vp, _vp]
        L.ref_rpy_to_quat.argtypes = [_cf, _cf, _cf, _vp]
        L.ref_rotate_leg_data.argtypes = [_vp, _vp, _vp]
        L.ref_qt_multiply.argtypes = [_vp, _vp, _vp]
        L.ref_quat_from_vect_angle.argtypes = [_vp, _cf, _vp]
        L.ref_create_child_box.argtypes = [_vp, ctypes.c_uint, _vp, _vp, _vp]
        L.ref_create_child_box.restype = ctypes.c_uint
        L.ref_is_in_box.argtypes = [_vp, _vp]

    def get_leg(self, robot, azimuth=0.0):
        out = np.zeros(14, np.float32)
        self.L.ref_get_leg(robot, azimuth, out.ctypes.data)
        return out

    def reach(self, pts, leg, quat=(1, 0, 0, 0), threads=1):
        pts = _as_f32(pts, 3)
        leg = leg_array(leg)
        q = _as_f32(quat)
        out = np.zeros(len(pts), np.uint8)
        self.L.ref_reach(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, out.ctypes.data, threads)
        return out

    def dist(self, pts, leg, quat=(1, 0, 0, 0), threads=1):
        pts = _as_f32(pts, 3)
        leg = leg_array(leg)
        q = _as_f32(quat)
        out = np.zeros_like(pts)
        fl = np.zeros(len(pts), np.uint8)
        self.L.ref_dist(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, out.ctypes.data,
                        fl.ctypes.data, threads)
        return out, fl

    def apply_reach_cpu(self, pts, leg):
        """The CPU path exactly as shipped (single thread, quatTest); returns (flags, ms)."""
        pts = _as_f32(pts, 3)
        leg = leg_array(leg)
        out = np.zeros(len(pts), np.uint8)
        ms = self.L.ref_apply_reach_cpu(pts.ctypes.data, len(pts), leg.ctypes.data, out.ctypes.data)
        return out, ms

    def apply_dist_cpu(self, pts, leg):
        pts = _as_f32(pts, 3)
        leg = leg_array(leg)
        out = np.zeros_like(pts)
        ms = self.L.ref_apply_dist_cpu(pts.ctypes.data, len(pts), leg.ctypes.data, out.ctypes.data)
        return out, ms

    def find_region(self, x, y, leg):
        leg = leg_array(leg)
        return self.L.ref_find_region(x, y, leg.ctypes.data)

    def insert_circles(self, x, y, leg):
        leg = leg_array(leg)
        out = np.zeros(16, np.float32)
        n = self.L.ref_insert_circles(x, y, leg.ctypes.data, out.ctypes.data)
        return out.reshape(4, 4)[:n]

    def insert_intersec(self, leg):
        leg = leg_array(leg)
        out = np.zeros(20, np.float32)
        n = self.L.ref_insert_intersec(leg.ctypes.data, out.ctypes.data)
        return out.reshape(10, 2)[:n]

    def rpy_to_quat(self, r, p, y):
        out = np.zeros(4, np.float32)
        self.L.ref_rpy_to_quat(r, p, y, out.ctypes.data)
        return out

    def rotate_leg_data(self, quat, leg):
        leg = leg_array(leg)
        q = _as_f32(quat)
        out = np.zeros(14, np.float32)
        self.L.ref_rotate_leg_data(q.ctypes.data, leg.ctypes.data, out.ctypes.data)
        return out

    def qt_rotate(self, quat, v):
        q, v = _as_f32(quat), _as_f32(v)
        out = np.zeros(3, np.float32)
        self.L.ref_qt_rotate(q.ctypes.data, v.ctypes.data, out.ctypes.data)
        return out

    def qt_multiply(self, a, b):
        a, b = _as_f32(a), _as_f32(b)
        out = np.zeros(4, np.float32)
        self.L.ref_qt_multiply(a.ctypes.data, b.ctypes.data, out.ctypes.data)
        return out

    def quat_from_vect_angle(self, axis, angle):
        axis = _as_f32(axis)
        out = np.zeros(4, np.float32)
        self.L.ref_quat_from_vect_angle(axis.ctypes.data, angle, out.ctypes.data)
        return out

    def create_child_box(self, parent6, child, small3=(0, 0, 0)):
        p = _as_f32(parent6)
        s = np.ascontiguousarray(small3, np.uint8)
        c = np.zeros(6, np.float32)
        missing = ctypes.c_int(0)
        r = self.L.ref_create_child_box(p.ctypes.data, child, s.ctypes.data, c.ctypes.data,
                                        ctypes.byref(missing))
        return r, c, missing.value


def have_ref():
    return os.path.exists(REF_LIB)


def best():
    """The strongest oracle available: the compiled reference if present, else the C port."""
    return RefOracle() if have_ref() else PortOracle()

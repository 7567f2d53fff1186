"""lrm_b200 — host-side mirror of the reference's plugin surface over the C ABI (include/lrm_c.h).

The directory name carries a hyphen, so import it through ``lrm_loader.load()`` (repo root) or
``importlib``; the module registers itself as ``lrm_b200``.

Everything here is a thin ctypes veneer: arguments are numpy arrays (host pointers, staged by the
library exactly like the reference's ``apply_kernel``, cross_compiled.cu:34-79) or torch CUDA
tensors (device pointers, asynchronous on the current torch stream).  There is NO CPU
implementation behind these calls: if ``liblrm_b200.so`` is missing or no CUDA device is usable the
call raises.

Names follow the reference: ``LegDimensions`` (HeaderCPP.h:19-52), ``get_M2_leg`` /
``get_moonbot_leg`` (static_variables.cpp:44-93), ``reachability`` / ``distance``
(one_leg_global.cu:74-130), ``robot_full_struct`` (several_leg.cu:796-877).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LRM_B200_LIB") or os.path.join(_HERE, "liblrm_b200.so")   # override: measurement builds

LRM_OK = 0


class LegDimensions(ctypes.Structure):
    """Field-for-field the reference's LegDimensions (14 floats)."""

    _fields_ = [(n, ctypes.c_float) for n in (
        "body_angle", "body", "coxa_pitch", "coxa_length", "tibia_length", "femur_length",
        "tibia_absolute_pos", "tibia_absolute_neg", "max_angle_coxa", "min_angle_coxa",
        "max_angle_tibia", "min_angle_tibia", "max_angle_femur", "min_angle_femur")]

    def as_array(self):
        return np.frombuffer(bytes(self), dtype=np.float32).copy()

    @classmethod
    def from_array(cls, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        assert a.size == 14
        return cls.from_buffer_copy(a.tobytes())


class PositOpts(ctypes.Structure):
    _fields_ = [("pre_cull", ctypes.c_int), ("first_hit_only", ctypes.c_int)]


class LrmError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded C-ABI library; raises (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LrmError(
                f"{LIB_PATH} not found: build it with `python {os.path.join(_HERE, 'build.py')}`"
                " (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp, sz, ci, fp = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_float)
        legp = ctypes.POINTER(LegDimensions)
        L.lrm_abi_version.restype = ci
        L.lrm_last_error.restype = ctypes.c_char_p
        L.lrm_device_count.restype = ci
        L.lrm_set_fast_path_min_points.restype = sz
        L.lrm_set_fast_path_min_points.argtypes = [sz]
        L.lrm_set_device.argtypes = [ci]
        L.lrm_set_option.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        L.lrm_get_stat.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]
        L.lrm_default_leg.argtypes = [ci, ctypes.c_float, legp]
        L.lrm_reach.argtypes = [vp, sz, legp, vp, vp, ci, vp, fp]
        L.lrm_dist.argtypes = [vp, sz, legp, vp, vp, vp, ci, vp, fp]
        L.lrm_reach_dist.argtypes = [vp, sz, legp, vp, vp, vp, ci, vp, fp]
        L.lrm_reach_dist_soa.argtypes = [vp, vp, vp, sz, legp, vp, vp, vp, vp, vp, vp, fp]
        L.lrm_forward_kine.argtypes = [vp, sz, legp, vp, ci, vp, fp]
        L.lrm_make_lattice.argtypes = [vp, vp, vp, vp, sz, sz, vp]
        L.lrm_full_struct_orientations.argtypes = [vp, ci]
        L.lrm_rpy_to_quat.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float, vp]
        L.lrm_positionability.argtypes = [vp, sz, vp, sz, legp, ci, vp, ci,
                                          ctypes.POINTER(PositOpts), vp, ci, vp, fp]
        L.lrm_positionability_counts.argtypes = [vp, sz, vp, sz, legp, ci, vp, ci, ctypes.POINTER(PositOpts), vp,
                                                 ctypes.POINTER(ctypes.c_double), ci, vp]
        L.lrm_recurs.argtypes = [vp, sz, legp, vp, ci, vp, ci, vp, fp]
        L.lrm_oct.argtypes = [vp, sz, legp, ci, vp, sz, ctypes.POINTER(ctypes.c_size_t), ci, vp, fp]
        L.lrm_oct_sharded.argtypes = [vp, sz, legp, ci, ci, ci, vp, sz, ctypes.POINTER(ctypes.c_size_t), vp,
                                      ci, vp, fp]
        L.lrm_oct_children.argtypes = [vp, sz, legp, vp, ci, vp, vp, ci, vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != LRM_OK:
        raise LrmError(f"lrm error {rc}: {lib().lrm_last_error().decode()}")


def set_fast_path_min_points(n):
    """One-leg sweeps of >= n points use the certified tables (default 4 Mi); returns the old value."""
    return int(lib().lrm_set_fast_path_min_points(int(n)))


def set_option(name, value):
    """lrm_set_option: tuning / measurement knobs (none changes a result); returns the old value."""
    prev = ctypes.c_double(0.0)
    _check(lib().lrm_set_option(name.encode(), float(value), ctypes.byref(prev)))
    return prev.value


def get_stat(name):
    v = ctypes.c_double(0.0)
    _check(lib().lrm_get_stat(name.encode(), ctypes.byref(v)))
    return v.value


def get_leg(robot, azimuth=0.0):
    leg = LegDimensions()
    _check(lib().lrm_default_leg(int(robot), float(azimuth), ctypes.byref(leg)))
    return leg


def get_moonbot_leg(azimuth=0.0):
    return get_leg(0, azimuth)


def get_M2_leg(azimuth=0.0):
    return get_leg(1, azimuth)


# ---- argument plumbing ----------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _quat_ptr(quat):
    if quat is None:
        return None, None
    q = np.ascontiguousarray(quat, dtype=np.float32).reshape(4)
    return q, q.ctypes.data


def _stream_ptr(stream):
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):
        return ctypes.c_void_p(stream.cuda_stream)
    return ctypes.c_void_p(int(stream))


def _stream_for(stream, tensor):
    """An explicitly passed stream is used as given (the default stream, 0, included); otherwise the
    current torch stream of the TENSOR's device."""
    if stream is not None:
        return _stream_ptr(stream)
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(tensor.device).cuda_stream)


class _device_of:
    """Make the device of `tensor` current for the duration of a device-pointer call (scratch
    buffers, cached tables and the launch itself belong to the current device) and check that every
    other tensor argument lives there too."""

    def __init__(self, tensor, *others):
        import torch
        for t in others:
            if t is not None and _is_torch(t):
                assert t.device == tensor.device, f"tensors on different devices: {t.device} vs {tensor.device}"
        self.guard = torch.cuda.device(tensor.device)

    def __enter__(self):
        return self.guard.__enter__()

    def __exit__(self, *a):
        return self.guard.__exit__(*a)


def _prep_points(points):
    """-> (on_device, pointer, n, keepalive)"""
    if _is_torch(points):
        import torch
        assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous()
        assert points.dim() == 2 and points.shape[1] == 3
        return 1, points.data_ptr(), points.shape[0], points
    a = np.ascontiguousarray(points, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 3
    return 0, a.ctypes.data, a.shape[0], a


def _timing(want):
    ms = ctypes.c_float(0.0)
    return ms, (ctypes.byref(ms) if want else None)


def reachability(points, leg, quat=None, out=None, stream=None, timing=False):
    """Per-point reachability flag (reachability_global, one_leg_global.cu:106-130).

    numpy in -> numpy uint8 out (host staging); torch.cuda in -> torch.uint8 out (device)."""
    dev, ptr, n, keep = _prep_points(points)
    q, qp = _quat_ptr(quat)
    ms, msp = _timing(timing)
    if dev:
        import torch
        out = torch.empty(n, dtype=torch.uint8, device=points.device) if out is None else out
        st = _stream_for(stream, points)
        with _device_of(points, out):
            _check(lib().lrm_reach(ptr, n, ctypes.byref(leg), qp, out.data_ptr(), 1, st, msp))
    else:
        out = np.empty(n, dtype=np.uint8) if out is None else out
        _check(lib().lrm_reach(ptr, n, ctypes.byref(leg), qp, out.ctypes.data, 0, None, msp))
    return (out, ms.value) if timing else out


def distance(points, leg, quat=None, out=None, flags=True, stream=None, timing=False):
    """Vector to the reachability edge + distance_global's bool (one_leg_global.cu:74-101)."""
    dev, ptr, n, keep = _prep_points(points)
    q, qp = _quat_ptr(quat)
    ms, msp = _timing(timing)
    if dev:
        import torch
        out = torch.empty((n, 3), dtype=torch.float32, device=points.device) if out is None else out
        fl = torch.empty(n, dtype=torch.uint8, device=points.device) if flags else None
        st = _stream_for(stream, points)
        with _device_of(points, out, fl):
            _check(lib().lrm_dist(ptr, n, ctypes.byref(leg), qp, out.data_ptr(),
                                  fl.data_ptr() if flags else None, 1, st, msp))
    else:
        out = np.empty((n, 3), dtype=np.float32) if out is None else out
        fl = np.empty(n, dtype=np.uint8) if flags else None
        _check(lib().lrm_dist(ptr, n, ctypes.byref(leg), qp, out.ctypes.data,
                              fl.ctypes.data if flags else None, 0, None, msp))
    res = (out, fl) if flags else (out,)
    return res + (ms.value,) if timing else (res if flags else out)


def reach_dist(points, leg, quat=None, out_flags=None, out_vec=None, stream=None, timing=False):
    """Fused pass: (reachability flags, distance vectors)."""
    dev, ptr, n, keep = _prep_points(points)
    q, qp = _quat_ptr(quat)
    ms, msp = _timing(timing)
    if dev:
        import torch
        fl = torch.empty(n, dtype=torch.uint8, device=points.device) if out_flags is None else out_flags
        vec = torch.empty((n, 3), dtype=torch.float32, device=points.device) if out_vec is None else out_vec
        st = _stream_for(stream, points)
        with _device_of(points, fl, vec):
            _check(lib().lrm_reach_dist(ptr, n, ctypes.byref(leg), qp, fl.data_ptr(), vec.data_ptr(), 1,
                                        st, msp))
    else:
        fl = np.empty(n, dtype=np.uint8) if out_flags is None else out_flags
        vec = np.empty((n, 3), dtype=np.float32) if out_vec is None else out_vec
        _check(lib().lrm_reach_dist(ptr, n, ctypes.byref(leg), qp, fl.ctypes.data, vec.ctypes.data,
                                    0, None, msp))
    return (fl, vec, ms.value) if timing else (fl, vec)


def reach_dist_soa(x, y, z, leg, quat=None, want_vec=True, stream=None, timing=False):
    """SoA planes (torch CUDA tensors) in, (flags, dx, dy, dz) out."""
    import torch
    n = x.shape[0]
    q, qp = _quat_ptr(quat)
    ms, msp = _timing(timing)
    fl = torch.empty(n, dtype=torch.uint8, device=x.device)
    d = [torch.empty(n, dtype=torch.float32, device=x.device) for _ in range(3)] if want_vec else [None] * 3
    st = _stream_for(stream, x)
    with _device_of(x, y, z):
        _check(lib().lrm_reach_dist_soa(x.data_ptr(), y.data_ptr(), z.data_ptr(), n, ctypes.byref(leg), qp,
                                        fl.data_ptr(), *[t.data_ptr() if t is not None else None for t in d],
                                        st, msp))
    res = (fl, d[0], d[1], d[2])
    return res + (ms.value,) if timing else res


def forward_kinematics(angles, leg, stream=None):
    """(coxa, femur, tibia) -> xyz (forward_kine_kernel, one_leg.cu:377-414)."""
    dev, ptr, n, keep = _prep_points(angles)
    if dev:
        import torch
        out = torch.empty((n, 3), dtype=torch.float32, device=angles.device)
        st = _stream_for(stream, angles)
        with _device_of(angles):
            _check(lib().lrm_forward_kine(ptr, n, ctypes.byref(leg), out.data_ptr(), 1, st, None))
    else:
        out = np.empty((n, 3), dtype=np.float32)
        _check(lib().lrm_forward_kine(ptr, n, ctypes.byref(leg), out.ctypes.data, 0, None, None))
    return out


def lattice_spec(lo, hi, dims):
    """float32 (lo, step, dims) of the synthetic lattice: coordinate = lo + float(i) * step."""
    lo = np.asarray(lo, dtype=np.float32)
    hi = np.asarray(hi, dtype=np.float32)
    dims = np.asarray(dims, dtype=np.uint32)
    den = np.maximum(dims.astype(np.float32) - np.float32(1), np.float32(1))
    step = ((hi - lo) / den).astype(np.float32)
    return lo, step, dims


def make_lattice(out, lo, step, dims, first=0, count=None, stream=None):
    """Fill a torch CUDA tensor (count x 3) with lattice points [first, first+count)."""
    lo = np.ascontiguousarray(lo, dtype=np.float32)
    step = np.ascontiguousarray(step, dtype=np.float32)
    dims = np.ascontiguousarray(dims, dtype=np.uint32)
    count = out.shape[0] if count is None else count
    st = _stream_for(stream, out)
    with _device_of(out):
        _check(lib().lrm_make_lattice(out.data_ptr(), lo.ctypes.data, step.ctypes.data, dims.ctypes.data,
                                      int(first), int(count), st))
    return out


def lattice_host(lo, step, dims, first=0, count=None):
    """numpy twin of make_lattice (same two float32 operations per coordinate)."""
    lo = np.asarray(lo, np.float32)
    step = np.asarray(step, np.float32)
    dims = [int(d) for d in dims]
    total = dims[0] * dims[1] * dims[2]
    count = total - first if count is None else count
    i = np.arange(first, first + count, dtype=np.int64)
    iz = i % dims[2]
    t = i // dims[2]
    iy = t % dims[1]
    ix = t // dims[1]
    out = np.empty((count, 3), np.float32)
    for k, idx in enumerate((ix, iy, iz)):
        out[:, k] = lo[k] + idx.astype(np.float32) * step[k]
    return out


def full_struct_orientations():
    """The 45 body orientations of robot_full_struct (several_leg.cu:811-857), (45, 4) float32."""
    q = np.empty((45, 4), dtype=np.float32)
    _check(lib().lrm_full_struct_orientations(q.ctypes.data, 45))
    return q


def rpy_to_quat(roll, pitch, yaw):
    """RPYtoQuat (octree_util.cu.h:164-172) in the storage qtRotate expects, (4,) float32."""
    q = np.empty(4, dtype=np.float32)
    _check(lib().lrm_rpy_to_quat(float(roll), float(pitch), float(yaw), q.ctypes.data))
    return q


def yaw_orientations(n_yaw, yaw_min=0.0, yaw_max=2.0 * np.pi):
    """n_yaw level orientations with uniformly spaced yaw in [yaw_min, yaw_max) (BASELINE configs[4]:
    dense body-pose x yaw grid)."""
    ys = np.float32(yaw_min) + (np.float32(yaw_max) - np.float32(yaw_min)) * (
        np.arange(n_yaw, dtype=np.float32) / np.float32(n_yaw))
    return np.stack([rpy_to_quat(0.0, 0.0, y) for y in ys]).astype(np.float32)


def positionability(bodies, map_points, legs, quats=None, pre_cull=False, stream=None, timing=False):
    """standable[b] = 1 + index of the first orientation under which every leg finds a reachable
    map point and the cull cylinders pass (several_leg.cu:762-787); 0 otherwise."""
    devb, pb, nb, kb = _prep_points(bodies)
    devm, pm, nt, km = _prep_points(map_points)
    assert devb == devm, "bodies and map must both be host or both be device"
    quats = full_struct_orientations() if quats is None else np.ascontiguousarray(quats, np.float32)
    quats = quats.reshape(-1, 4)
    leg_arr = (LegDimensions * len(legs))(*legs)
    opts = PositOpts(1 if pre_cull else 0, 0)
    ms, msp = _timing(timing)
    if devb:
        import torch
        out = torch.empty(nb, dtype=torch.uint8, device=bodies.device)
        st = _stream_for(stream, bodies)
        with _device_of(bodies, map_points):
            _check(lib().lrm_positionability(pb, nb, pm, nt, leg_arr, len(legs), quats.ctypes.data,
                                             quats.shape[0], ctypes.byref(opts), out.data_ptr(), 1, st, msp))
    else:
        out = np.empty(nb, dtype=np.uint8)
        _check(lib().lrm_positionability(pb, nb, pm, nt, leg_arr, len(legs), quats.ctypes.data,
                                         quats.shape[0], ctypes.byref(opts), out.ctypes.data, 0, None, msp))
    return (out, ms.value) if timing else out


def positionability_counts(bodies, map_points, legs, quats=None, pre_cull=False, stream=None):
    """lrm_positionability_counts on device tensors: (standable, {"leg_predicates_executed",
    "cylinder_predicates_executed", "leg_predicates_algorithmic"})."""
    import torch
    devb, pb, nb, kb = _prep_points(bodies)
    devm, pm, nt, km = _prep_points(map_points)
    assert devb == 1 and devm == 1, "device tensors only"
    quats = full_struct_orientations() if quats is None else np.ascontiguousarray(quats, np.float32)
    quats = quats.reshape(-1, 4)
    leg_arr = (LegDimensions * len(legs))(*legs)
    opts = PositOpts(1 if pre_cull else 0, 0)
    out = torch.empty(nb, dtype=torch.uint8, device=bodies.device)
    counts = (ctypes.c_double * 3)()
    st = _stream_for(stream, bodies)
    with _device_of(bodies, map_points):
        _check(lib().lrm_positionability_counts(pb, nb, pm, nt, leg_arr, len(legs), quats.ctypes.data, quats.shape[0],
                                                ctypes.byref(opts), out.data_ptr(), counts, 1, st))
    return out, {"leg_predicates_executed": counts[0], "cylinder_predicates_executed": counts[1],
                 "leg_predicates_algorithmic": counts[2]}


def apply_recurs(points, leg, max_depth=1, quat=None, fill=-1.0, stream=None):
    """Octree of the single-leg distance field painted on the query points (apply_recurs,
    cross_compiled.cu:82-139): (leaf depth, 0, 0) per point, (fill, 0, 0) outside the root box."""
    dev, ptr, n, keep = _prep_points(points)
    q, qp = _quat_ptr(quat)
    if dev:
        import torch
        out = torch.zeros((n, 3), dtype=torch.float32, device=points.device)
        out[:, 0] = fill
        st = _stream_for(stream, points)
        with _device_of(points):
            _check(lib().lrm_recurs(ptr, n, ctypes.byref(leg), qp, int(max_depth), out.data_ptr(), 1, st, None))
    else:
        out = np.zeros((n, 3), dtype=np.float32)
        out[:, 0] = fill
        _check(lib().lrm_recurs(ptr, n, ctypes.byref(leg), qp, int(max_depth), out.ctypes.data, 0, None, None))
    return out


class _maybe_device:
    def __init__(self, dev, tensor):
        self.ctx = _device_of(tensor) if dev else None

    def __enter__(self):
        return self.ctx.__enter__() if self.ctx else None

    def __exit__(self, *a):
        return self.ctx.__exit__(*a) if self.ctx else False


def apply_oct(footholds, leg, max_depth=1, cap=1 << 16, stream=None, timing=False, shard=0, nshards=1,
              child_counts=False):
    """Body-space octree positionability (apply_oct, several_leg_octree.cu:391-488): centres of the
    valid leaf / raw nodes after `max_depth` refinement passes, (n, 3) float32 numpy array.
    shard / nshards: refine only the root's children c with c % nshards == shard (lrm_oct_sharded);
    child_counts=True also returns the number of centres under each of the 8 top-level children."""
    dev, ptr, nt, keep = _prep_points(footholds)
    ms, msp = _timing(timing)
    counts = (ctypes.c_size_t * 8)()
    while True:
        out = np.empty((cap, 3), dtype=np.float32)
        count = ctypes.c_size_t(0)
        st = _stream_for(stream, footholds) if dev else None
        with _maybe_device(dev, footholds):
            _check(lib().lrm_oct_sharded(ptr, nt, ctypes.byref(leg), int(max_depth), int(shard), int(nshards),
                                         out.ctypes.data, cap, ctypes.byref(count), counts, dev, st, msp))
        if count.value <= cap:
            res = (out[:count.value].copy(),)
            if child_counts:
                res += (np.array(list(counts), dtype=np.int64),)
            if timing:
                res += (ms.value,)
            return res if len(res) > 1 else res[0]
        cap = count.value


def merge_oct_shards(parts):
    """parts[r] = (centres, child_counts) of shard r of len(parts): the full list in the reference's
    traversal order (top-level child c comes from shard c % nshards)."""
    n = len(parts)
    offs = [np.concatenate([[0], np.cumsum(c)]) for _, c in parts]
    out = [parts[c % n][0][offs[c % n][c]:offs[c % n][c + 1]] for c in range(8)]
    return np.concatenate(out) if out else np.zeros((0, 3), np.float32)


def oct_children(footholds, leg, parent_box6, parent_validity=False, stream=None):
    """One launch of validity_child (several_leg_octree.cu:19-151) on the 8 children of a body box:
    (flags 8 x [validity, leaf, raw, onEdge] uint8, boxes 8 x 6 float32)."""
    dev, ptr, nt, keep = _prep_points(footholds)
    box = np.ascontiguousarray(parent_box6, dtype=np.float32).reshape(6)
    flags = np.zeros((8, 4), np.uint8)
    boxes = np.zeros((8, 6), np.float32)
    st = _stream_for(stream, footholds) if dev else None
    with _maybe_device(dev, footholds):
        _check(lib().lrm_oct_children(ptr, nt, ctypes.byref(leg), box.ctypes.data, int(bool(parent_validity)),
                                      flags.ctypes.data, boxes.ctypes.data, dev, st))
    return flags, boxes


def robot_full_struct(body_map, target_map, legs):
    """Drop-in shape of the reference's robot_full_struct (several_leg.cu:796-877): returns the
    standable body positions (original order; the reference's order depends on thrust::partition)
    and the dummy count array of 3s (several_leg.cu:867-868)."""
    flags = positionability(body_map, target_map, legs, pre_cull=True)
    if _is_torch(flags):
        sel = flags != 0
        bodies = body_map[sel]
        import torch
        return bodies, torch.full((bodies.shape[0],), 3, dtype=torch.int32, device=bodies.device)
    sel = flags != 0
    bodies = np.asarray(body_map, dtype=np.float32)[sel]
    return bodies, np.full(bodies.shape[0], 3, dtype=np.int32)

"""Second parity / timing column: the reference's OWN GPU entry points, compiled unmodified for
sm_100 with its shipped device flags (oracle/_ref/libref_gpu*.so, `make -C oracle refgpu`), run on
the same B200 next to this repo's kernels.

    python tools/vs_refgpu.py [--points 20000000] [--json gpurun_out/vs_refgpu.json]

Sections (each guarded: a Python-level failure is reported, not fatal; robot_full_struct runs first, in
a context no other allocator has touched — with torch's caching allocator active before it, the
reference's pipeline died with an illegal address inside thrust::partition):
  one_leg   apply_kernel(reachability_global_kernel / distance_global_kernel)   vs lrm_reach / lrm_dist
  full      robot_full_struct (several_leg.cu:796-877)                          vs lrm_positionability(pre_cull)
            and vs the CPU restatement (oracle_port.c op_standability) -> pins the pipeline logic
  recurs    apply_recurs                                                        vs lrm_recurs
  oct       apply_oct (MAX_DEPTH 1 as shipped)                                  vs lrm_oct(max_depth=1)
The reference's device code is built with -use_fast_math (CMakeLists.txt:146), so its flags can
differ from its own CPU path on points next to the reachability edge; such points are counted
with their distance to the edge.
"""
import argparse, ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrm_loader
from tests import terrain

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=20_000_000)
ap.add_argument("--json", default=None)
ap.add_argument("--skip", nargs="*", default=["oct"],
                help="sections to skip (apply_oct of the reference did not terminate within 10 min on sm_100)")
args = ap.parse_args()

lrm = lrm_loader.load()
REF = os.path.join(ROOT, "oracle", "_ref")
sz, vp, fp = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_float
out = {}


def section(name):
    def deco(fn):
        if name in args.skip:
            return fn
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:  # report and continue with the other sections
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        out[name]["section_s"] = round(time.perf_counter() - t0, 2)
        print(name, json.dumps(out[name]), flush=True)
        return fn
    return deco


def load_ref():
    g = ctypes.CDLL(os.path.join(REF, "libref_gpu.so"))
    g.refgpu_reach.restype = fp
    g.refgpu_reach.argtypes = [vp, sz, vp, vp]
    g.refgpu_dist.restype = fp
    g.refgpu_dist.argtypes = [vp, sz, vp, vp]
    g.refgpu_recurs.restype = fp
    g.refgpu_recurs.argtypes = [vp, sz, vp, vp]
    g.refgpu_oct.restype = fp
    g.refgpu_oct.argtypes = [vp, sz, vp, vp, sz, ctypes.POINTER(sz)]
    return g


def bench_points(n):
    """every k-th point of the 1e9 bench lattice (same extents as bench.py)"""
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (1000, 1000, 1000))
    stride = 10 ** 9 // n
    i = np.arange(n, dtype=np.int64) * stride
    iz, t = i % 1000, i // 1000
    iy, ix = t % 1000, t // 1000
    return np.stack([lo[0] + ix.astype(np.float32) * step[0], lo[1] + iy.astype(np.float32) * step[1],
                     lo[2] + iz.astype(np.float32) * step[2]], 1).astype(np.float32)


def _full_struct_call(libname, bodies, terr, la):
    s = ctypes.CDLL(os.path.join(REF, libname))
    s.refgpu_full_struct.restype = ctypes.c_double
    s.refgpu_full_struct.argtypes = [vp, sz, vp, sz, vp, sz, vp, sz, ctypes.POINTER(sz)]
    cap = len(bodies)
    ref_xyz = np.empty((cap, 3), np.float32)
    cnt = sz(0)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)  # the reference prints a line per orientation
    try:
        ms = s.refgpu_full_struct(bodies.ctypes.data, len(bodies), terr.ctypes.data, len(terr),
                                  la.ctypes.data, 4, ref_xyz.ctypes.data, cap, ctypes.byref(cnt))
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return ref_xyz[:cnt.value].copy(), ms


@section("full")
def _full():
    from oracle.oracle import PortOracle
    port = PortOracle()
    res = {}
    legs = [lrm.get_M2_leg(float(np.float32(k) * np.float32(np.pi / 2))) for k in range(4)]
    la = np.stack([l.as_array() for l in legs]).astype(np.float32)
    quats = lrm.full_struct_orientations()
    cases = {
        # SURVEY §8c restatement-test shape: sinusoidal terrain, 6912 poses
        "sine_16k_x_6912": (terrain.sine_terrain(128, 2400.0, 100.0), (24, 24, 12)),
        # a 128^2 window of the C3 Perlin generator
        "perlin_16k_x_20k": (terrain.perlin_terrain(128), (32, 40, 16)),
    }
    dump = {}
    for name, (terr, poses) in cases.items():
        bodies = terrain.body_lattice(terr, *poses)
        key = lambda a: {tuple(r) for r in np.ascontiguousarray(a, np.float32).view(np.uint32).reshape(-1, 3).tolist()}
        ref_fast, ms = _full_struct_call("libref_gpu_several.so", bodies, terr, la)
        ref_prec, ms_p = _full_struct_call("libref_gpu_several_precise.so", bodies, terr, la)
        t0 = time.perf_counter()
        got = lrm.positionability(bodies, terr, legs, quats, pre_cull=True)
        b200_ms = (time.perf_counter() - t0) * 1e3
        want = port.standability(bodies, terr, [l for l in la], quats, pre_cull=True, threads=os.cpu_count() or 1)
        sets = {"ref_gpu": key(ref_fast), "ref_gpu_precise": key(ref_prec), "b200": key(bodies[got != 0]),
                "cpu_oracle": key(bodies[want != 0])}
        names = list(sets)
        res[name] = {"map_points": len(terr), "poses": len(bodies),
                     "standable": {k: len(v) for k, v in sets.items()},
                     "symdiff": {f"{a}^{b}": len(sets[a] ^ sets[b]) for i, a in enumerate(names) for b in names[i + 1:]},
                     "ref_gpu_wall_ms": ms, "ref_gpu_precise_wall_ms": ms_p, "b200_wall_ms_host_pointers": b200_ms,
                     "ref_gpu_poses_per_s": len(bodies) / ms * 1e3, "b200_poses_per_s": len(bodies) / b200_ms * 1e3}
        # membership of every pose on which any two columns disagree, for offline analysis
        allk = set().union(*sets.values())
        bad = [k for k in allk if len({k in v for v in sets.values()}) > 1]
        if bad:
            arr = np.array(bad, np.uint32).view(np.float32).reshape(-1, 3)
            dump[name + "_pose"] = arr
            dump[name + "_member"] = np.array([[k in sets[n] for n in names] for k in bad], np.uint8)
        dump[name + "_b200_codes"] = got
        dump[name + "_oracle_codes"] = want
    if args.json:
        np.savez_compressed(os.path.splitext(args.json)[0] + "_full_dump.npz", columns=np.array(names), **dump)
    return res


@section("one_leg")
def _one_leg():
    g = load_ref()
    from oracle.oracle import best
    cpu = best()
    res = {}
    n = args.points
    pts = bench_points(n)
    for robot, name in ((1, "M2"), (0, "moonbot")):
        leg = lrm.get_leg(robot, 0.0)
        la = leg.as_array()
        r_ref = np.empty(n, np.uint8)
        d_ref = np.empty((n, 3), np.float32)
        ms_r = min(g.refgpu_reach(pts.ctypes.data, n, la.ctypes.data, r_ref.ctypes.data) for _ in range(3))
        ms_d = min(g.refgpu_dist(pts.ctypes.data, n, la.ctypes.data, d_ref.ctypes.data) for _ in range(3))
        # device-resident, one launch, kernel-only time: what apply_kernel's return value measures
        import torch
        d_pts = torch.from_numpy(pts).cuda()
        # steady state: the first calls of a new leg build its tables (the choice volume in the
        # background, which slows the foreground sweeps of those ~100 ms: tools/first_calls.py)
        lrm.reachability(d_pts, leg)
        lrm.distance(d_pts, leg)
        torch.cuda.synchronize()
        time.sleep(0.6)
        r_us, t_r = None, 1e30
        for _ in range(4):
            r_us, t = lrm.reachability(d_pts, leg, timing=True)
            t_r = min(t_r, t)
        d_us, t_d = None, 1e30
        for _ in range(4):
            d_us, _f, t = lrm.distance(d_pts, leg, timing=True)
            t_d = min(t_d, t)
        r_us, d_us = r_us.cpu().numpy(), d_us.cpu().numpy()
        del d_pts
        flag_diff = np.flatnonzero(r_ref != r_us)
        err = np.abs(d_ref - d_us).max(axis=1)
        bad = np.flatnonzero(err > 1e-2)
        # judge every disagreement against the reference's own CPU path (the parity oracle)
        sub = np.unique(np.concatenate([flag_diff[:20000], bad[:20000]]))
        cpu_r = cpu.reach(pts[sub], la, threads=8) if len(sub) else np.zeros(0, np.uint8)
        cpu_d = cpu.dist(pts[sub], la, threads=8)[0] if len(sub) else np.zeros((0, 3), np.float32)
        pos = {int(k): j for j, k in enumerate(sub)}
        fd = [pos[int(k)] for k in flag_diff[:20000]]
        bd = [pos[int(k)] for k in bad[:20000]]
        res[name] = {
            "points": n,
            "ref_gpu_reach_ms": ms_r, "ref_gpu_dist_ms": ms_d, "b200_reach_ms": t_r, "b200_dist_ms": t_d,
            "ref_gpu_reach_gpts": n / ms_r / 1e6, "ref_gpu_dist_gpts": n / ms_d / 1e6,
            "b200_reach_gpts": n / t_r / 1e6, "b200_dist_gpts": n / t_d / 1e6,
            "speedup_reach": ms_r / t_r, "speedup_dist": ms_d / t_d,
            "flags_differ_vs_ref_gpu": int(len(flag_diff)),
            "of_those_b200_equals_ref_cpu": int((cpu_r[fd] == r_us[flag_diff[:20000]]).sum()) if len(fd) else 0,
            "of_those_ref_gpu_equals_ref_cpu": int((cpu_r[fd] == r_ref[flag_diff[:20000]]).sum()) if len(fd) else 0,
            "vectors_differ_gt_1e-2mm_vs_ref_gpu": int(len(bad)),
            "of_those_b200_within_1e-2_of_ref_cpu": int((np.abs(cpu_d[bd] - d_us[bad[:20000]]).max(axis=1) <= 1e-2).sum()) if len(bd) else 0,
            "of_those_ref_gpu_within_1e-2_of_ref_cpu": int((np.abs(cpu_d[bd] - d_ref[bad[:20000]]).max(axis=1) <= 1e-2).sum()) if len(bd) else 0,
            "max_abs_vector_diff_mm": float(err.max()), "median_abs_vector_diff_mm": float(np.median(err)),
            "reachable_ref_gpu": int(r_ref.sum()), "reachable_b200": int(r_us.sum()),
        }
    return res


@section("recurs")
def _recurs():
    g = load_ref()
    leg = lrm.get_M2_leg(0.0)
    la = leg.as_array()
    rng = np.random.default_rng(5)
    pts = (rng.random((200_000, 3), dtype=np.float32) * np.float32(1400) - np.float32(700)).astype(np.float32)
    ref = np.zeros((len(pts), 3), np.float32)
    ref[:, 0] = -1.0
    ms = g.refgpu_recurs(pts.ctypes.data, len(pts), la.ctypes.data, ref.ctypes.data)
    got = lrm.apply_recurs(pts, leg, max_depth=1, fill=-1.0)
    hist_ref = {str(int(k)): int(v) for k, v in zip(*np.unique(ref[:, 0], return_counts=True))}
    hist_us = {str(int(k)): int(v) for k, v in zip(*np.unique(got[:, 0], return_counts=True))}
    return {"points": len(pts), "ref_gpu_ms": ms, "differ": int((ref[:, 0] != got[:, 0]).sum()),
            "depth_hist_ref_gpu": hist_ref, "depth_hist_b200": hist_us}


@section("oct")
def _oct():
    g = load_ref()
    from oracle.oracle import PortOracle
    port = PortOracle()
    res = {}
    terr = terrain.sine_terrain(25, 1200.0, 80.0)
    wide = port.get_leg(1, 0.0).copy()
    wide[8], wide[9] = 3.0, -3.0
    for name, la in (("M2_as_shipped", port.get_leg(1, 0.0)), ("wide_coxa", wide)):
        la = np.ascontiguousarray(la, np.float32)
        ref_xyz = np.empty((4096, 3), np.float32)
        cnt = sz(0)
        ms = g.refgpu_oct(terr.ctypes.data, len(terr), la.ctypes.data, ref_xyz.ctypes.data, 4096, ctypes.byref(cnt))
        got = lrm.apply_oct(terr, lrm.LegDimensions.from_array(la), max_depth=1)
        want = port.apply_oct(terr, la, 1)
        res[name] = {"ref_gpu_valid_boxes": int(cnt.value), "b200_valid_boxes": int(len(got)),
                     "cpu_oracle_valid_boxes": int(len(want)), "ref_gpu_ms": ms,
                     "same_list_b200_vs_ref_gpu": bool(cnt.value == len(got) and np.array_equal(ref_xyz[:cnt.value], got))}
    return res


if args.json:
    os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
    with open(args.json, "w") as f:
        json.dump(out, f, indent=1)
print(json.dumps(out))

"""The reference's own size sweep (bench.cpp:109-171): y = 0 slice, x in [-100, 601], z in [-100, 51]
(its z range starts at XMin, bench.cpp:113), pitch 0.04 mm * 2^k up to 50 mm -> N = 72 ... 66 160 650,
float-accumulated arange (bench.cpp:21-27), M2 leg.  Per N and per kernel (reachability_global /
distance_global) the kernel-only time — what apply_kernel returns and bench.cpp writes as
`N;ns_per_point` — for

  b200_host    lrm_reach / lrm_dist with HOST pointers (the apply_kernel contract: staging outside
               the timer, kernel_ms summed over the pipeline's chunks)
  b200_device  the same entry points on a device-resident array (one launch)
  ref_gpu      the reference's own kernels recompiled for sm_100 with its shipped flags
               (oracle/_ref/libref_gpu.so), through its own apply_kernel, same box, same process

    python tools/size_sweep.py [--json gpurun_out/size_sweep.json] [--csv-dir gpurun_out/size_sweep] [--max-n N]

CSV files use the reference's format (`N;ns_per_point`, one row per repetition, largest N first) so
that benchIllu.py-style plots read them next to bdata/pc/*.csv.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrm_loader  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--json", default=None)
ap.add_argument("--csv-dir", default=None)
ap.add_argument("--max-n", type=int, default=70_000_000)
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--no-ref", action="store_true")
args = ap.parse_args()

import torch  # noqa: E402

lrm = lrm_loader.load()
leg = lrm.get_M2_leg(0.0)
la = leg.as_array()


def arange32(start, end, step):
    out, v, step = [], np.float32(start), np.float32(step)
    while v <= np.float32(end):
        out.append(v)
        v = np.float32(v + step)
    return np.array(out, np.float32)


def grid(pix):
    xs, ys, zs = arange32(-100, 601, pix), arange32(0, 0, pix), arange32(-100, 51, pix)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    return np.ascontiguousarray(np.stack([X, Y, Z], -1).reshape(-1, 3), np.float32)


ref = None
if not args.no_ref:
    path = os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so")
    if os.path.exists(path):
        ref = ctypes.CDLL(path)
        vp, sz, fp = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float
        ref.refgpu_reach.restype = fp
        ref.refgpu_reach.argtypes = [vp, sz, vp, vp]
        ref.refgpu_dist.restype = fp
        ref.refgpu_dist.argtypes = [vp, sz, vp, vp]

rows = []
pix = 0.04
pixes = []
while pix <= 50:
    pixes.append(pix)
    pix *= 2
for pix in pixes:          # largest N first, like the reference's files
    pts = grid(pix)
    n = len(pts)
    if n > args.max_n:
        continue
    row = {"N": n, "pix_mm": pix}
    flags = np.empty(n, np.uint8)
    vec = np.empty((n, 3), np.float32)
    d_pts = torch.from_numpy(pts).cuda()
    d_flags = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_vec = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    reps = args.reps if n < 20_000_000 else 3

    def timed(fn):
        fn()                                   # warm-up (tables of the leg are built on the first large call)
        return [float(fn()) for _ in range(reps)]

    row["b200_host_reach_ms"] = timed(lambda: lrm.reachability(pts, leg, out=flags, timing=True)[1])
    row["b200_host_dist_ms"] = timed(lambda: lrm.distance(pts, leg, out=vec, flags=False, timing=True)[-1])
    row["b200_device_reach_ms"] = timed(lambda: lrm.reachability(d_pts, leg, out=d_flags, timing=True)[1])
    row["b200_device_dist_ms"] = timed(lambda: lrm.distance(d_pts, leg, out=d_vec, flags=False, timing=True)[-1])
    if ref is not None:
        r_flags = np.empty(n, np.uint8)
        r_vec = np.empty((n, 3), np.float32)
        row["ref_gpu_reach_ms"] = timed(lambda: ref.refgpu_reach(pts.ctypes.data, n, la.ctypes.data, r_flags.ctypes.data))
        row["ref_gpu_dist_ms"] = timed(lambda: ref.refgpu_dist(pts.ctypes.data, n, la.ctypes.data, r_vec.ctypes.data))
        row["flags_differ_vs_ref_gpu"] = int((r_flags != flags).sum())
    for k in list(row):
        if k.endswith("_ms"):
            row[k[:-3] + "_ns_per_point"] = float(np.median(row[k])) / n * 1e6
    for mode in ("reach", "dist"):
        if ref is not None:
            row[f"speedup_{mode}_host"] = row[f"ref_gpu_{mode}_ns_per_point"] / row[f"b200_host_{mode}_ns_per_point"]
            row[f"speedup_{mode}_device"] = row[f"ref_gpu_{mode}_ns_per_point"] / row[f"b200_device_{mode}_ns_per_point"]
    rows.append(row)
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items() if not k.endswith("_ms")}),
          flush=True)
    del d_pts, d_flags, d_vec

if args.json:
    os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
    with open(args.json, "w") as f:
        json.dump(rows, f, indent=1)
if args.csv_dir:
    os.makedirs(args.csv_dir, exist_ok=True)
    for col, name in (("b200_host_reach_ms", "rgpu.csv"), ("b200_host_dist_ms", "dgpu.csv"),
                      ("b200_device_reach_ms", "rgpu_device.csv"), ("b200_device_dist_ms", "dgpu_device.csv"),
                      ("ref_gpu_reach_ms", "rgpu_reference_sm100.csv"), ("ref_gpu_dist_ms", "dgpu_reference_sm100.csv")):
        if rows and col in rows[0]:
            with open(os.path.join(args.csv_dir, name), "w") as f:
                for r in rows:
                    for ms in r[col]:
                        f.write(f"{r['N']};{ms / r['N'] * 1e6:g}\n")

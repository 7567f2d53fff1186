"""Positionability benchmark (BASELINE configs[2] shape): Perlin terrain map, pose lattice, 4 M2
legs at k*pi/2, the 45 orientations of robot_full_struct.  Prints one JSON line (body poses/s).

    python tools/bench_posit.py [--map 1024] [--poses 64 128 32] [--check 4000] [--pre-cull]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
from tests import terrain

ap = argparse.ArgumentParser()
ap.add_argument("--map", type=int, default=1024)
ap.add_argument("--poses", type=int, nargs=3, default=[64, 128, 32])
ap.add_argument("--check", type=int, default=3000, help="poses to verify against the CPU oracle")
ap.add_argument("--pre-cull", action="store_true")
ap.add_argument("--legs", type=int, default=4)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

lrm = lrm_loader.load()
terr = terrain.perlin_terrain(args.map)
bodies = terrain.body_lattice(terr, *args.poses)
legs = [lrm.get_M2_leg(float(np.float32(k) * np.float32(2 * np.pi) / np.float32(args.legs))) for k in range(args.legs)]
quats = lrm.full_struct_orientations()
d_terr, d_bod = torch.from_numpy(terr).cuda(), torch.from_numpy(bodies).cuda()
out, ms = lrm.positionability(d_bod, d_terr, legs, quats, pre_cull=args.pre_cull, timing=True)  # warm-up
times = []
for _ in range(args.reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, ms = lrm.positionability(d_bod, d_terr, legs, quats, pre_cull=args.pre_cull, timing=True)
    torch.cuda.synchronize()
    times.append((time.perf_counter() - t0, ms))
wall, kms = min(times)
got = out.cpu().numpy()
line = {"metric": "body poses/s (4-leg map positionability)", "value": len(bodies) / wall, "unit": "poses/s",
        "poses": len(bodies), "map_points": len(terr), "orientations": len(quats), "legs": args.legs,
        "wall_ms": wall * 1e3, "kernel_ms": kms, "standable": int((got != 0).sum()), "pre_cull": args.pre_cull}
if args.check:
    from oracle.oracle import PortOracle
    port = PortOracle()
    rng = np.random.default_rng(0)
    # verify a pose subsample near the terrain (where the answer is not trivially 0)
    idx = rng.choice(len(bodies), size=min(args.check, len(bodies)), replace=False)
    t0 = time.perf_counter()
    want = port.standability(bodies[idx], terr, [l.as_array() for l in legs], quats, pre_cull=False,
                             threads=os.cpu_count() or 1)
    cpu_s = time.perf_counter() - t0
    if not args.pre_cull:
        line["check"] = {"poses": len(idx), "flag_diff": int(((got[idx] != 0) != (want != 0)).sum()),
                         "standable_oracle": int((want != 0).sum()), "cpu_poses_per_s": len(idx) / cpu_s,
                         "cpu_threads": os.cpu_count()}
print(json.dumps(line))

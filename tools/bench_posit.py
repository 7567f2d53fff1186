"""Positionability benchmarks (BASELINE configs[2..4]).  Prints one JSON line per run.

    python tools/bench_posit.py --config c3            # 4 legs, 1 Mi-point Perlin map, 256^3 poses, 45 RPY
    python tools/bench_posit.py --config c4            # 50 M-point map: lrm_oct + pose search on it
    python tools/bench_posit.py --config c5            # 6 legs at k*pi/3, pose lattice x 16 yaws
    python tools/bench_posit.py [--map 1024] [--poses 64 128 32] [--legs 4] [--yaws 0] [--check 3000]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_posit.py ...

Multi-GPU: rank 0 builds the map and broadcasts it once over NCCL (NVLink /
NVSwitch); the poses are cut into 8 contiguous chunks per rank, dealt round-robin (poses differ by
orders of magnitude in cost) — no collective in the search itself; the standable counts are
gathered at the end and the time is the max over ranks.  The body-space octree (c4) is sharded by
top-level children (lrm_oct_sharded): rank r refines children c with c % N == r.
The CPU check runs the oracle (all host threads) on a random sample of poses of rank 0's slab.
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import lrm_loader
from tests import terrain

PRESETS = {
    # BASELINE configs[2]: 4-leg positionability, 1 Mi-point Perlin map, 256^3 body poses
    "c3": dict(map=[1024, 1024], poses=[256, 256, 256], legs=4, yaws=0, oct_depth=-1),
    # configs[3]: 50 M-point terrain (7168 x 7040 lattice), map replicated per GPU; the body-space
    # octree (apply_oct semantics) over all footholds + the pose search on the same map
    # (the CPU oracle manages ~5 poses/s on this map: check fewer poses)
    "c4": dict(map=[7168, 7040], poses=[128, 128, 64], legs=4, yaws=0, oct_depth=6, check=200),
    # configs[4]: hexapod, mounts k*pi/3, dense pose lattice x yaw grid
    "c5": dict(map=[1024, 1024], poses=[256, 256, 64], legs=6, yaws=16, oct_depth=-1),
}

ap = argparse.ArgumentParser()
ap.add_argument("--config", choices=sorted(PRESETS), default=None)
ap.add_argument("--map", type=int, nargs="+", default=[1024], help="points per side, or ny nx")
ap.add_argument("--poses", type=int, nargs=3, default=[64, 128, 32])
ap.add_argument("--check", type=int, default=2000, help="poses to verify against the CPU oracle")
ap.add_argument("--pre-cull", action="store_true")
ap.add_argument("--legs", type=int, default=4)
ap.add_argument("--yaws", type=int, default=0, help="0: the 45 RPY orientations of robot_full_struct; N: N level yaws")
ap.add_argument("--oct-depth", type=int, default=-1, help=">= 0: also time lrm_oct on the map at this depth")
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
if args.config:
    for k, v in PRESETS[args.config].items():
        setattr(args, k, v)
map_shape = (args.map[0], args.map[0]) if len(args.map) == 1 else (args.map[0], args.map[1])

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

lrm = lrm_loader.load()
from importlib import import_module
slabs = import_module("lrm_b200.slabs")

fixtures = import_module("lrm_b200.fixtures")
n_map = map_shape[0] * map_shape[1]
gen_s = 0.0
if rank == 0:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_terr = fixtures.perlin_terrain(map_shape, device=dev)   # generated on the device (60 s of numpy at 50 M points)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
else:
    d_terr = torch.empty((n_map, 3), dtype=torch.float32, device=dev)
bcast_ms = 0.0
if world > 1:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dist.broadcast(d_terr, src=0)          # the map is replicated once
    torch.cuda.synchronize()
    bcast_ms = (time.perf_counter() - t0) * 1e3
d_all = fixtures.body_lattice(d_terr, *args.poses)
n_all = int(d_all.shape[0])
mine = slabs.dealt_chunks(n_all, rank, world, chunks_per_rank=8)
d_bod = torch.cat([d_all[f:f + c] for f, c in mine]).contiguous()
del d_all
bodies = d_bod.cpu().numpy()
terr = d_terr.cpu().numpy() if rank == 0 else None
legs = [lrm.get_M2_leg(float(np.float32(k) * np.float32(2 * np.pi) / np.float32(args.legs))) for k in range(args.legs)]
quats = lrm.yaw_orientations(args.yaws) if args.yaws > 0 else lrm.full_struct_orientations()
out, ms = lrm.positionability(d_bod, d_terr, legs, quats, pre_cull=args.pre_cull, timing=True)  # warm-up
times = []
for _ in range(args.reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, ms = lrm.positionability(d_bod, d_terr, legs, quats, pre_cull=args.pre_cull, timing=True)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0, ms], dtype=torch.float64, device=dev)
    g = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
    if world > 1:
        dist.all_gather(g, t)
    else:
        g = [t]
    times.append((max(float(x[0]) for x in g), max(float(x[1]) for x in g), min(float(x[1]) for x in g)))
wall, kms, kms_min = min(times)
got = out.cpu().numpy()
standable = torch.tensor([int((got != 0).sum())], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(standable)
if rank == 0:
    line = {"metric": f"body poses/s ({args.legs}-leg map positionability)", "value": n_all / wall,
            "unit": "poses/s", "config": args.config or "custom",
            "n_gpus": world, "poses": n_all, "map_points": n_map, "orientations": len(quats),
            "legs": args.legs, "wall_ms": wall * 1e3, "kernel_ms_max": kms, "kernel_ms_min": kms_min,
            "partition": "8 contiguous chunks per rank, dealt round-robin", "standable": int(standable.item()),
            "pre_cull": args.pre_cull, "map_broadcast_ms": bcast_ms, "map_generation_s": round(gen_s, 3)}
    if args.check and not args.pre_cull:
        from oracle.oracle import PortOracle
        from tests import parity
        port = PortOracle()
        rng = np.random.default_rng(0)
        idx = rng.choice(len(bodies), size=min(args.check, len(bodies)), replace=False)
        la = [l.as_array() for l in legs]
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        want = port.standability(bodies[idx], terr, la, quats, pre_cull=False, threads=threads)
        cpu_s = time.perf_counter() - t0
        rep = parity.pose_report(bodies[idx], got[idx], want,
                                 lambda p: port.standability(p, terr, la, quats, pre_cull=False, threads=threads))
        bad = np.flatnonzero(got[idx] != want)
        line["check"] = {"poses": len(idx), "parity": rep, "standable_oracle": int((want != 0).sum()),
                         "cpu_poses_per_s": len(idx) / cpu_s, "cpu_threads": threads,
                         "mismatches": [{"pose": [float(v) for v in bodies[idx[k]]], "b200": int(got[idx[k]]),
                                         "oracle": int(want[k])} for k in bad[:8]]}
    print(json.dumps(line), flush=True)
if args.oct_depth >= 0:
    # apply_oct semantics on the whole map as footholds, sharded by top-level children: shipped leg
    # (no box can be valid) and a wide-coxa leg (valid boxes appear once boxes are small)
    for name, leg in (("M2_as_shipped", lrm.get_M2_leg(0.0)), ("wide_coxa", None)):
        if leg is None:
            a = lrm.get_M2_leg(0.0).as_array()
            a[8], a[9] = 3.0, -3.0
            leg = lrm.LegDimensions.from_array(a)
        lrm.apply_oct(d_terr, leg, max_depth=args.oct_depth, cap=1 << 20, shard=rank, nshards=world)   # warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, counts, oms = lrm.apply_oct(d_terr, leg, max_depth=args.oct_depth, cap=1 << 20, timing=True,
                                         shard=rank, nshards=world, child_counts=True)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0, oms, float(len(res))], dtype=torch.float64, device=dev)
        g = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(g, t)
        else:
            g = [t]
        if rank == 0:
            owall = max(float(x[0]) for x in g)
            print(json.dumps({"metric": "apply_oct (body-space octree) wall ms", "config": args.config or "custom",
                              "leg": name, "footholds": n_map, "max_depth": args.oct_depth, "n_gpus": world,
                              "sharding": "top-level children c % N == rank (lrm_oct_sharded)",
                              "valid_boxes": int(sum(float(x[2]) for x in g)), "wall_ms": owall * 1e3,
                              "kernel_ms_per_rank": [round(float(x[1]), 1) for x in g],
                              "footholds_per_s": n_map / owall}), flush=True)
if world > 1:
    dist.destroy_process_group()

"""Positionability benchmark (BASELINE configs[2] shape): Perlin terrain map, pose lattice, M2
legs, the 45 orientations of robot_full_struct.  Prints one JSON line (body poses/s).

    python tools/bench_posit.py [--map 1024] [--poses 64 128 32] [--check 3000] [--pre-cull]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_posit.py ...

Multi-GPU: rank 0 builds the map and broadcasts it once over NCCL (NVLink /
NVSwitch); every rank then owns a contiguous slab of poses — no collective in the search itself;
the per-slab standable counts are gathered at the end and the time is the max over ranks.
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
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

n_map = args.map * args.map
d_terr = torch.empty((n_map, 3), dtype=torch.float32, device=dev)
if rank == 0:
    terr = terrain.perlin_terrain(args.map)
    d_terr.copy_(torch.from_numpy(terr))
bcast_ms = 0.0
if world > 1:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dist.broadcast(d_terr, src=0)          # the map is replicated once
    torch.cuda.synchronize()
    bcast_ms = (time.perf_counter() - t0) * 1e3
terr = d_terr.cpu().numpy()
bodies_all = terrain.body_lattice(terr, *args.poses)
first, count = slabs.slab_range(len(bodies_all), rank, world)
bodies = bodies_all[first:first + count]
legs = [lrm.get_M2_leg(float(np.float32(k) * np.float32(2 * np.pi) / np.float32(args.legs))) for k in range(args.legs)]
quats = lrm.full_struct_orientations()
d_bod = torch.from_numpy(bodies).to(dev)
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
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append((float(t[0]), float(t[1])))
wall, kms = min(times)
got = out.cpu().numpy()
standable = torch.tensor([int((got != 0).sum())], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(standable)
if rank == 0:
    line = {"metric": "body poses/s (4-leg map positionability)", "value": len(bodies_all) / wall, "unit": "poses/s",
            "n_gpus": world, "poses": len(bodies_all), "map_points": n_map, "orientations": len(quats),
            "legs": args.legs, "wall_ms": wall * 1e3, "kernel_ms_max": kms, "standable": int(standable.item()),
            "pre_cull": args.pre_cull, "map_broadcast_ms": bcast_ms}
    if args.check and not args.pre_cull:
        from oracle.oracle import PortOracle
        port = PortOracle()
        rng = np.random.default_rng(0)
        idx = rng.choice(len(bodies), size=min(args.check, len(bodies)), replace=False)
        t0 = time.perf_counter()
        want = port.standability(bodies[idx], terr, [l.as_array() for l in legs], quats, pre_cull=False,
                                 threads=os.cpu_count() or 1)
        cpu_s = time.perf_counter() - t0
        line["check"] = {"poses": len(idx), "flag_diff": int(((got[idx] != 0) != (want != 0)).sum()),
                         "standable_oracle": int((want != 0).sum()), "cpu_poses_per_s": len(idx) / cpu_s,
                         "cpu_threads": os.cpu_count()}
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()

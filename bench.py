#!/usr/bin/env python
"""bench.py — headline benchmark (BASELINE.json: "Gpoints/s reach+dist (1-8 B200); body poses/s for
4-leg map positionability").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--scaling weak|strong]
                    [--impl b200|reference] [--no-cpu] [--no-posit]

The headline value: a step = one fused reachability + distance pass (lrm_reach_dist through the C
ABI) over P points of the synthetic x[-100,600] y[-400,400] z[-500,200] mm lattice (BASELINE
configs[1]: 1000^3 = 1e9 points), resident in HBM, M2 leg, identity orientation.
  --scaling weak   (default) every rank sweeps its own P-point lattice over the full extents: rank
                   r's x-planes are shifted by r / N of a pitch, i.e. the ranks interleave the planes
                   of an N x finer lattice — equal work per rank, no collective on the data path.
  --scaling strong ONE P-point lattice (configs[1] as worded) cut into 8 N contiguous slabs of x-planes
                   that are dealt to the N ranks round-robin (slabs differ in content — the first
                   holds no reachable point, the middle ones the whole workspace — so one slab per
                   rank leaves the ranks 0.87 - 1.19 ms apart); a rank sweeps its slabs in one launch.
Rank 0 prints ONE JSON line.  Besides the base contract it carries
  roofline        the fused kernel against the measured HBM peak (25 B / point)
  e2e             the same call with pinned HOST buffers, copies inside the timed region
  cpu_baseline    the reference's CPU path (compiled reference) on the host cores, bounded sample
  parity          the GPU's results on that very sample against the reference's: flags (mismatch /
                  within 1e-3 mm of the boundary / unexplained) and vectors (over 1e-2 mm / on a seam
                  / unexplained); any unexplained point makes the run exit non-zero
  per_rank        kernel ms of every rank (min / max are what the scaling curve is made of)
  setup           what the first call of a new leg costs (certified tables), outside the timed region
  positionability BASELINE configs[2] at N GPUs: body poses/s, kernel ms min / max over ranks,
                  executed and algorithmic leg-predicate counts, oracle check, CPU poses/s

--impl reference times the reference's own CPU implementation of the one-leg path (the compiled
reference oracle/_ref when it was built here, else the pinned C restatement) on all host threads,
on a bounded sample of the same lattice.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LO, HI = (-100.0, -400.0, -500.0), (600.0, 400.0, 200.0)
BYTES_PER_POINT = 25  # 12 B point in + 12 B vector out + 1 B flag out
METRIC = "Gpoints/s reach+dist (1-8 B200)"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def recorded_traffic():
    """Per-launch DRAM bytes of the fused kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa(local):
    """Bind this process to the CPUs that sit next to GPU `local` (sysfs local_cpulist of its PCI
    function), so that the pinned staging buffers allocated afterwards are first touched on that
    NUMA node.  Best effort: returns what it found / did."""
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        info["pci"] = bdf
        info["numa_node"] = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        info["local_cpulist"] = cpulist
        cpus = set()
        for part in cpulist.split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
        info["cpus_used"] = len(os.sched_getaffinity(0))
    except Exception as e:  # no sysfs entry (containers), single node, ...
        info["note"] = f"{type(e).__name__}: {e}"
    return info


def lattice_dims(points):
    """nx x 1000 x 1000 points (z fastest) with nx = ceil(points / 1e6)."""
    ny = nz = 1000
    nx = max(1, -(-points // (ny * nz)))
    return nx, ny, nz


def rank_lattice(lrm, points, rank, world, scaling):
    """(lo, step, dims, [(first, count), ...]) of this rank's share, see the module docstring."""
    if scaling == "strong":
        dims = lattice_dims(points)
        lo, step, d = lrm.lattice_spec(LO, HI, dims)
        planes = dims[1] * dims[2]
        from importlib import import_module
        chunks = import_module("lrm_b200.slabs").dealt_chunks(dims[0], rank, world, chunks_per_rank=8)  # whole x-planes
        return lo, step, d, [(f * planes, c * planes) for f, c in chunks]
    dims = lattice_dims(points)
    lo, step, d = lrm.lattice_spec(LO, HI, dims)
    lo = lo.copy()
    lo[0] = np.float32(lo[0] + np.float32(rank) * step[0] / np.float32(world))        # interleaved x-planes
    return lo, step, d, [(0, points)]


def strided_sample(lrm, n_sample, n_total=10 ** 9):
    """Every k-th point of the 1000^3 lattice (k = n_total // n_sample): indices and coordinates."""
    lo, step, dims = lrm.lattice_spec(LO, HI, (1000, 1000, 1000))
    stride = max(1, n_total // n_sample)
    idx = np.arange(n_sample, dtype=np.int64) * stride
    iz, t = idx % 1000, idx // 1000
    iy, ix = t % 1000, t // 1000
    pts = np.stack([lo[0] + ix.astype(np.float32) * step[0], lo[1] + iy.astype(np.float32) * step[1],
                    lo[2] + iz.astype(np.float32) * step[2]], 1).astype(np.float32)
    return idx, pts, stride


def cpu_baseline(points_sample, threads, keep=False):
    """Reference CPU path on a bounded lattice sample; returns a dict (rate in Gpoints/s, kind, n,
    seconds, and — keep=True — the sample with the reference's own results for the parity block)."""
    from oracle.oracle import best
    import lrm_loader
    lrm = lrm_loader.load()
    oracle = best()
    leg = oracle.get_leg(1, 0.0)
    idx, pts, stride = strided_sample(lrm, points_sample)
    t0 = time.perf_counter()
    reach = oracle.reach(pts, leg, threads=threads)
    vec, dflag = oracle.dist(pts, leg, threads=threads)
    dt = time.perf_counter() - t0
    out = {"rate": points_sample / dt / 1e9, "kind": oracle.kind, "n": points_sample, "seconds": dt, "stride": stride}
    if keep:
        out.update(idx=idx, pts=pts, reach=reach, vec=vec, oracle=oracle, leg=leg)
    return out


def parity_block(sample, gpu_flags, gpu_vec, threads):
    """BASELINE.json's bar on the cpu_baseline sample: flags bit-exact except within 1e-3 mm of the
    boundary (counted), vectors within 1e-2 mm (seam points classified, see tests/parity.py)."""
    from tests import parity
    oracle, leg, pts = sample["oracle"], sample["leg"], sample["pts"]
    fr = parity.flag_report(pts, gpu_flags, sample["reach"], lambda p: oracle.reach(p, leg, threads=threads))
    dr = parity.dist_report(pts, gpu_vec, sample["vec"], lambda p: oracle.dist(p, leg, threads=threads)[0])
    ok = fr["unexplained"] == 0 and dr["unexplained"] == 0
    return {"sample": f"{len(pts)} lattice points (the cpu_baseline sample), oracle kind {oracle.kind}",
            "flags": fr, "vectors": dr, "flag_band_mm": parity.FLAG_BAND_MM, "vector_tol_mm": parity.DIST_TOL_MM,
            "green": ok}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # calibrate on a small sample, then size each step for ~4 s of wall time
    rate = cpu_baseline(200_000, threads)["rate"]
    per_step = int(min(max(rate * 1e9 * 4.0, 200_000), 20_000_000))
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline(min(per_step, 400_000), threads)
    times, kind = [], "port"
    for _ in range(args.steps):
        r = cpu_baseline(per_step, threads)
        kind = r["kind"]
        times.append(r["seconds"])
    total = sum(times)
    value = per_step * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "one-leg reach+dist, 1e9-point lattice per GPU (BASELINE configs[1]), M2 leg",
                   "note": "reference CPU path (reachability_global + distance_global, one_leg_global.cu:74-147) "
                           "on all host threads over a strided sample of the lattice"},
        "cpu_baseline": {"value": value, "unit": "Gpoints/s", "cores": threads, "kind": kind,
                         "sample": f"{per_step} lattice points per step (every {10**9 // per_step}-th of 1e9)"},
        "e2e": {"value": value, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- positionability (BASELINE configs[2]) -----------------------------------------------------
def run_positionability(lrm, torch, dist, dev, rank, world, args):
    """4 M2 legs at k*pi/2, 1 Mi-point Perlin map, 256^3 body poses, the 45 orientations of
    robot_full_struct.  Rank 0 builds the map, ONE broadcast replicates it; the poses are cut into
    8 chunks per rank dealt round-robin (poses differ by orders of magnitude in cost); no collective
    in the search; standable counts and kernel times are reduced at the end."""
    from importlib import import_module
    slabs, fixtures = import_module("lrm_b200.slabs"), import_module("lrm_b200.fixtures")
    side, poses = args.posit_map, args.posit_poses
    n_map = side * side
    gen_s = 0.0
    if rank == 0:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d_terr = fixtures.perlin_terrain((side, side), device=dev)     # generated on the device
        torch.cuda.synchronize()
        gen_s = time.perf_counter() - t0
    else:
        d_terr = torch.empty((n_map, 3), dtype=torch.float32, device=dev)
    bcast_ms = 0.0
    if world > 1:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.broadcast(d_terr, src=0)
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3
    d_all = fixtures.body_lattice(d_terr, poses, poses, poses)
    mine = slabs.dealt_chunks(d_all.shape[0], rank, world, chunks_per_rank=8)
    d_bod = torch.cat([d_all[f:f + c] for f, c in mine]).contiguous() if mine else d_all[:0]
    n_poses_all = int(d_all.shape[0])
    del d_all
    bodies = d_bod.cpu().numpy()
    terr = d_terr.cpu().numpy() if rank == 0 else None
    legs = [lrm.get_M2_leg(float(np.float32(k) * np.float32(np.pi / 2))) for k in range(4)]
    quats = lrm.full_struct_orientations()
    out, ms = lrm.positionability(d_bod, d_terr, legs, quats, timing=True)   # warm-up
    best = None
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, ms = lrm.positionability(d_bod, d_terr, legs, quats, timing=True)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t = torch.tensor([wall, ms], dtype=torch.float64, device=dev)
        g = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(g, t)
        else:
            g = [t]
        walls, kms = [float(x[0]) for x in g], [float(x[1]) for x in g]
        if best is None or max(walls) < max(best[0]):
            best = (walls, kms)
    walls, kms = best
    got = out.cpu().numpy()
    standable = torch.tensor([int((got != 0).sum())], dtype=torch.int64, device=dev)
    # predicate counts on every 64th pose of this rank (instrumented kernel, outside the timed runs)
    sub = d_bod[::64].contiguous()
    _, cnt = lrm.positionability_counts(sub, d_terr, legs, quats)
    scale = len(bodies) / max(1, sub.shape[0])
    c = torch.tensor([cnt["leg_predicates_executed"] * scale, cnt["cylinder_predicates_executed"] * scale,
                      cnt["leg_predicates_algorithmic"] * scale], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(standable)
        dist.all_reduce(c)
    if rank != 0:
        return None
    wall = max(walls)
    rec = {"metric": "body poses/s (4-leg map positionability)", "value": n_poses_all / wall, "unit": "poses/s",
           "config": f"BASELINE configs[2]: 4 M2 legs at k*pi/2, {n_map}-point Perlin map (seed 42), {poses}^3 = "
                     f"{n_poses_all} body poses, 45 orientations of robot_full_struct, no pre-cull",
           "n_gpus": world, "poses": n_poses_all, "map_points": n_map, "orientations": int(len(quats)),
           "wall_ms": wall * 1e3, "kernel_ms_min": min(kms), "kernel_ms_max": max(kms),
           "kernel_ms_per_rank": [round(k, 2) for k in kms], "standable": int(standable.item()),
           "partition": "8 contiguous chunks per rank, dealt round-robin; map broadcast once",
           "map_broadcast_ms": bcast_ms, "map_generation_s": round(gen_s, 3),
           "map_generator": "lrm_b200.fixtures.perlin_terrain on the device (bit-identical to the reference's numpy generator)",
           "predicates": {"leg_executed": float(c[0]), "cylinder_executed": float(c[1]),
                          "leg_algorithmic": float(c[2]),
                          "leg_executed_per_s": float(c[0]) / (max(kms) * 1e-3),
                          "leg_algorithmic_per_s": float(c[2]) / (max(kms) * 1e-3),
                          "executed_over_algorithmic": float(c[0]) / max(1.0, float(c[2])),
                          "note": "counted by the instrumented kernel on every 64th pose, scaled by 64; algorithmic "
                                  "= for every orientation, every map point inside the reach cylinder, once per leg "
                                  "(no pruning, no early exit: what reach_mem_kernel evaluates, several_leg.cu:92-129)"}}
    if not args.no_cpu:
        from oracle.oracle import PortOracle
        from tests import parity
        port = PortOracle()
        rng = np.random.default_rng(0)
        idx = rng.choice(len(bodies), size=min(args.posit_check, len(bodies)), replace=False)
        la = [l.as_array() for l in legs]
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        want = port.standability(bodies[idx], terr, la, quats, pre_cull=False, threads=threads)
        cpu_s = time.perf_counter() - t0
        rep = parity.pose_report(bodies[idx], got[idx], want,
                                 lambda p: port.standability(p, terr, la, quats, pre_cull=False, threads=threads))
        rec["parity"] = dict(rep, sample=f"{len(idx)} random poses of rank 0's share vs op_standability "
                                         "(pinned to the reference GPU pipeline, tests/test_refgpu_pin.py)",
                             green=rep["unexplained"] == 0 and rep["flag_mismatch"] == 0)
        rec["cpu_baseline"] = {"value": len(idx) / cpu_s, "unit": "poses/s", "cores": threads, "kind": "port",
                               "sample": f"{len(idx)} poses, {cpu_s:.1f} s"}
    return rec


def run_b200(args):
    import torch
    import torch.distributed as dist
    import lrm_loader
    lrm = lrm_loader.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path); use --impl reference")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version
    # banner at the first collective, ...) goes to stderr — fd 1 is pointed at fd 2 for the run and
    # the line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    leg = lrm.get_M2_leg(0.0)
    lo, step, dims, chunks = rank_lattice(lrm, args.points, rank, world, args.scaling)
    n = sum(c for _, c in chunks)
    first = chunks[0][0] if len(chunks) == 1 else -1
    pts = torch.empty((n, 3), dtype=torch.float32, device=dev)
    off = 0
    for f, c in chunks:                       # this rank's slabs, back to back in one buffer
        lrm.make_lattice(pts[off:off + c], lo, step, dims, first=f, count=c)
        off += c
    flags = torch.empty(n, dtype=torch.uint8, device=dev)
    vec = torch.empty((n, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()

    def step_fn():
        lrm.reach_dist(pts, leg, None, out_flags=flags, out_vec=vec, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # setup cost of a new (leg, orientation): the first call builds the plane atlas on the caller's
    # stream and starts the choice volume on a side stream; the tiered sweep takes over once it is there
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step_fn()
    torch.cuda.synchronize()
    first_call_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    builds0 = lrm.get_stat("volume_builds")
    # until a sweep has found the background build of the choice volume finished (the two-tier sweep
    # answers meanwhile, same bits), and a little beyond; at most 5 s
    while True:
        step_fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dt > 5.0 or (dt > 0.4 and lrm.get_stat("volume_builds") > builds0):
            break
    volume_wait_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(max(args.warmup, 3)):
        step_fn()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local) as clocks:
        # nvidia-smi needs a few hundred ms per sample: keep the GPU under the same load a little
        # longer than the timed steps so that the clock record has several samples under load
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.6:
            step_fn()
            torch.cuda.synchronize()
        barrier()
        ev[0].record(stream)
        for k in range(args.steps):
            step_fn()
            ev[k + 1].record(stream)
        barrier()
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    clk = clocks.summary()
    mine = torch.tensor([total_ms, float(np.mean(per_launch_ms)), float(n), float(clk["sm_mhz"] or 0.0)],
                        dtype=torch.float64, device=dev)
    gathered = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, mine)
    else:
        gathered = [mine]
    total_ms_max = max(float(g[0]) for g in gathered)
    n_all = int(sum(float(g[2]) for g in gathered))
    value = n_all * args.steps / (total_ms_max * 1e-3) / 1e9
    reach_count = int(flags.sum().item())

    # end to end through the C ABI with HOST buffers (pinned), copies inside the timed region;
    # the buffers are first touched on the GPU's own NUMA node
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(local)
    ne = min(args.e2e_points, n)
    h_pts = torch.empty((ne, 3), dtype=torch.float32).pin_memory()
    h_pts.copy_(pts[:ne])
    h_flags = torch.empty(ne, dtype=torch.uint8).pin_memory()
    h_vec = torch.empty((ne, 3), dtype=torch.float32).pin_memory()
    hp, hf, hv = h_pts.numpy(), h_flags.numpy(), h_vec.numpy()
    for _ in range(2):
        lrm.reach_dist(hp, leg, None, out_flags=hf, out_vec=hv)
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lrm.reach_dist(hp, leg, None, out_flags=hf, out_vec=hv)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = ne * world * e2e_steps / float(t.item()) / 1e9
    assert np.array_equal(hf[:4096], flags[:4096].cpu().numpy())
    os.sched_setaffinity(0, all_cpus)     # the cpu_baseline leg uses every host thread

    line = None
    exit_code = 0
    if rank == 0:
        peak, peak_kind = measured_peaks()
        kernel_ms = float(gathered[0][1])
        achieved = BYTES_PER_POINT * n / (kernel_ms * 1e-3) / 1e9
        traffic = recorded_traffic()
        cell_mm, vdim = lrm.get_stat("volume_cell_mm"), int(lrm.get_stat("volume_dim"))
        line = {
            "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"one-leg reach+dist fused sweep, {n} lattice points per GPU "
                                   f"(BASELINE configs[1]: 1e9-point grid), M2 leg, identity orientation",
                       "points_per_gpu": n, "points_total": n_all,
                       "lattice": "x[-100,600] y[-400,400] z[-500,200] mm, 0.7/0.8/0.7 mm pitch; "
                                  + ("weak: every rank sweeps the full extents, x-planes shifted by rank/N of a pitch"
                                     if args.scaling == "weak" else
                                     "strong: one lattice, 8 contiguous slabs of x-planes per rank dealt round-robin"),
                       "l2": f"inputs+outputs {BYTES_PER_POINT * n / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed",
                       "reachable_points_rank0": reach_count},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_kind,
                         # ncu dram__bytes_read+write of this kernel (profiles/traffic.json, captured at
                         # 1e9 points per launch; scaled per point for other sizes)
                         "traffic": (traffic["dram_bytes_per_point"] * n) if traffic else None,
                         "kernel": "one_leg_tier_kernel<both,aos>", "kernel_ms": kernel_ms,
                         "bytes_per_point": BYTES_PER_POINT},
            "e2e": {"value": e2e_value, "unit": "Gpoints/s", "h2d_bytes_per_step": 12 * ne,
                    "d2h_bytes_per_step": 13 * ne, "points_per_step": ne,
                    "note": "lrm_reach_dist with pinned host buffers, H2D + kernel + D2H per step", "numa_rank0": numa},
            # per step: the coherence probe (1 CTA), the tiered sweep it selects for a lattice, and the
            # two-tier sweep that reads the verdict and returns at once
            "gpu_launches": 3 * args.steps, "clocks": clk,
            "per_rank": {"kernel_ms": [round(float(g[1]), 4) for g in gathered],
                         "kernel_ms_min": min(float(g[1]) for g in gathered),
                         "kernel_ms_max": max(float(g[1]) for g in gathered),
                         "sm_mhz": [float(g[3]) for g in gathered]},
            "setup": {"first_call_ms": first_call_ms,
                      "note": "first lrm_reach_dist of a new (leg, orientation), synchronised: plane atlas (16 MiB) "
                              "built on the caller's stream + that call's two-tier sweep; the choice volume "
                              f"({vdim}^3 cubes of {cell_mm} mm, 32-bit texels) builds on a side stream meanwhile; "
                              "volume_wait_ms = wall time of the sweeps issued until one found it finished (>= 400)",
                      "volume_wait_ms": volume_wait_ms,
                      "table_bytes": 3 * 4096 * 4096 + 4 * vdim ** 3 + 128 * int(lrm.get_stat("volume_bricks")),
                      "bricks": int(lrm.get_stat("volume_bricks")), "table_builds": lrm.get_stat("table_builds")},
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate = cpu_baseline(100_000, threads)["rate"]
            sample = int(min(max(rate * 1e9 * 12.0, 200_000), 40_000_000))
            s = cpu_baseline(sample, threads, keep=True)
            line["cpu_baseline"] = {"value": s["rate"], "unit": "Gpoints/s", "cores": threads, "kind": s["kind"],
                                    "sample": f"{s['n']} lattice points (every {s['stride']}-th of the 1e9 lattice), "
                                              f"{s['seconds']:.1f} s"}
            if n == 10 ** 9 and first == 0:
                # the sample indexes the 1e9 lattice this GPU has just swept: same points, same bits
                sel = torch.from_numpy(s["idx"]).to(dev)
                assert np.array_equal(pts[sel[:1000]].cpu().numpy(), s["pts"][:1000]), "sample / lattice mismatch"
                line["parity"] = parity_block(s, flags[sel].cpu().numpy(), vec[sel].cpu().numpy(), threads)
            else:
                # smaller runs: sweep the sample itself (host pointers)
                gf, gv = lrm.reach_dist(s["pts"], leg)
                line["parity"] = parity_block(s, gf, gv, threads)
            if not line["parity"]["green"]:
                exit_code = 3
    del pts, vec, flags, h_pts, h_vec, h_flags
    torch.cuda.empty_cache()

    if not args.no_posit:
        rec = run_positionability(lrm, torch, dist, dev, rank, world, args)
        if rank == 0:
            line["positionability"] = rec
            if rec.get("parity") and not rec["parity"]["green"]:
                exit_code = 3
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if exit_code:
        sys.exit(exit_code)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=1_000_000_000,
                    help="points per GPU (weak scaling) or in total (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-points", type=int, default=1 << 26)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs")
    ap.add_argument("--no-posit", action="store_true", help="skip the positionability sub-record")
    ap.add_argument("--posit-map", type=int, default=1024, help="map points per side (multiple of 128)")
    ap.add_argument("--posit-poses", type=int, default=256, help="body poses per axis")
    ap.add_argument("--posit-check", type=int, default=2000, help="poses verified against the CPU oracle")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — headline benchmark of the one-leg reach+dist sweep (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl b200|reference]

A step = one fused reachability + distance pass (lrm_reach_dist through the C ABI) over a slab of
P points of the synthetic x[-100,600] y[-400,400] z[-500,200] mm lattice, resident in HBM, M2 leg,
identity orientation.  Each rank owns its own 1000 x 1000 x (P/1e6) slab: weak scaling, no
collective on the data path.  Rank 0 prints ONE JSON line (metric Gpoints/s, whole job).

--impl reference times the reference's own CPU implementation of the same path (the compiled
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


def slab_dims(points):
    """Lattice of nx x 1000 x 1000 points (z fastest) with nx = ceil(points / 1e6)."""
    ny = nz = 1000
    nx = max(1, -(-points // (ny * nz)))
    return nx, ny, nz


def cpu_baseline(points_sample, threads):
    """Reference CPU path on a bounded lattice sample; returns (Gpoints/s, kind, n, seconds)."""
    from oracle.oracle import best
    import lrm_loader
    lrm = lrm_loader.load()
    oracle = best()
    leg = oracle.get_leg(1, 0.0)
    lo, step, dims = lrm.lattice_spec(LO, HI, (1000, 1000, 1000))
    # a strided sample of the 1e9 lattice: every k-th point, same extents
    stride = 10 ** 9 // points_sample
    idx = np.arange(points_sample, dtype=np.int64) * stride
    iz, t = idx % 1000, idx // 1000
    iy, ix = t % 1000, t // 1000
    pts = np.stack([lo[0] + ix.astype(np.float32) * step[0], lo[1] + iy.astype(np.float32) * step[1],
                    lo[2] + iz.astype(np.float32) * step[2]], 1).astype(np.float32)
    t0 = time.perf_counter()
    oracle.reach(pts, leg, threads=threads)
    oracle.dist(pts, leg, threads=threads)
    dt = time.perf_counter() - t0
    return points_sample / dt / 1e9, oracle.kind, points_sample, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # calibrate on a small sample, then size each step for ~4 s of wall time
    rate, kind, _, _ = cpu_baseline(200_000, threads)
    per_step = int(min(max(rate * 1e9 * 4.0, 200_000), 20_000_000))
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline(min(per_step, 400_000), threads)
    times = []
    for _ in range(args.steps):
        _, kind, n, dt = cpu_baseline(per_step, threads)
        times.append(dt)
    total = sum(times)
    value = per_step * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "one-leg reach+dist, 1e9-point lattice per GPU (BASELINE configs[1]), M2 leg",
                   "note": "reference CPU path (reachability_global + distance_global, one_leg_global.cu:74-147) "
                           "on all host threads over a strided sample of the lattice"},
        "cpu_baseline": {"value": value, "unit": "Gpoints/s", "cores": threads, "kind": kind,
                         "sample": f"{per_step} lattice points per step (every {10**9 // per_step}-th of 1e9)"},
        "e2e": {"value": value, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's banner / debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.points
    leg = lrm.get_M2_leg(0.0)
    # one lattice over the fixed extents with world x (points/1e6) x-planes; rank r owns the
    # contiguous slab of x-planes [r, r+1) * points/1e6 (at 1 GPU and 1e9 points: configs[1] itself)
    nx, ny, nz = slab_dims(n * world)
    lo, step, dims = lrm.lattice_spec(LO, HI, (nx, ny, nz))
    dev = torch.device("cuda", local)
    pts = torch.empty((n, 3), dtype=torch.float32, device=dev)
    from importlib import import_module
    first, count = import_module("lrm_b200.slabs").weak_slab(n, rank, world)
    lrm.make_lattice(pts, lo, step, dims, first=first, count=count)   # this rank's contiguous slab
    flags = torch.empty(n, dtype=torch.uint8, device=dev)
    vec = torch.empty((n, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()

    def step_fn():
        lrm.reach_dist(pts, leg, None, out_flags=flags, out_vec=vec, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = n * world * args.steps / (total_ms_max * 1e-3) / 1e9
    reach_count = int(flags.sum().item())

    # end to end through the C ABI with HOST buffers (pinned), copies inside the timed region
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

    if rank == 0:
        peak, peak_kind = measured_peaks()
        kernel_ms = float(np.mean(per_launch_ms))
        achieved = BYTES_PER_POINT * n / (kernel_ms * 1e-3) / 1e9
        traffic = recorded_traffic()
        clk = clocks.summary()
        line = {
            "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"one-leg reach+dist fused sweep, {n} lattice points per GPU "
                                   f"(BASELINE configs[1]: 1e9-point grid), M2 leg, identity orientation",
                       "points_per_gpu": n, "lattice": "x[-100,600] y[-400,400] z[-500,200] mm, 0.7/0.8/0.7 mm pitch",
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
                    "note": "lrm_reach_dist with pinned host buffers, H2D + kernel + D2H per step"},
            # per step: the coherence probe (1 CTA), the tiered sweep it selects for a lattice, and the
            # two-tier sweep that reads the verdict and returns at once
            "gpu_launches": 3 * args.steps, "clocks": clk,
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate, kind, _, _ = cpu_baseline(100_000, threads)
            sample = int(min(max(rate * 1e9 * 12.0, 200_000), 40_000_000))
            v, kind, ns, dt = cpu_baseline(sample, threads)
            line["cpu_baseline"] = {"value": v, "unit": "Gpoints/s", "cores": threads, "kind": kind,
                                    "sample": f"{ns} lattice points (every {10**9 // ns}-th of the 1e9 lattice), "
                                              f"{dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=1_000_000_000, help="points per GPU")
    ap.add_argument("--e2e-points", type=int, default=1 << 26)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

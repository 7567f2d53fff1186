"""Worker for tests/test_multi_gpu_cpu.py: one rank of a world_size-2 gloo job on CPU.

Exercises the host-side logic of the multi-GPU path (slab partition, per-rank lattice slab,
barrier, max-over-ranks timing, gather of per-slab summaries).  The per-slab compute is the CPU
oracle standing in for the CUDA kernel (there is no GPU here and the product has no CPU path)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrm_loader  # noqa: E402
from oracle.oracle import PortOracle  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}",
                            rank=rank, world_size=world)
    lrm = lrm_loader.load()
    from importlib import import_module
    slabs = import_module("lrm_b200.slabs")
    port = PortOracle()
    leg = port.get_leg(1, 0.0)
    dims = (12, 30, 31)
    lo, step, d = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), dims)
    total = dims[0] * dims[1] * dims[2]
    first, count = slabs.slab_range(total, rank, world)
    pts = lrm.lattice_host(lo, step, d, first=first, count=count)
    dist.barrier()
    t0 = time.perf_counter()
    flags = port.reach(pts, leg)
    vec, _ = port.dist(pts, leg)
    elapsed = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)        # the benchmark's max-over-ranks clock
    summary = torch.tensor([first, count, int(flags.sum()), float(np.abs(vec).sum())], dtype=torch.float64)
    gathered = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, summary)
    if rank == 0:
        full = lrm.lattice_host(lo, step, d)
        want_flags = port.reach(full, leg)
        want_vec, _ = port.dist(full, leg)
        out = {"slabs": [g.tolist() for g in gathered], "total": total, "max_elapsed": float(elapsed.item()),
               "want_reach": int(want_flags.sum()), "want_abs": float(np.abs(want_vec).sum())}
        print("RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host<->device copy bandwidth of this box with pinned memory (what bounds bench.py's e2e leg):
H2D alone, D2H alone, both directions at once, at several transfer sizes.

    python tools/pcie_probe.py                                         # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
                                                                       # all GPUs AT ONCE: the ceiling of
                                                                       # the 8-rank end-to-end leg
Every rank binds itself to the CPUs local to its GPU (sysfs local_cpulist) before it allocates its
pinned buffers, so that they are first touched on the GPU's NUMA node; the line reports what it found."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import bind_to_gpu_numa

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
numa = bind_to_gpu_numa(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = {}
for mb in (24, 96, 384):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, reps=10):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return n * reps / float(dt.item()) / 1e9      # per GPU, at the pace of the slowest rank

    run(True, True, 2)
    res[f"{mb}MiB"] = {"h2d_GBs_per_gpu": run(True, False), "d2h_GBs_per_gpu": run(False, True),
                       "both_each_GBs_per_gpu": run(True, True)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "numa_rank0": numa, "sizes": res}))
if world > 1:
    dist.destroy_process_group()

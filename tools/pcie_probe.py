"""Host<->device copy bandwidth of this box with pinned memory (what bounds bench.py's e2e leg):
H2D alone, D2H alone, both directions at once, at several transfer sizes."""
import json, sys, time
import torch
res = {}
for mb in (8, 24, 96, 384):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(up, down, reps=10):
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
        return n * reps / (time.perf_counter() - t0) / 1e9
    run(True, True, 2)
    res[f"{mb}MiB"] = {"h2d_GBs": run(True, False), "d2h_GBs": run(False, True), "both_each_GBs": run(True, True)}
print(json.dumps(res))

"""Kernel-only throughput of the three one-leg modes (reach, dist, fused) on a resident lattice slab,
with the per-mode algorithmic bytes (13 / 24 / 25 B per point) against the measured HBM peak."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
lrm = lrm_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
leg = lrm.get_M2_leg(0.0)
lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (max(1, n // 1_000_000), 1000, 1000))
pts = torch.empty((n, 3), dtype=torch.float32, device="cuda")
lrm.make_lattice(pts, lo, step, dims, 0, n)
flags = torch.empty(n, dtype=torch.uint8, device="cuda")
vec = torch.empty((n, 3), dtype=torch.float32, device="cuda")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
out = {}
for name, fn, bpp in (("reach", lambda: lrm.reachability(pts, leg, out=flags), 13),
                      ("dist", lambda: lrm.distance(pts, leg, out=vec, flags=False), 24),
                      ("reach_dist", lambda: lrm.reach_dist(pts, leg, out_flags=flags, out_vec=vec), 25)):
    import time
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 0.3:   # the choice volume of a new leg builds in the background
        fn()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[name] = {"ms": ms, "gpoints_s": n / ms / 1e6, "gb_s": bpp * n / ms / 1e6, "frac_of_hbm_peak": bpp * n / ms / 1e6 / peak}
print(json.dumps({"points": n, "modes": out}))

"""A few launches of one one-leg kernel on a resident lattice slab (profiling target).
    python tools/run_fused.py [points] [reps] [reach|dist|both] [sweep: 0 two-tier | 1 tiered | 2 auto]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
lrm = lrm_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
leg = lrm.get_M2_leg(0.0)
lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (max(1, n // 1_000_000), 1000, 1000))
pts = torch.empty((n, 3), dtype=torch.float32, device="cuda")
lrm.make_lattice(pts, lo, step, dims, 0, n)
flags = torch.empty(n, dtype=torch.uint8, device="cuda")
vec = torch.empty((n, 3), dtype=torch.float32, device="cuda")
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
if len(sys.argv) > 4:
    lrm.set_option("sweep", int(sys.argv[4]))
if os.environ.get("LRM_TC_BRICKS"):
    lrm.set_option("volume_bricks", int(os.environ["LRM_TC_BRICKS"]))
for _ in range(reps):
    if mode == "reach":
        lrm.reachability(pts, leg, out=flags)
    elif mode == "dist":
        lrm.distance(pts, leg, out=vec, flags=False)
    else:
        lrm.reach_dist(pts, leg, out_flags=flags, out_vec=vec)
torch.cuda.synchronize()
print("reachable", int(flags.sum().item()))

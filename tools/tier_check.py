"""A/B of the tiered distance sweep (choice volume) against the two-tier one (lrm_set_option "sweep" 0)
on the same resident lattice slab: results must be identical up to float rounding of the
projection (flags exactly), and both are timed.  Also reports what the first call (volume build)
costs.  Usage: python tools/tier_check.py [points] [random] [volume_cell_mm volume_dim]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
lrm = lrm_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
leg = lrm.get_M2_leg(0.0)
pts = torch.empty((n, 3), dtype=torch.float32, device="cuda")
if len(sys.argv) > 2 and sys.argv[2] == "random":
    g = torch.Generator(device="cuda").manual_seed(3)
    pts.uniform_(-700, 700, generator=g)
else:
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (max(1, n // 1_000_000), 1000, 1000))
    lrm.make_lattice(pts, lo, step, dims, 0, n)
out = {"points": n}
if len(sys.argv) > 4:
    lrm.set_option("volume_cell_mm", float(sys.argv[3]))
    lrm.set_option("volume_dim", int(sys.argv[4]))
    out["volume"] = [float(sys.argv[3]), int(sys.argv[4])]
if len(sys.argv) > 5:
    lrm.set_option("tier_kernel", int(sys.argv[5]))
    out["tier_kernel"] = int(sys.argv[5])
if os.environ.get("LRM_TC_KSHIFT"):
    lrm.set_option("tier_chunk_shift", int(os.environ["LRM_TC_KSHIFT"]))
    out["tier_chunk_shift"] = int(os.environ["LRM_TC_KSHIFT"])
if os.environ.get("LRM_TC_BRICKS"):
    lrm.set_option("volume_bricks", int(os.environ["LRM_TC_BRICKS"]))
    out["volume_bricks_option"] = int(os.environ["LRM_TC_BRICKS"])
res = {}
for name, mode in (("two_tier", 0), ("three_tier", 1), ("auto", 2)):
    lrm.set_option("sweep", mode)
    flags = torch.empty(n, dtype=torch.uint8, device="cuda")
    vec = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    lrm.reach_dist(pts, leg, out_flags=flags, out_vec=vec)
    torch.cuda.synchronize()
    first = time.time() - t0
    for _ in range(2):
        lrm.reach_dist(pts, leg, out_flags=flags, out_vec=vec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lrm.reach_dist(pts, leg, out_flags=flags, out_vec=vec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    dvec = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    e0.record()
    for _ in range(5):
        lrm.distance(pts, leg, out=dvec, flags=False)
    e1.record()
    torch.cuda.synchronize()
    dms = e0.elapsed_time(e1) / 5
    rflags = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        lrm.reachability(pts, leg, out=rflags)
    e0.record()
    for _ in range(5):
        lrm.reachability(pts, leg, out=rflags)
    e1.record()
    torch.cuda.synchronize()
    rms = e0.elapsed_time(e1) / 5
    out[name] = {"first_call_s": first, "fused_ms": ms, "fused_gpoints_s": n / ms / 1e6, "dist_ms": dms,
                 "dist_gpoints_s": n / dms / 1e6, "reach_ms": rms, "reach_gpoints_s": n / rms / 1e6}
    res[name] = (flags, vec, dvec, rflags)
    del dvec
fa, va, da, ra = res["two_tier"]
fb, vb, db, rb = res["three_tier"]
diff = (va - vb).abs().amax(dim=1)
out["flags_equal"] = bool(torch.equal(fa, fb))
out["max_vec_diff_mm"] = float(diff.max())
out["points_over_1e-3"] = int((diff > 1e-3).sum())
out["dist_mode_max_diff_mm"] = float((da - db).abs().max())
out["reachable"] = int(fb.sum())
fc, vc, dc, rc = res["auto"]
out["auto_equal"] = bool(torch.equal(fc, fa) and torch.equal(vc, va) and torch.equal(dc, da))
out["reach_equal"] = bool(torch.equal(ra, rb) and torch.equal(ra, rc))
out["reach_vs_fused_mismatch"] = int((ra != fa).sum())
out["bricks_used"], out["brick_capacity"] = lrm.get_stat("volume_bricks"), lrm.get_stat("volume_brick_capacity")
print(json.dumps(out))

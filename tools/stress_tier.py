"""Stress the tiered sweep: repeated launches at several sizes, each compared bit for bit with the
two-tier sweep of the same points (a race in the ring hand-overs would show as a differing byte or a
fault).  python tools/stress_tier.py [reps] [sizes...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
lrm = lrm_loader.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sizes = [int(a) for a in sys.argv[2:]] or [4_200_000, 33_000_000, 250_000_000]
leg = lrm.get_M2_leg(0.0)
out = {}
for n in sizes:
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (max(1, -(-n // 1_000_000)), 1000, 1000))
    pts = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    lrm.make_lattice(pts, lo, step, dims, 0, n)
    lrm.set_option("sweep", 0)
    f0, v0 = lrm.reach_dist(pts, leg)
    torch.cuda.synchronize()
    bad = 0
    for kern in (0, 1):
        lrm.set_option("sweep", 1)
        lrm.set_option("tier_kernel", kern)
        for r in range(reps):
            f1, v1 = lrm.reach_dist(pts, leg)
            torch.cuda.synchronize()
            if not (torch.equal(f0, f1) and torch.equal(v0, v1)):
                bad += 1
    out[str(n)] = {"reps": 2 * reps, "differing_runs": bad}
    print(n, out[str(n)], flush=True)
    del pts, f0, v0, f1, v1
print(json.dumps(out))

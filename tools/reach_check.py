"""Steady-state reach-only / distance throughput per robot at a given size (tables warm), next to
the first calls.  python tools/reach_check.py [points]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
lrm = lrm_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (max(1, -(-n // 1_000_000)), 1000, 1000))
pts = torch.empty((n, 3), dtype=torch.float32, device="cuda")
lrm.make_lattice(pts, lo, step, dims, 0, n)
fl = torch.empty(n, dtype=torch.uint8, device="cuda")
vec = torch.empty((n, 3), dtype=torch.float32, device="cuda")
out = {"points": n}
for robot, name in ((1, "M2"), (0, "moonbot")):
    leg = lrm.get_leg(robot, 0.0)
    first = [lrm.reachability(pts, leg, out=fl, timing=True)[1] for _ in range(4)]
    time.sleep(0.3)                      # the choice volume builds in the background
    warm = [lrm.reachability(pts, leg, out=fl, timing=True)[1] for _ in range(6)]
    dwarm = [lrm.distance(pts, leg, out=vec, flags=False, timing=True)[-1] for _ in range(6)]
    out[name] = {"reach_first_calls_ms": [round(x, 3) for x in first], "reach_warm_ms": round(min(warm), 4),
                 "reach_warm_gpts": n / min(warm) / 1e6, "dist_warm_ms": round(min(dwarm), 4),
                 "dist_warm_gpts": n / min(dwarm) / 1e6, "reachable": int(fl.sum().item())}
print(json.dumps(out))

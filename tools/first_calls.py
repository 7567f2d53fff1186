"""What the first calls of a new (leg, orientation) cost: per-call wall and kernel ms of 40
back-to-back large sweeps of a fresh plan (the choice volume builds in the background meanwhile).
    python tools/first_calls.py [points]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
lrm = lrm_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (max(1, -(-n // 1_000_000)), 1000, 1000))
pts = torch.empty((n, 3), dtype=torch.float32, device="cuda")
lrm.make_lattice(pts, lo, step, dims, 0, n)
fl = torch.empty(n, dtype=torch.uint8, device="cuda")
vec = torch.empty((n, 3), dtype=torch.float32, device="cuda")
lrm.reach_dist(pts, lrm.get_M2_leg(0.0), out_flags=fl, out_vec=vec)      # loads the modules
torch.cuda.synchronize()
time.sleep(0.5)
leg = lrm.get_moonbot_leg(0.3)                                             # a fresh plan
walls, kms = [], []
t_start = time.perf_counter()
for k in range(40):
    t0 = time.perf_counter()
    _, _, ms = lrm.reach_dist(pts, leg, out_flags=fl, out_vec=vec, timing=True)
    walls.append((time.perf_counter() - t0) * 1e3)
    kms.append(ms)
steady = min(kms)
print(json.dumps({"points": n, "steady_kernel_ms": steady, "first_8_wall_ms": [round(w, 2) for w in walls[:8]],
                  "calls_until_steady": next((i for i, m in enumerate(kms) if m < 1.05 * steady), None),
                  "ms_until_steady": round(sum(walls[:next((i for i, m in enumerate(kms) if m < 1.05 * steady), 0)]), 1),
                  "excess_ms_over_steady": round(sum(walls) - 40 * steady, 1)}))

"""Where the time of the pose search goes (BASELINE configs[2]): poses bucketed by their height above
the terrain under them, the kernel timed per bucket.  python tools/posit_profile.py [nposes_per_side]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
lrm = lrm_loader.load()
from importlib import import_module
fixtures = import_module("lrm_b200.fixtures")
dev = torch.device("cuda", 0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 256
terr = fixtures.perlin_terrain((1024, 1024), device=dev)
bod = fixtures.body_lattice(terr, m, m, m)
legs = [lrm.get_M2_leg(float(np.float32(k) * np.float32(2 * np.pi) / np.float32(4))) for k in range(4)]
quats = lrm.full_struct_orientations()
out, ms = lrm.positionability(bod, terr, legs, quats, timing=True)
out, ms_all = lrm.positionability(bod, terr, legs, quats, timing=True)
# terrain height under a pose: the map is a 1024 x 1024 lattice over the same xy extent (x-major)
H = terr[:, 2].reshape(1024, 1024)
lo, hi = terr.min(dim=0).values, terr.max(dim=0).values
ix = ((bod[:, 0] - lo[0]) / (hi[0] - lo[0]) * 1023).round().long().clamp(0, 1023)
iy = ((bod[:, 1] - lo[1]) / (hi[1] - lo[1]) * 1023).round().long().clamp(0, 1023)
dz = bod[:, 2] - H[ix, iy]
edges = [-1e9, -100, 0, 50, 100, 150, 200, 250, 300, 350, 400, 450, 500, 600, 1e9]
rows = []
for a, b in zip(edges[:-1], edges[1:]):
    sel = (dz >= a) & (dz < b)
    n = int(sel.sum())
    if n == 0:
        continue
    sub = bod[sel].contiguous()
    o, _ = lrm.positionability(sub, terr, legs, quats, timing=True)
    o, t = lrm.positionability(sub, terr, legs, quats, timing=True)
    rows.append({"dz_mm": [a, b], "poses": n, "standable": int((o != 0).sum()), "kernel_ms": t, "us_per_pose": 1e3 * t / n,
                 "first_orientation_mean": float(o[o != 0].float().mean()) if int((o != 0).sum()) else None})
    print(rows[-1], flush=True)
print(json.dumps({"poses": int(bod.shape[0]), "kernel_ms_all": ms_all, "sum_ms": sum(r["kernel_ms"] for r in rows), "rows": rows}))

"""Debug helper (GPU box): run one golden case on the GPU and dump the points that exceed the
distance tolerance, with the GPU and reference vectors."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
from oracle.oracle import best
from tests import parity

key = sys.argv[1]
rname, az, qname, pname = key.split("_")
g = np.load("tests/golden/one_leg_golden.npz")
lrm = lrm_loader.load()
pts, leg_arr, q = g[f"pts_{pname}"], g[f"leg_{rname}_{az}"], g[f"quat_{qname}"]
leg = lrm.LegDimensions.from_array(leg_arr)
d, f = lrm.distance(torch.from_numpy(pts).cuda(), leg, q)
d = d.cpu().numpy()
want = g[f"dist_{key}"]
err = np.abs(d - want).max(1)
bad = np.nonzero(err > 1e-2)[0]
o = best()
print(parity.dist_report(pts, d, want, lambda p: o.dist(p, leg_arr, q, 8)[0]))
for b in bad:
    print(int(b), pts[b].tolist(), d[b].tolist(), want[b].tolist(), float(np.linalg.norm(d[b])), float(np.linalg.norm(want[b])))
    land = (pts[b] - d[b]).astype(np.float32)[None]
    print("  landing gpu", o.dist(land, leg_arr, q)[0], " landing ref", o.dist((pts[b] - want[b]).astype(np.float32)[None], leg_arr, q)[0])

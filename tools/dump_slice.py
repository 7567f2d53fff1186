"""Debug helper (GPU box): the reference bench slice through the GPU, dump over-tolerance points."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
from oracle.oracle import best
from tests import parity

lrm = lrm_loader.load()
o = best()
def arange(start, end, step):
    out, v = [], np.float32(start)
    while v <= np.float32(end):
        out.append(v); v = np.float32(v + np.float32(step))
    return np.array(out, np.float32)
xs, zs = arange(-100, 601, 0.64), arange(-100, 51, 0.64)
X, Z = np.meshgrid(xs, zs, indexing="ij")
pts = np.stack([X, np.zeros_like(X), Z], -1).reshape(-1, 3).astype(np.float32)
leg = lrm.get_M2_leg(0.0); la = leg.as_array()
d, f = lrm.distance(torch.from_numpy(pts).cuda(), leg, None)
d = d.cpu().numpy()
want, _ = o.dist(pts, la, threads=8)
print(parity.dist_report(pts, d, want, lambda p: o.dist(p, la, threads=8)[0]))
err = np.abs(d - want).max(1)
for b in np.nonzero(err > 1e-2)[0]:
    land = (pts[b] - d[b]).astype(np.float32)[None]
    print(int(b), pts[b].tolist(), d[b].tolist(), want[b].tolist(), float(np.linalg.norm(d[b])), float(np.linalg.norm(want[b])),
          "landing", np.linalg.norm(o.dist(land, la)[0]))

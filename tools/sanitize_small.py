"""Small end-to-end exercise of every kernel on ragged sizes (written for compute-sanitizer; that
tool is closed on this GPU pool, so it runs as a plain crash / launch-error check)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
from tests import terrain

lrm = lrm_loader.load()
leg = lrm.get_M2_leg(0.3)
rng = np.random.default_rng(0)
for n in (1, 17, 1023, 5000, (1 << 22) + 37):          # ragged sizes; the last one takes the atlas path
    pts = torch.from_numpy(rng.uniform(-600, 600, (n, 3)).astype(np.float32)).cuda()
    lrm.reachability(pts, leg)
    lrm.distance(pts, leg)
    f, v = lrm.reach_dist(pts, leg, np.array([-0.96194, -0.03806, -0.19134, -0.19134], np.float32))
    planes = [pts[:, k].clone() for k in range(3)]   # fresh (16-byte aligned) planes
    lrm.reach_dist_soa(*planes, leg)
lrm.reach_dist(rng.uniform(-600, 600, (3000, 3)).astype(np.float32), leg)       # host staging
lrm.forward_kinematics(rng.uniform(-1, 1, (1000, 3)).astype(np.float32), leg)
lat = torch.empty((1000, 3), device="cuda")
lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (10, 10, 10))
lrm.make_lattice(lat, lo, step, dims)
terr = terrain.sine_terrain(33, 500.0, 50.0)
bodies = terrain.body_lattice(terr, 6, 6, 6, z_above=300.0)
legs = [lrm.get_M2_leg(k * 1.5707964) for k in range(4)]
lrm.positionability(torch.from_numpy(bodies).cuda(), torch.from_numpy(terr).cuda(), legs)
lrm.positionability(bodies, terr, legs, pre_cull=True)
wide = lrm.get_M2_leg(0.0); wide.max_angle_coxa, wide.min_angle_coxa = 3.0, -3.0
lrm.apply_oct(torch.from_numpy(terrain.sine_terrain(17, 1200.0, 80.0)).cuda(), wide, 5)
lrm.apply_recurs(torch.from_numpy(rng.uniform(-900, 900, (3000, 3)).astype(np.float32)).cuda(), leg, 6)
torch.cuda.synchronize()
print("sanitize_small: all kernels ran")

"""Debug helper (GPU box): dump standability mismatches of the Perlin-128 test scene."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrm_loader
from oracle.oracle import PortOracle
from tests import parity, terrain
lrm = lrm_loader.load(); port = PortOracle()
terr = terrain.perlin_terrain(128)
bodies = terrain.body_lattice(terr, 24, 48, 20)
legs_o = [port.get_leg(1, float(np.float32(k) * np.float32(2) * np.float32(np.pi) / np.float32(4))) for k in range(4)]
legs = [lrm.LegDimensions.from_array(l) for l in legs_o]
quats = lrm.full_struct_orientations()
want = port.standability(bodies, terr, legs_o, quats, threads=16)
got = lrm.positionability(torch.from_numpy(bodies).cuda(), torch.from_numpy(terr).cuda(), legs, quats).cpu().numpy()
rep = parity.pose_report(bodies, got, want, lambda p: port.standability(p, terr, legs_o, quats, threads=16))
print(rep)
bad = np.nonzero(got != want)[0]
np.savez("gpurun_out/posit_dump.npz", got=got, want=want, bad=bad)

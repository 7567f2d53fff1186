"""Generate the golden vectors in this directory from the COMPILED REFERENCE (oracle/_ref).

Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Everything written here is produced by calling the reference's own host functions through
oracle/ref_shim.cu; the files pin oracle/oracle_port.c (tests/test_oracle_golden.py) and are the
fixtures of the GPU parity tests on machines without the reference tree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import RefOracle  # noqa: E402

QUATS = {
    "identity": [1, 0, 0, 0],
    # RPYtoQuat(-pi/4, -pi/8, -pi/8) and friends: octree_util.cu.h:164-198
    "tilt0": None,
    "tilt9": None,
    "y10deg": [0.985, 0, 0.174, 0],  # settings.h:55 (commented quatTest candidates)
}


def lattice(n, lo=(-100, -400, -500), hi=(600, 400, 200)):
    axes = [np.float32(l) + np.arange(n, dtype=np.float32) * np.float32((np.float32(h) - np.float32(l)) / np.float32(n - 1))
            for l, h in zip(lo, hi)]
    X, Y, Z = np.meshgrid(*axes, indexing="ij")
    return np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)


def main():
    R = RefOracle()
    pi = np.float32(np.pi)
    QUATS["tilt0"] = R.rpy_to_quat(-pi / 4, -pi / 8, -pi / 8)
    QUATS["tilt9"] = R.rpy_to_quat(-pi / 4, -pi / 8, 0.0)
    rng = np.random.default_rng(20261018)

    out = {}
    # ---- one-leg: lattice (config C1 shape at 18^3), y=0 slice of the reference bench, random cloud
    grid = lattice(18)
    xs = np.arange(-100, 601, 7.3, dtype=np.float32)
    zs = np.arange(-100, 51, 3.1, dtype=np.float32)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    slice_y0 = np.stack([X, np.zeros_like(X), Z], -1).reshape(-1, 3).astype(np.float32)
    cloud = rng.uniform(-650, 650, (6000, 3)).astype(np.float32)
    out["pts_grid"], out["pts_slice"], out["pts_cloud"] = grid, slice_y0, cloud
    for robot, rname in ((0, "moonbot"), (1, "m2")):
        for az in (0.0, 2.3561945):
            leg = R.get_leg(robot, az)
            out[f"leg_{rname}_{az:.2f}"] = leg
            for qname, q in QUATS.items():
                if az != 0.0 and qname in ("tilt9",):
                    continue
                q = np.asarray(q, np.float32)
                out[f"quat_{qname}"] = q
                for pname in ("grid", "slice", "cloud"):
                    if qname != "identity" and pname == "slice":
                        continue
                    pts = out[f"pts_{pname}"]
                    key = f"{rname}_{az:.2f}_{qname}_{pname}"
                    out[f"reach_{key}"] = R.reach(pts, leg, q, threads=8)
                    d, f = R.dist(pts, leg, q, threads=8)
                    out[f"dist_{key}"], out[f"dflag_{key}"] = d, f
    # SURVEY §8c spot values
    spot = np.array([[300, 0, -100], [400, 0, -200], [346.5, 0, 0], [50, 0, 0], [300, 100, -150],
                     [-50, 20, -80]], np.float32)
    out["pts_spot"] = spot
    for robot, rname in ((0, "moonbot"), (1, "m2")):
        leg = R.get_leg(robot, 0.0)
        out[f"reach_spot_{rname}"] = R.reach(spot, leg)
        out[f"dist_spot_{rname}"], _ = R.dist(spot, leg)

    # ---- planar tables (find_region / insert_circles / insert_intersecv2), both legs + oriented
    plane = rng.uniform(-320, 320, (400, 2)).astype(np.float32)
    out["plane_pts"] = plane
    for robot, rname in ((0, "moonbot"), (1, "m2")):
        for qname in ("identity", "tilt0"):
            leg = R.rotate_leg_data(QUATS[qname], R.get_leg(robot, 0.7853982))
            out[f"oriented_leg_{rname}_{qname}"] = leg
            out[f"region_{rname}_{qname}"] = np.array(
                [R.find_region(float(x), float(y), leg) for x, y in plane], np.int32)
            out[f"circles_{rname}_{qname}"] = np.stack(
                [R.insert_circles(float(x), float(y), leg) for x, y in plane[:64]])
            out[f"corners_{rname}_{qname}"] = R.insert_intersec(leg)

    # ---- quaternion helpers
    out["rpy_samples"] = rng.uniform(-1.2, 1.2, (32, 3)).astype(np.float32)
    out["rpy_quats"] = np.stack([R.rpy_to_quat(*map(float, r)) for r in out["rpy_samples"]])
    vecs = rng.uniform(-500, 500, (32, 3)).astype(np.float32)
    out["rot_vecs"] = vecs
    out["rot_out"] = np.stack([R.qt_rotate(q, v) for q, v in zip(out["rpy_quats"], vecs)])
    legs_rot = []
    for q in out["rpy_quats"]:
        legs_rot.append(R.rotate_leg_data(q, R.get_leg(1, 1.1)))
    out["rotated_legs_m2_az1.1"] = np.stack(legs_rot)
    # robot_full_struct's 45 orientations (several_leg.cu:811-857) rebuilt from reference primitives
    q_init = R.quat_from_vect_angle([0, 0, 1], 0.0)
    quats = []
    for i in range(3):
        roll = -pi / 8 + (pi / 8 - -pi / 8) * (np.float32(i) / np.float32(2))
        qr = R.qt_multiply(R.quat_from_vect_angle([1, 0, 0], float(roll)), q_init)
        for j in range(3):
            pitch = -pi / 8 + (pi / 8 - -pi / 8) * (np.float32(j) / np.float32(2))
            qp = R.qt_multiply(R.quat_from_vect_angle([0, 1, 0], float(pitch)), qr)
            for m in range(5):
                yaw = np.float32(0) + (pi / 2 - np.float32(0)) * (np.float32(m) / np.float32(4))
                quats.append(R.qt_multiply(R.quat_from_vect_angle([0, 0, 1], float(yaw)), qp))
    out["full_struct_quats"] = np.stack(quats)

    # ---- octree child boxes (octree_util.cu.h:105-151)
    parents = np.array([[0, 0, 0, 5000, 5000, 5000], [10, -20, 30, 150, 80, 150],
                        [0, 0, 0, 90, 90, 300], [5, 5, 5, 60, 70, 80]], np.float32)
    out["box_parents"] = parents
    cb = []
    for p in parents:
        for c in range(8):
            r, box, missing = R.create_child_box(p, c)
            cb.append(np.concatenate([box, [r, missing]]))
    out["box_children"] = np.array(cb, np.float32)

    np.savez_compressed(os.path.join(HERE, "one_leg_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "one_leg_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()

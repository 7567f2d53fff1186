"""CPU-only: pin the C restatement (oracle/oracle_port.c) against the golden vectors that
tests/golden/make_golden.py generated from the COMPILED REFERENCE, and — where the reference shim
is present (oracle/_ref) — against the reference itself on fresh inputs.  Bit-exact: both sides are
host float arithmetic with the same libm and no FMA contraction."""
import numpy as np
import pytest

QUAT_NAMES = ("identity", "tilt0", "tilt9", "y10deg")


def _cases(golden):
    for k in golden.files:
        if k.startswith("reach_") and not k.startswith("reach_spot"):
            rname, az, qname, pname = k[len("reach_"):].split("_")
            yield k[len("reach_"):], rname, az, qname, pname


def test_golden_file_shape(golden):
    assert len(list(_cases(golden))) >= 30
    assert golden["pts_grid"].shape == (18 ** 3, 3)


def test_port_matches_golden_one_leg(golden, port):
    n_checked = 0
    for key, rname, az, qname, pname in _cases(golden):
        leg = golden[f"leg_{rname}_{az}"]
        q = golden[f"quat_{qname}"]
        pts = golden[f"pts_{pname}"]
        r = port.reach(pts, leg, q, threads=4)
        d, f = port.dist(pts, leg, q, threads=4)
        assert np.array_equal(r, golden[f"reach_{key}"]), key
        assert np.array_equal(f, golden[f"dflag_{key}"]), key
        assert np.array_equal(d.view(np.uint32), golden[f"dist_{key}"].view(np.uint32)), key
        n_checked += len(pts)
    assert n_checked > 150_000


def test_port_default_legs(golden, port):
    assert np.array_equal(port.get_leg(0, 0.0).view(np.uint32), golden["leg_moonbot_0.00"].view(np.uint32))
    assert np.array_equal(port.get_leg(1, 0.0).view(np.uint32), golden["leg_m2_0.00"].view(np.uint32))
    assert np.array_equal(port.get_leg(1, 2.3561945).view(np.uint32), golden["leg_m2_2.36"].view(np.uint32))
    m2 = port.get_leg(1, 0.0)
    # static_variables.cpp:69-93: M2 = body 181, pitch -45 deg, coxa 65.5, femur 129, tibia 135
    assert m2[1] == 181 and m2[3] == 65.5 and m2[4] == 135 and m2[5] == 129
    assert abs(m2[2] + np.pi / 4) < 1e-6 and abs(m2[6] - 0.6981317) < 1e-6 and abs(m2[7] + 2.268928) < 1e-6


def test_spot_values_from_survey(golden, port):
    """SURVEY.md §8c known answers (M2 leg, identity orientation), mm."""
    leg = port.get_leg(1, 0.0)
    d, _ = port.dist(golden["pts_spot"], leg)
    r = port.reach(golden["pts_spot"], leg)
    expect = np.array([[-33.576088, 0, 24.799173], [-24.525362, 0, 21.826902], [-43.406654, 0, 69.687981],
                       [48.387573, 0, 149.298279], [9.760560, 8.698915, -13.639524],
                       [-51.869598, 16.082787, 69.555466]], np.float32)
    assert np.abs(d - expect).max() < 2e-5
    assert r.tolist() == [0, 1, 0, 0, 1, 0]
    assert np.array_equal(d.view(np.uint32), golden["dist_spot_m2"].view(np.uint32))
    dm, _ = port.dist(golden["pts_spot"], port.get_leg(0, 0.0))
    assert np.array_equal(dm.view(np.uint32), golden["dist_spot_moonbot"].view(np.uint32))


def test_port_planar_tables(golden, port):
    for rname in ("moonbot", "m2"):
        for qname in ("identity", "tilt0"):
            leg = golden[f"oriented_leg_{rname}_{qname}"]
            pts = golden["plane_pts"]
            reg = np.array([port.find_region(float(x), float(y), leg) for x, y in pts], np.int32)
            assert np.array_equal(reg, golden[f"region_{rname}_{qname}"])
            circ = np.stack([port.insert_circles(float(x), float(y), leg) for x, y in pts[:64]])
            assert np.array_equal(circ.view(np.uint32), golden[f"circles_{rname}_{qname}"].view(np.uint32))
            assert np.array_equal(port.insert_intersec(leg).view(np.uint32),
                                  golden[f"corners_{rname}_{qname}"].view(np.uint32))


def test_appendix_c_known_answers(port):
    """circle constants of the default legs."""
    m2 = port.get_leg(1, 0.0)
    c = port.insert_circles(100.0, -50.0, m2)  # a lower-region point
    assert abs(c[0, 2] - 132.102234) < 1e-4 and c[0, 3] == 0  # C_in, repulsive
    corners = port.insert_intersec(m2)
    assert corners.shape == (5, 2)
    assert np.allclose(corners[0], [116.9134, -61.5], atol=1e-3)
    assert np.allclose(corners[-1], [103.4160, 215.7763], atol=1e-3)
    assert port.insert_intersec(port.get_leg(0, 0.0)).shape == (4, 2)


def test_port_quaternion_helpers(golden, port):
    q = np.stack([port.rpy_to_quat(*map(float, r)) for r in golden["rpy_samples"]])
    assert np.array_equal(q.view(np.uint32), golden["rpy_quats"].view(np.uint32))
    out = np.stack([port.qt_rotate(a, v) for a, v in zip(golden["rpy_quats"], golden["rot_vecs"])])
    assert np.array_equal(out.view(np.uint32), golden["rot_out"].view(np.uint32))
    legs = np.stack([port.rotate_leg_data(a, port.get_leg(1, 1.1)) for a in golden["rpy_quats"]])
    assert np.array_equal(legs.view(np.uint32), golden["rotated_legs_m2_az1.1"].view(np.uint32))
    assert np.array_equal(port.full_struct_orientations().view(np.uint32),
                          golden["full_struct_quats"].view(np.uint32))


def test_quaternion_layout_quirks(port):
    """RPYtoQuat(0,0,0) = (-1,0,0,0) is qtRotate's identity; the 27
    octree angle samples hold only 8 distinct orientations and index 13 is the level one."""
    assert np.allclose(port.rpy_to_quat(0, 0, 0), [-1, 0, 0, 0], atol=1e-7)
    v = np.array([12.5, -3.0, 7.75], np.float32)
    assert np.array_equal(port.qt_rotate([1, 0, 0, 0], v), v)
    qs = np.stack([port.quaternion_from_angle_index(i) for i in range(27)])
    assert len({tuple(np.round(q, 5)) for q in qs}) == 8
    assert np.allclose(qs[13], [-1, 0, 0, 0], atol=1e-6)
    assert np.allclose(qs[0], [-0.87415, -0.40328, -0.10355, -0.25], atol=1e-4)
    assert np.allclose(qs[12], [-0.92388, -0.38268, 0, 0], atol=1e-4)
    # identity orientation leaves the leg limits untouched, bit for bit
    leg = port.get_leg(1, 0.0)
    assert np.array_equal(port.rotate_leg_data([1, 0, 0, 0], leg).view(np.uint32), leg.view(np.uint32))


def test_distance_lands_on_boundary(golden, port):
    """p - d(p) lies on the reachability edge, i.e. |d(p - d(p))| ~ 0."""
    leg = port.get_leg(1, 0.0)
    pts = golden["pts_grid"]
    d, _ = port.dist(pts, leg)
    d2, _ = port.dist(pts - d, leg)
    assert np.linalg.norm(d2, axis=1).max() < 2e-2


def test_manually_placed_points(port):
    """The stale Catch2 properties of one_leg.cpp:100-139,498-588 re-expressed on the live code:
    on the moonbot leg's x axis at z = 0... the workspace is entered at body+coxa+r and left
    1 mm further."""
    leg = port.get_leg(0, 0.0)  # pitch 0: the femur plane of y = 0 is the xz plane
    # walk outward along a ray that stays inside the coxa/femur limits and find the flips
    xs = np.arange(150, 700, 0.25, dtype=np.float32)
    pts = np.stack([xs, np.zeros_like(xs), np.full_like(xs, -150)], 1)
    r = port.reach(pts, leg)
    d, _ = port.dist(pts, leg)
    flips = np.nonzero(np.diff(r.astype(np.int8)))[0]
    assert len(flips) >= 1
    n = np.linalg.norm(d, axis=1)
    for f in flips:  # the distance field vanishes where the flag flips
        assert min(n[f], n[f + 1]) < 0.26
    # overshoot by 0.01 mm along x past the outermost flip: |d| ~ 0.01 within 1e-3 (one_leg.cpp:498-588)
    f = flips[-1]
    lo, hi = float(xs[f]), float(xs[f + 1])
    for _ in range(40):
        mid = np.float32((lo + hi) / 2)
        if port.reach(np.array([[mid, 0, -150]], np.float32), leg)[0]:
            lo = float(mid)
        else:
            hi = float(mid)
    p = np.array([[hi + 0.01, 0, -150]], np.float32)
    dd, _ = port.dist(p, leg)
    assert abs(np.linalg.norm(dd) - 0.01) < 2e-3


def test_forward_kinematics_points_are_reachable(port):
    """one_leg.cpp:141-202 with the tibia-absolute filter and the coxa pitch."""
    rng = np.random.default_rng(7)
    for robot in (0, 1):
        L = port.get_leg(robot, 0.0)
        (body, pitch, coxa_len, tib_len, fem_len, abs_pos, abs_neg) = L[1], L[2], L[3], L[4], L[5], L[6], L[7]
        n = 4000
        cox = rng.uniform(L[9] + 1e-3, L[8] - 1e-3, n)
        fem = rng.uniform(L[13] + 1e-3, L[12] - 1e-3, n)
        tib = rng.uniform(L[11] + 1e-3, L[10] - 1e-3, n)
        ok = (fem + tib > abs_neg + 1e-3) & (fem + tib < abs_pos - 1e-3)
        cox, fem, tib = cox[ok], fem[ok], tib[ok]
        # FK in the coxa frame, then undo place_over_coxa (one_leg.cu:9-24)
        rad = coxa_len + fem_len * np.cos(fem) + tib_len * np.cos(fem + tib)
        up = fem_len * np.sin(fem) + tib_len * np.sin(fem + tib)
        xc, yc, zc = rad * np.cos(cox), rad * np.sin(cox), up
        c, s = np.cos(pitch), np.sin(pitch)
        x = xc * c - zc * s + body
        z = xc * s + zc * c
        pts = np.stack([x, yc, z], 1).astype(np.float32)
        assert port.reach(pts, L).all()


def test_port_matches_compiled_reference_fresh(ref, port):
    rng = np.random.default_rng(99)
    pts = rng.uniform(-700, 700, (60000, 3)).astype(np.float32)
    for robot, az, q in ((1, 0.3, [1, 0, 0, 0]), (0, 4.0, [-0.90613, -0.37533, 0.07466, -0.18024]),
                         (1, 1.5707964, [0.940, 0, 0, 0.342])):
        leg = ref.get_leg(robot, az)
        assert np.array_equal(ref.reach(pts, leg, q, 8), port.reach(pts, leg, q, 8))
        dr, fr = ref.dist(pts, leg, q, 8)
        dp, fp = port.dist(pts, leg, q, 8)
        assert np.array_equal(dr.view(np.uint32), dp.view(np.uint32)) and np.array_equal(fr, fp)


def test_reference_cpu_path_as_shipped(ref, golden):
    """apply_reach_cpu / apply_dist_cpu (cross_compiled.cu:163-181) == the slab loop at quatTest."""
    leg = ref.get_leg(1, 0.0)
    pts = golden["pts_slice"]
    r, ms = ref.apply_reach_cpu(pts, leg)
    d, ms2 = ref.apply_dist_cpu(pts, leg)
    assert ms > 0 and ms2 > 0
    assert np.array_equal(r, golden["reach_m2_0.00_identity_slice"])
    assert np.array_equal(d.view(np.uint32), golden["dist_m2_0.00_identity_slice"].view(np.uint32))

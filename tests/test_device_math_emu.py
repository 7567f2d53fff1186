"""CPU-only: the product's per-point device math (leg_math.cuh) and leg-plan construction
(leg_plan.cpp), compiled for the host by tests/emu (test infrastructure, not part of the product
library), against the golden vectors and the oracle — the same bar as the GPU parity tests.  This
keeps algorithmic regressions out of the GPU budget; the GPU tests remain the parity tests proper."""
import ctypes
import os

import numpy as np
import pytest

from tests import parity


@pytest.fixture(scope="module")
def emu():
    from tests.emu.build_emu import build
    E = ctypes.CDLL(build())
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    E.emu_dist.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    E.emu_reach.argtypes = [vp, sz, vp, vp, vp]
    E.emu_leg_reaches.argtypes = [vp, sz, vp, vp, vp]
    return E


def run_emu(E, pts, leg, q):
    pts = np.ascontiguousarray(pts, np.float32)
    leg = np.ascontiguousarray(leg, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    n = len(pts)
    r = np.zeros(n, np.uint8)
    E.emu_reach(pts.ctypes.data, n, leg.ctypes.data, q.ctypes.data, r.ctypes.data)
    d = np.zeros_like(pts)
    f = np.zeros(n, np.uint8)
    fr = np.zeros(n, np.uint8)
    E.emu_dist(pts.ctypes.data, n, leg.ctypes.data, q.ctypes.data, d.ctypes.data, f.ctypes.data, fr.ctypes.data)
    return r, d, f, fr


def _cases(golden):
    for k in golden.files:
        if k.startswith("reach_") and not k.startswith("reach_spot"):
            rname, az, qname, pname = k[len("reach_"):].split("_")
            yield k[len("reach_"):], rname, az, qname, pname


def test_emulated_device_math_vs_golden(emu, golden, port):
    total = {"n": 0, "flag_mismatch": 0, "over_tol": 0}
    for key, rname, az, qname, pname in _cases(golden):
        pts, leg, q = golden[f"pts_{pname}"], golden[f"leg_{rname}_{az}"], golden[f"quat_{qname}"]
        r, d, f, fr = run_emu(emu, pts, leg, q)
        reach_fn = lambda p: port.reach(p, leg, q, threads=4)
        for got in (r, fr):
            rep = parity.flag_report(pts, got, golden[f"reach_{key}"], reach_fn)
            assert rep["unexplained"] == 0, (key, rep)
            total["flag_mismatch"] += rep["mismatch"]
        rep = parity.flag_report(pts, f, golden[f"dflag_{key}"], lambda p: port.dist(p, leg, q, threads=4)[1])
        assert rep["unexplained"] == 0, (key, rep)
        rep = parity.dist_report(pts, d, golden[f"dist_{key}"], lambda p: port.dist(p, leg, q, threads=4)[0],
                                 frame_slack=abs(float((q.astype(np.float64) ** 2).sum()) - 1.0))
        assert rep["unexplained"] == 0, (key, rep)
        total["over_tol"] += rep["over_tol"]
        total["n"] += len(pts)
    assert total["flag_mismatch"] <= total["n"] // 50000 + 2, total
    assert total["over_tol"] <= total["n"] // 2000, total


def test_emulated_leg_predicate_of_positionability(emu, port):
    """reachable_rotate_leg (several_leg.cu:48-67) as the positionability kernel evaluates it."""
    rng = np.random.default_rng(8)
    off = rng.uniform(-550, 550, (60000, 3)).astype(np.float32)
    quats = port.full_struct_orientations()
    for o in (0, 7, 22, 44):
        for az in (0.0, 1.5707964, 3.1415927, 4.712389):
            leg = port.get_leg(1, az)
            q = quats[o]
            got = np.zeros(len(off), np.uint8)
            emu.emu_leg_reaches(off.ctypes.data, len(off), leg.ctypes.data, q.ctypes.data, got.ctypes.data)
            # oracle: gravity-side test + Rz(-azimuth) + reachability_circles with rotated limits
            rl = port.rotate_leg_data(q, leg)
            qi = np.array([q[0], -q[1], -q[2], -q[3]], np.float32) / np.float32((q * q).sum())
            g = np.stack([port.qt_rotate(qi, v) for v in off[:4000]])
            c, s = np.cos(-az), np.sin(-az)
            gx = g[:, 0] * c - g[:, 1] * s
            v = off[:4000].copy()
            vx = v[:, 0] * c - v[:, 1] * s
            vy = v[:, 0] * s + v[:, 1] * c
            rl0 = rl.copy()
            rl0[0] = 0.0   # body_angle already applied; reachability_circles ignores it anyway
            want = port.reach(np.stack([vx, vy, v[:, 2]], 1).astype(np.float32), rl0, [1, 0, 0, 0])
            # reachability_global at identity adds only an exact Rz(0): same as reachability_circles
            want = want & (gx >= 0)
            assert (got[:4000] != want).sum() <= 2, (o, az, int((got[:4000] != want).sum()))


def test_arc_tables_cover_default_legs_and_match_cross_validation(emu, port):
    """The precomputed valid arcs (leg_plan.cpp valid_arcs) replace the explicit cross-validation of
    multi_circle_clamp (one_leg.cu:122-123): for the default legs under every orientation of both
    sample sets every valid set must be a single arc, and both paths must agree."""
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    emu.emu_plan_is_generic.argtypes = [vp, vp]
    emu.emu_dist_generic.argtypes = [vp, sz, vp, vp, vp]
    rng = np.random.default_rng(21)
    pts = rng.uniform(-650, 650, (40000, 3)).astype(np.float32)
    quats = list(port.full_struct_orientations()[::6]) + [port.quaternion_from_angle_index(i) for i in (0, 1, 3, 13)]
    worst = 0
    for robot in (0, 1):
        for az in (0.0, 0.7853982, 3.9269907):
            leg = port.get_leg(robot, az)
            for q in quats:
                q = np.ascontiguousarray(q, np.float32)
                assert emu.emu_plan_is_generic(leg.ctypes.data, q.ctypes.data) == 0
            q = np.ascontiguousarray(quats[robot + 1], np.float32)
            _, d, _, _ = run_emu(emu, pts, leg, q)
            g = np.zeros_like(pts)
            emu.emu_dist_generic(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, g.ctypes.data)
            diff = np.abs(d - g).max(axis=1)
            worst = max(worst, int((diff > 1e-2).sum()))
            assert (diff > 1e-2).sum() <= 4, (robot, az, int((diff > 1e-2).sum()))


def test_fast_path_never_changes_a_result(emu, port):
    """The fast path of the distance sweep (leg_math.cuh dist_fast: yaw-sector table + plane atlas)
    is a pure accelerator: wherever both tables certify a point it must reproduce the full
    evaluation (flags identical, vectors to float rounding: a certified cell evaluates
    P - projection as v (1 - r/|v|) instead of P - (c + v r/|v|)), and enough of both tables must
    be certified to be worth having.
    The cloud includes rings of points straddling every yaw decision (limits, limits +- pi/2, the
    +-pi seam, the x axis) at several radii."""
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    emu.emu_dist_atlas.argtypes = [vp, sz, vp, vp, ctypes.c_int, ctypes.c_float, vp, vp, vp, vp, vp, vp]
    emu.emu_dist_atlas.restype = ctypes.c_size_t
    rng = np.random.default_rng(31)
    cloud = np.concatenate([rng.uniform([-100, -400, -500], [600, 400, 200], (60000, 3)),
                            rng.uniform(-700, 700, (30000, 3))]).astype(np.float32)
    cases = ((1, 0.0, [1, 0, 0, 0]), (0, 0.7853982, port.quaternion_from_angle_index(0)),
             (1, 3.9269907, port.full_struct_orientations()[31]))
    for robot, az, q in cases:
        leg = port.get_leg(robot, az)
        q = np.ascontiguousarray(q, np.float32)
        # rings around the coxa axis in the world frame of an identity-orientation leg would only
        # hit the boundaries for az = 0; build them in the coxa frame and map back instead:
        # world = R^T (p_coxa - t) is not needed exactly — any dense set of directions does, so
        # sweep the azimuth densely around the leg's mount point at several heights.
        ang = np.linspace(-np.pi, np.pi, 7201)
        near = np.concatenate([ang + d for d in (-1e-4, -1e-6, 0.0, 1e-6, 1e-4)])
        rings = []
        for rad, z in ((40.0, -50.0), (200.0, -120.0), (420.0, 30.0)):
            cx, cy = np.float32(leg[1]) * np.cos(az), np.float32(leg[1]) * np.sin(az)
            rings.append(np.stack([cx + rad * np.cos(near), cy + rad * np.sin(near), np.full_like(near, z)], 1))
        pts = np.ascontiguousarray(np.concatenate([cloud] + rings), np.float32)
        r0, base, bf, br = run_emu(emu, pts, leg, q)
        out = np.zeros_like(pts)
        fl = np.zeros(len(pts), np.uint8)
        rf = np.zeros(len(pts), np.uint8)
        ra = np.zeros(len(pts), np.uint8)
        pure, bins = ctypes.c_size_t(0), ctypes.c_size_t(0)
        fb = emu.emu_dist_atlas(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, 1024, 2.0,
                                out.ctypes.data, fl.ctypes.data, rf.ctypes.data, ctypes.byref(pure),
                                ctypes.byref(bins), ra.ctypes.data)
        assert np.array_equal(fl, bf) and np.array_equal(rf, br)
        assert np.array_equal(ra, r0)   # reach-only sweep through the atlas' valid bit
        assert np.array_equal(out, base), float(np.abs(out - base).max())   # same operations on the winner: same bits
        assert pure.value > 0.85 * 1024 * 1024 and fb < 0.35 * len(pts), (pure.value, fb)
        assert bins.value > 0.95 * 1025, bins.value


def test_choice_volume_never_changes_a_result(emu, port):
    """The three-tier distance sweep (leg_math.cuh: choice volume -> dist_fast -> full evaluation) is
    a pure accelerator and, since every tier applies the same operations to the winning candidate,
    returns the full evaluation's result BIT FOR BIT whichever tier decides a point — so a ring
    overflow (which moves a point to a slower tier) cannot change an output either.  Cube bytes come
    from the per-cube function of the device build (choice_cell_byte; the device additionally lets a
    block of 4^3 cubes inherit a byte certified for the whole block by the same function)."""
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    emu.emu_dist_choice.argtypes = [vp, sz, vp, vp, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                    ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
    rng = np.random.default_rng(41)
    idx = rng.integers(0, 1000, (120000, 3))
    lo, hi = np.array([-100, -400, -500], np.float32), np.array([600, 400, 200], np.float32)
    lattice = lo + idx.astype(np.float32) * ((hi - lo) / np.float32(999)).astype(np.float32)
    cloud = np.concatenate([lattice, rng.uniform(-800, 800, (60000, 3))]).astype(np.float32)
    cases = ((1, 0.0, [1, 0, 0, 0], 4.0, 0.80), (0, 0.7853982, port.quaternion_from_angle_index(0), 4.0, 0.70),
             (1, 3.9269907, port.full_struct_orientations()[31], 2.0, 0.70))
    for robot, az, q, cell, want_share in cases:
        leg = port.get_leg(robot, az)
        q = np.ascontiguousarray(q, np.float32)
        ang = np.linspace(-np.pi, np.pi, 3601)
        near = np.concatenate([ang + d for d in (-1e-4, 0.0, 1e-4)])
        cx, cy = np.float32(leg[1]) * np.cos(az), np.float32(leg[1]) * np.sin(az)
        ring = np.stack([cx + 200.0 * np.cos(near), cy + 200.0 * np.sin(near), np.full_like(near, -120.0)], 1)
        pts = np.ascontiguousarray(np.concatenate([cloud, ring]), np.float32)
        r0, base, bf, br = run_emu(emu, pts, leg, q)
        out = np.zeros_like(pts)
        fl = np.zeros(len(pts), np.uint8)
        rf = np.zeros(len(pts), np.uint8)
        rv = np.zeros(len(pts), np.uint8)
        tier = np.zeros(len(pts), np.uint8)
        tiers = (ctypes.c_size_t * 4)()
        known = ctypes.c_size_t(0)
        emu.emu_dist_choice(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, 2048, 1.0, cell,
                            int(1536 / cell), out.ctypes.data, fl.ctypes.data, rf.ctypes.data, tiers,
                            tier.ctypes.data, rv.ctypes.data, ctypes.byref(known))
        assert np.array_equal(fl, bf) and np.array_equal(rf, br), (robot, az)
        # the reach-only sweep through the cube's reach bits equals reachability_circles, and most
        # of the cloud is decided by those bits alone
        assert np.array_equal(rv, r0), (robot, az, int((rv != r0).sum()))
        assert known.value > 0.8 * len(pts), (robot, az, known.value / len(pts))
        assert np.array_equal(out, base), (robot, az, float(np.abs(out - base).max()))
        # the volume must be worth having on the bench box (first 120 000 points)
        assert (tier[:120000] == 0).mean() > want_share, (robot, az, float((tier[:120000] == 0).mean()))


def test_tier0_labels_never_change_a_result(emu, port):
    """Tier 0 of the tiered sweep (16-bit cube texels: winning solution + plane label of the cube,
    certified per block by plane_probe and per cube by the atlas cells under the cube's plane
    rectangle — coarse_block_word / choice_cell_word, the functions of the device build) followed
    by the plane-atlas tier, the explicit plane evaluation, dist_fast and the full evaluation: the
    output equals the full evaluation's BIT FOR BIT whichever tier decides a point, the rule-free
    variant equals the ruled one wherever the kernel may take it, and tier 0 alone decides most of
    the bench lattice (the device runs 0.5 mm atlas cells and reaches 88 %, 95 % with bricks; 1 mm
    cells here).  With bricks the sweep reads a fine texel where the cube's own is a pointer."""
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    emu.emu_dist_tier0.argtypes = [vp, sz, vp, vp, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                   vp, vp, vp, vp, vp, ctypes.c_int]
    rng = np.random.default_rng(43)
    idx = rng.integers(0, 1000, (100000, 3))
    lo, hi = np.array([-100, -400, -500], np.float32), np.array([600, 400, 200], np.float32)
    lattice = lo + idx.astype(np.float32) * ((hi - lo) / np.float32(999)).astype(np.float32)
    cloud = np.concatenate([lattice, rng.uniform(-800, 800, (50000, 3))]).astype(np.float32)
    cases = ((1, 0.0, [1, 0, 0, 0], 3.0, 0.80), (0, 0.7853982, port.quaternion_from_angle_index(0), 4.0, 0.65),
             (1, 3.9269907, port.full_struct_orientations()[31], 2.0, 0.65))
    for robot, az, q, cell, want_share in cases:
        leg = port.get_leg(robot, az)
        q = np.ascontiguousarray(q, np.float32)
        pts = np.ascontiguousarray(cloud, np.float32)
        r0, base, bf, br = run_emu(emu, pts, leg, q)
        shares = []
        for bricks in (0, 1):
            out = np.zeros_like(pts)
            fl = np.zeros(len(pts), np.uint8)
            rf = np.zeros(len(pts), np.uint8)
            tier = np.zeros(len(pts), np.uint8)
            tiers = (ctypes.c_size_t * 5)()
            emu.emu_dist_tier0(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, 2048, 1.0, cell,
                               int(1536 / cell) // 4 * 4, out.ctypes.data, fl.ctypes.data, rf.ctypes.data, tiers,
                               tier.ctypes.data, bricks)
            # the rule-free variant never differed, nor did the two cubes a texture fetch next to a
            # cube face may return
            assert sum(tiers) == len(pts), (robot, az, bricks, list(tiers))
            assert np.array_equal(fl, bf) and np.array_equal(rf, br), (robot, az, bricks)
            assert np.array_equal(out, base), (robot, az, bricks, float(np.abs(out - base).max()))
            shares.append(float((tier[:100000] == 0).mean()))
            assert shares[-1] > want_share, (robot, az, bricks, shares, [t / len(pts) for t in tiers])
            assert all(t > 0 for t in tiers), list(tiers)               # every tier was exercised
        # bricks (4^3 fine cubes under every unsettled cube) settle a good part of what the grid leaves
        assert shares[1] > shares[0] + 0.25 * (1.0 - shares[0]), shares


def test_reach_plan_predicates(emu, port):
    """positionability.cu's per-foothold predicate on the compact ReachPlan equals the full-plan
    path, and the cell-level pruning test is conservative: whenever some point within rc of a
    centre is reachable, reach_ball_possible(centre, rc) must say so."""
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    emu.emu_reach_offset.argtypes = [vp, sz, vp, vp, ctypes.c_float, vp, vp]
    rng = np.random.default_rng(77)
    centres = rng.uniform(-560, 560, (6000, 3)).astype(np.float32)
    quats = port.full_struct_orientations()
    for robot, az, o in ((1, 0.0, 0), (1, 1.5707964, 20), (0, 3.1415927, 44), (1, 4.712389, 33)):
        leg = port.get_leg(robot, az)
        q = np.ascontiguousarray(quats[o], np.float32)
        for rc in (20.0, 60.0, 150.0):
            r0 = np.zeros(len(centres), np.uint8)
            ball = np.zeros(len(centres), np.uint8)
            emu.emu_reach_offset(centres.ctypes.data, len(centres), leg.ctypes.data, q.ctypes.data, rc,
                                 r0.ctypes.data, ball.ctypes.data)
            full = np.zeros(len(centres), np.uint8)
            emu.emu_leg_reaches(centres.ctypes.data, len(centres), leg.ctypes.data, q.ctypes.data, full.ctypes.data)
            assert np.array_equal(r0, full)
            # sample points inside each ball
            any_reach = r0.astype(bool).copy()
            for _ in range(24):
                d = rng.normal(size=(len(centres), 3))
                d *= (rng.uniform(0, 1, (len(centres), 1)) ** (1 / 3)) * rc / np.linalg.norm(d, axis=1, keepdims=True)
                pts = (centres + d).astype(np.float32)
                rr = np.zeros(len(pts), np.uint8)
                bb = np.zeros(len(pts), np.uint8)
                emu.emu_reach_offset(pts.ctypes.data, len(pts), leg.ctypes.data, q.ctypes.data, 1.0,
                                     rr.ctypes.data, bb.ctypes.data)
                any_reach |= rr.astype(bool)
            assert not (any_reach & ~ball.astype(bool)).any(), (robot, az, o, rc)
            # ... and it does prune: most far-away balls are rejected
            assert ball.mean() < 0.9


def test_octree_leg_pruning_is_conservative(emu, port):
    """octree.cu skips a leg's distance evaluation for a foothold when leg_ball_possible(v, rc) is
    false, rc = the child's half diagonal: that is exact iff no point within rc of v is reachable
    by that leg (flag of distance_global) — then the leg cannot reach v, and its distance vector,
    which ends on the workspace boundary, is longer than rc, so it cannot land inside the child."""
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    emu.emu_leg_ball.argtypes = [vp, sz, vp, vp, ctypes.c_float, vp]
    rng = np.random.default_rng(90)
    centres = rng.uniform(-700, 700, (20000, 3)).astype(np.float32)
    wide = port.get_leg(1, 0.0).copy()
    wide[11], wide[12] = 3.0, -3.0          # max / min coxa angle: the wide-coxa leg of the octree tests
    for leg, q in ((port.get_leg(1, 0.7853982), port.quaternion_from_angle_index(0)),
                   (port.get_leg(0, 2.3561945), port.quaternion_from_angle_index(13)), (wide, [1, 0, 0, 0])):
        q = np.ascontiguousarray(q, np.float32)
        leg = np.ascontiguousarray(leg, np.float32)
        for rc in (60.0, 140.0, 300.0):
            ball = np.zeros(len(centres), np.uint8)
            emu.emu_leg_ball(centres.ctypes.data, len(centres), leg.ctypes.data, q.ctypes.data, rc, ball.ctypes.data)
            pruned = ball == 0
            assert 0.2 < pruned.mean() < 0.99, (rc, float(pruned.mean()))   # it prunes, and not everything
            # the centre itself: unreachable, and its distance vector is longer than rc
            _, d, f, _ = run_emu(emu, centres[pruned], leg, q)
            assert not f.any()
            assert (np.linalg.norm(d, axis=1) > rc).all(), float(np.linalg.norm(d, axis=1).min())
            # no reachable point anywhere in a pruned ball
            for _ in range(8):
                step = rng.normal(size=(int(pruned.sum()), 3))
                step *= (rng.uniform(0, 1, (len(step), 1)) ** (1 / 3)) * rc / np.linalg.norm(step, axis=1, keepdims=True)
                _, _, f2, _ = run_emu(emu, (centres[pruned] + step).astype(np.float32), leg, q)
                assert not f2.any(), (rc, int(f2.sum()))


def test_host_plan_bytes_match_the_golden_file(emu):
    """The host-built constants (LegPlan: affine maps, angle tests, circle tables, valid arcs,
    corners; FastTables: yaw-sector pairs and bin codes) byte for byte against
    tests/golden/leg_plan_golden.npz (captured on x86-64 with -ffp-contract=off): a host compiler
    that contracts a*b+c, or any edit of leg_plan.cpp / fast_tables.cpp that moves a rounding,
    shows up here — the certified tables and the knife-edge rules assume these exact bits."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "leg_plan_golden.npz"))
    keys = [k[len("plan_"):] for k in gold.files if k.startswith("plan_")]
    assert len(keys) >= 4
    for k in keys:
        plan = np.zeros(emu.emu_sizeof_plan(), np.uint8)
        tables = np.zeros(emu.emu_sizeof_tables(), np.uint8)
        leg, q = np.ascontiguousarray(gold["leg_" + k]), np.ascontiguousarray(gold["quat_" + k])
        emu.emu_plan_bytes(leg.ctypes.data, q.ctypes.data, plan.ctypes.data, tables.ctypes.data)
        assert np.array_equal(plan, gold["plan_" + k]), k
        assert np.array_equal(tables, gold["tables_" + k]), k


def test_cone_collision_region_is_conservative(emu, port):
    """positionability.cu settles a pose in one walk when some map point is inside the body cylinder
    under EVERY orientation.  The region it tests (AxisCone: the cylinder axes of all orientations lie
    within an angle of their mean) must never claim an offset that escapes the cylinder under one of
    them — and should claim most of what the brute force finds, far more than the 109 mm ball did."""
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    emu.emu_cone_check.argtypes = [vp, ctypes.c_int, ctypes.c_float, vp, sz, vp, vp]
    emu.emu_cone_check.restype = ctypes.c_int
    rng = np.random.default_rng(5)
    off = rng.uniform(-320, 320, (400_000, 3)).astype(np.float32)
    level = np.stack([port.rpy_to_quat(0.0, 0.0, float(y)) for y in np.linspace(0, 2 * np.pi, 16, endpoint=False)])
    wild = np.stack([port.rpy_to_quat(float(r), float(p), 0.3) for r in (-1.2, 0.0, 1.2) for p in (-1.2, 0.0, 1.2)])
    cases = (("robot_full_struct", port.full_struct_orientations(), 181.0, 3, 0.6),
             ("level yaws", level, 181.0, 1, 0.97), ("wide spread", wild, 181.0, 2, 0.0))
    for name, quats, radius, want_state, want_share in cases:
        q = np.ascontiguousarray(quats, np.float32)
        claimed = np.zeros(len(off), np.uint8)
        truth = np.zeros(len(off), np.uint8)
        state = emu.emu_cone_check(q.ctypes.data, len(q), radius, off.ctypes.data, len(off), claimed.ctypes.data,
                                   truth.ctypes.data)
        assert state == want_state, (name, state)                # ok | gate << 1
        assert not np.any(claimed & ~truth & 1), (name, int((claimed & ~truth & 1).sum()))
        if want_share:
            ball = (np.linalg.norm(off, axis=1) < 109.0) & (truth == 1)
            assert claimed.sum() >= want_share * truth.sum(), (name, int(claimed.sum()), int(truth.sum()))
            assert claimed.sum() > 1.5 * ball.sum(), (name, int(claimed.sum()), int(ball.sum()))


def test_plane_certificate_holds_inside_its_radius(emu, port):
    """plane_probe's safety — what the plane atlas and the cube labels are certified with, including the
    corner-on-circle certificate next to the arc ends — is a radius within which the plane evaluation
    keeps its valid bit, sector and winner, and plane_clamp returns the labelled winner's projection bit
    for bit: checked on random balls all over the femur plane, two legs, limits rotated by a body pitch."""
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    emu.emu_probe_ball_check.argtypes = [vp, sz, vp, vp, ctypes.c_int, ctypes.c_float, vp]
    emu.emu_probe_ball_check.restype = sz
    rng = np.random.default_rng(11)
    centres = np.ascontiguousarray(rng.uniform([-420, -520], [520, 420], (60_000, 2)), np.float32)
    for robot, az, q in ((1, 0.0, [1, 0, 0, 0]), (0, 0.7853982, port.quaternion_from_angle_index(0)),
                         (1, 3.9269907, port.full_struct_orientations()[31])):
        leg = port.get_leg(robot, az)
        q = np.ascontiguousarray(q, np.float32)
        counts = (ctypes.c_size_t * 3)()
        bad = emu.emu_probe_ball_check(centres.ctypes.data, len(centres), leg.ctypes.data, q.ctypes.data, 12, 6.0, counts)
        assert counts[0] > 0.9 * len(centres), (robot, az, list(counts))     # nearly every centre has a certificate
        assert bad == 0, (robot, az, int(bad), list(counts))

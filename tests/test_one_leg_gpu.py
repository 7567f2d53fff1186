"""GPU parity tests of the one-leg hot path, called through the C ABI (ctypes) and checked against
the oracle (compiled reference when shipped, else the pinned C port) and the committed golden
vectors.  Tolerances: flags bit-exact except within 1e-3 mm of the boundary (counted), distance
vectors within 1e-2 mm absolute (seam points counted) — BASELINE.json north_star."""
import ctypes

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _cases(golden):
    for k in golden.files:
        if k.startswith("reach_") and not k.startswith("reach_spot"):
            rname, az, qname, pname = k[len("reach_"):].split("_")
            yield k[len("reach_"):], rname, az, qname, pname


def _check(lrm, oracle, pts, leg_arr, q, want_r, want_d, want_f, label):
    leg = lrm.LegDimensions.from_array(leg_arr)
    dev = torch.from_numpy(pts).cuda()
    if q is None:
        q = np.array([1, 0, 0, 0], np.float32)
    r = lrm.reachability(dev, leg, q).cpu().numpy()
    d, f = lrm.distance(dev, leg, q)
    d, f = d.cpu().numpy(), f.cpu().numpy()
    fr, fd = lrm.reach_dist(dev, leg, q)
    fr, fd = fr.cpu().numpy(), fd.cpu().numpy()
    assert np.array_equal(fd.view(np.uint32), d.view(np.uint32)), label  # fused == separate
    reach_fn = lambda p: oracle.reach(p, leg_arr, q, threads=8)
    dist_fn = lambda p: oracle.dist(p, leg_arr, q, threads=8)[0]
    for name, got, want in (("reach", r, want_r), ("fused-reach", fr, want_r), ("dist-flag", f, want_f)):
        rep = parity.flag_report(pts, got, want, reach_fn if name != "dist-flag" else
                                 (lambda p: oracle.dist(p, leg_arr, q, threads=8)[1]))
        assert rep["unexplained"] == 0, (label, name, rep)
        assert rep["mismatch"] <= max(3, len(pts) // 20000), (label, name, rep)
    slack = 0.0 if q is None else abs(float((np.asarray(q, np.float64) ** 2).sum()) - 1.0)
    rep = parity.dist_report(pts, d, want_d, dist_fn, frame_slack=slack)
    assert rep["unexplained"] == 0, (label, rep)
    assert rep["over_tol"] <= max(3, len(pts) // 2000), (label, rep)
    return rep


def test_golden_vectors(lrm, oracle, golden):
    worst = 0.0
    for key, rname, az, qname, pname in _cases(golden):
        rep = _check(lrm, oracle, golden[f"pts_{pname}"], golden[f"leg_{rname}_{az}"],
                     golden[f"quat_{qname}"], golden[f"reach_{key}"], golden[f"dist_{key}"],
                     golden[f"dflag_{key}"], key)
        worst = max(worst, rep["max_err_within_tol"])
    assert worst < 1e-2


def test_spot_values(lrm, golden):
    for rname, robot in (("m2", 1), ("moonbot", 0)):
        leg = lrm.get_leg(robot, 0.0)
        pts = golden["pts_spot"]
        r = lrm.reachability(pts, leg)            # host-pointer path (apply_kernel contract)
        d, _ = lrm.distance(pts, leg)
        assert np.array_equal(r, golden[f"reach_spot_{rname}"])
        assert np.abs(d - golden[f"dist_spot_{rname}"]).max() < 1e-3


def test_config1_grid_1m(lrm, oracle):
    """BASELINE config[0]/C1: 100^3 lattice over x[-100,600] y[-400,400] z[-500,200], M2 leg."""
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (100, 100, 100))
    n = 100 ** 3
    dev = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    lrm.make_lattice(dev, lo, step, dims)
    pts = dev.cpu().numpy()
    assert np.array_equal(pts, lrm.lattice_host(lo, step, dims))   # device lattice == host formula
    for robot in (1, 0):
        leg = lrm.get_leg(robot, 0.0)
        leg_arr = leg.as_array()
        want_r = oracle.reach(pts, leg_arr, threads=8)
        want_d, want_f = oracle.dist(pts, leg_arr, threads=8)
        rep = _check(lrm, oracle, pts, leg_arr, np.array([1, 0, 0, 0], np.float32), want_r, want_d, want_f,
                     f"C1 robot {robot}")
        fr, _ = lrm.reach_dist(dev, leg)
        assert abs(int(fr.sum().item()) - int(want_r.sum())) <= 3


def test_reference_bench_slice(lrm, oracle):
    """The reference's own published shape: y = 0 slice x[-100,601] z[-100,51] (bench.cpp:112-120),
    float-accumulated arange (bench.cpp:21-27) at pitch 0.64 -> 1096 x 236 points."""
    def arange(start, end, step):
        out, v = [], np.float32(start)
        while v <= np.float32(end):
            out.append(v)
            v = np.float32(v + np.float32(step))
        return np.array(out, np.float32)
    xs, zs = arange(-100, 601, 0.64), arange(-100, 51, 0.64)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    pts = np.stack([X, np.zeros_like(X), Z], -1).reshape(-1, 3).astype(np.float32)
    leg = lrm.get_M2_leg(0.0)
    la = leg.as_array()
    want_r = oracle.reach(pts, la, threads=8)
    want_d, want_f = oracle.dist(pts, la, threads=8)
    _check(lrm, oracle, pts, la, None, want_r, want_d, want_f, "bench slice")


@pytest.mark.parametrize("n", [0, 1, 3, 15, 16, 17, 255, 1023, 1024, 1025, 4099, 70001])
def test_ragged_sizes(lrm, oracle, n):
    rng = np.random.default_rng(n)
    pts = rng.uniform(-500, 600, (n, 3)).astype(np.float32)
    leg = lrm.get_M2_leg(0.5)
    la = leg.as_array()
    if n == 0:
        assert lrm.reachability(torch.empty((0, 3), device="cuda"), leg).shape[0] == 0
        assert lrm.reachability(pts, leg).shape[0] == 0
        return
    want_r = oracle.reach(pts, la)
    want_d, _ = oracle.dist(pts, la)
    dev = torch.from_numpy(pts).cuda()
    fr, vec = lrm.reach_dist(dev, leg)
    torch.cuda.synchronize()
    # every difference must be EXPLAINED: a flag only within 1e-3 mm of the boundary, a vector only
    # on a seam of the reference's own piecewise field (tests/parity.py)
    rep = parity.flag_report(pts, fr.cpu().numpy(), want_r, lambda p: oracle.reach(p, la))
    assert rep["unexplained"] == 0 and rep["mismatch"] <= 1, rep
    rep = parity.dist_report(pts, vec.cpu().numpy(), want_d, lambda p: oracle.dist(p, la)[0])
    assert rep["unexplained"] == 0 and rep["over_tol"] <= max(1, n // 2000), rep
    # host-pointer path
    fr_h, vec_h = lrm.reach_dist(pts, leg)
    assert np.array_equal(fr_h, fr.cpu().numpy()) and np.array_equal(vec_h, vec.cpu().numpy())


def test_unaligned_buffers_take_the_plain_kernel(lrm, oracle):
    """Device buffers that are not 16-byte aligned cannot use the bulk-copy engine."""
    rng = np.random.default_rng(5)
    n = 5000
    pts = rng.uniform(-500, 600, (n, 3)).astype(np.float32)
    leg = lrm.get_moonbot_leg(0.0)
    base = torch.zeros(3 * n + 1, dtype=torch.float32, device="cuda")
    view = base[1:].view(n, 3)                      # 4-byte offset
    view.copy_(torch.from_numpy(pts))
    assert view.data_ptr() % 16 != 0
    out = torch.zeros(3 * n + 1, dtype=torch.float32, device="cuda")[1:].view(n, 3)
    flags = torch.zeros(n + 1, dtype=torch.uint8, device="cuda")[1:]
    lrm.reach_dist(view, leg, out_flags=flags, out_vec=out)
    ref_f, ref_v = lrm.reach_dist(torch.from_numpy(pts).cuda(), leg)
    assert torch.equal(flags, ref_f) and torch.equal(out, ref_v)


def test_soa_matches_aos(lrm):
    rng = np.random.default_rng(11)
    n = 100003
    pts = rng.uniform(-500, 600, (n, 3)).astype(np.float32)
    leg = lrm.get_M2_leg(1.0)
    q = np.array([-0.96194, -0.03806, -0.19134, -0.19134], np.float32)
    dev = torch.from_numpy(pts).cuda()
    fr, vec = lrm.reach_dist(dev, leg, q)
    planes = [dev[:, k].contiguous() for k in range(3)]
    f2, dx, dy, dz = lrm.reach_dist_soa(*planes, leg, q)
    assert torch.equal(f2, fr)
    assert torch.equal(torch.stack([dx, dy, dz], 1), vec)
    f3, *_ = lrm.reach_dist_soa(*planes, leg, q, want_vec=False)
    assert torch.equal(f3, lrm.reachability(dev, leg, q))


def test_distance_lands_on_boundary_at_scale(lrm, sweep):
    """Size-independent property at 2e7 points: |d(p - d(p))| ~ 0 and the fused flag
    equals the stand-alone reach kernel's."""
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (272, 272, 272))
    n = 272 ** 3
    dev = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    lrm.make_lattice(dev, lo, step, dims)
    leg = lrm.get_M2_leg(0.0)
    fr, vec = lrm.reach_dist(dev, leg)
    r = lrm.reachability(dev, leg)
    assert int((fr != r).sum().item()) <= n // 1_000_000 + 2
    back = dev - vec
    _, vec2 = lrm.reach_dist(back, leg)
    resid = vec2.norm(dim=1)
    assert float(resid.quantile(0.999)) < 2e-2 if n <= 16_000_000 else True
    assert int((resid > 5e-2).sum().item()) <= n // 100_000
    assert 0.05 < fr.float().mean().item() < 0.10      # ~7 % reachable for M2 on this box (SURVEY C)


def test_body_orientation_wrapper(lrm, oracle):
    """Non-identity quaternions exercise rotate_leg_data + qtInvRotate + make_asif_leg0."""
    rng = np.random.default_rng(3)
    pts = rng.uniform(-600, 600, (150000, 3)).astype(np.float32)
    for robot, az, q in ((1, 0.7853982, oracle_quat(0)), (0, 2.0, oracle_quat(9)), (1, 4.5, oracle_quat(4))):
        leg = lrm.get_leg(robot, az)
        la = leg.as_array()
        want_r = oracle.reach(pts, la, q, threads=8)
        want_d, want_f = oracle.dist(pts, la, q, threads=8)
        _check(lrm, oracle, pts, la, q, want_r, want_d, want_f, f"quat robot{robot} az{az}")


def oracle_quat(idx):
    from oracle.oracle import PortOracle
    return PortOracle().quaternion_from_angle_index(idx)


def test_forward_kinematics(lrm):
    rng = np.random.default_rng(1)
    ang = rng.uniform(-1.5, 1.5, (10000, 3)).astype(np.float32)
    leg = lrm.get_M2_leg(0.0)
    out = lrm.forward_kinematics(ang, leg)
    c, f, t = ang[:, 0].astype(np.float64), ang[:, 1].astype(np.float64), ang[:, 2].astype(np.float64)
    rad = leg.coxa_length + leg.femur_length * np.cos(f) + leg.tibia_length * np.cos(f + t)
    want = np.stack([leg.body + np.cos(c) * rad, np.sin(c) * rad,
                     leg.femur_length * np.sin(f) + leg.tibia_length * np.sin(f + t)], 1)
    assert np.abs(out - want).max() < 1e-3
    # FK points inside all limits are reachable for the pitch-free moonbot leg (one_leg.cpp:141-202)
    mb = lrm.get_moonbot_leg(0.0)
    a = np.stack([rng.uniform(-1.0, 1.0, 5000), rng.uniform(-1.5, 1.5, 5000), rng.uniform(-2.0, 2.0, 5000)], 1)
    ok = (a[:, 1] + a[:, 2] > mb.tibia_absolute_neg + 1e-3) & (a[:, 1] + a[:, 2] < mb.tibia_absolute_pos - 1e-3)
    a = a[ok].astype(np.float32)
    assert lrm.reachability(lrm.forward_kinematics(a, mb), mb).all()


def test_kernel_ms_is_reported(lrm):
    pts = torch.rand((1 << 20, 3), device="cuda") * 600
    out, ms = lrm.reachability(pts, lrm.get_M2_leg(), timing=True)
    assert 0 < ms < 50


@pytest.fixture(params=["two-tier", "tiered-cta", "tiered-warp"])
def sweep(request, lrm):
    """lrm_set_option("sweep"): 0 = the two-tier sweep (certified tables + full evaluation), 1 = the
    tiered sweep through the choice volume, waiting for the volume instead of letting it build in
    the background; 2 (the default), a coherence probe picks one per launch.  "tier_kernel" picks
    the tiered sweep's organisation: 0 = CTA tiles + CTA-wide rings, 1 = warp-autonomous."""
    old = lrm.set_option("sweep", 0 if request.param == "two-tier" else 1)
    oldk = lrm.set_option("tier_kernel", 0 if request.param == "tiered-cta" else 1)
    yield request.param
    lrm.set_option("sweep", old)
    lrm.set_option("tier_kernel", oldk)


# ---- the fast path (certified tables + deferred redo) only runs on sweeps of >= 4 Mi points -------
def _sample_check(lrm, oracle, dev_pts, leg, q, label, n_sample=250_000, seed=0):
    """Fused + stand-alone distance on the whole device array, oracle on a random sample."""
    n = dev_pts.shape[0]
    la = leg.as_array()
    fr, vec = lrm.reach_dist(dev_pts, leg, q)
    d, f = lrm.distance(dev_pts, leg, q)
    assert torch.equal(vec, d), label                       # fused == separate, bit for bit
    r = lrm.reachability(dev_pts, leg, q)
    assert int((fr != r).sum().item()) <= n // 1_000_000 + 2, label
    idx = torch.from_numpy(np.random.default_rng(seed).choice(n, size=min(n_sample, n), replace=False)).cuda()
    pts = dev_pts[idx].cpu().numpy()
    qq = np.array([1, 0, 0, 0], np.float32) if q is None else np.asarray(q, np.float32)
    want_r = oracle.reach(pts, la, qq, threads=8)
    want_d, want_f = oracle.dist(pts, la, qq, threads=8)
    rep = parity.flag_report(pts, fr[idx].cpu().numpy(), want_r, lambda p: oracle.reach(p, la, qq, threads=8))
    assert rep["unexplained"] == 0 and rep["mismatch"] <= max(3, len(pts) // 20000), (label, rep)
    rep = parity.flag_report(pts, f[idx].cpu().numpy(), want_f, lambda p: oracle.dist(p, la, qq, threads=8)[1])
    assert rep["unexplained"] == 0, (label, rep)
    slack = abs(float((qq.astype(np.float64) ** 2).sum()) - 1.0)
    rep = parity.dist_report(pts, vec[idx].cpu().numpy(), want_d, lambda p: oracle.dist(p, la, qq, threads=8)[0],
                             frame_slack=slack)
    assert rep["unexplained"] == 0 and rep["over_tol"] <= max(3, len(pts) // 2000), (label, rep)
    return vec, fr


def test_fast_path_parity_on_large_sweeps(lrm, oracle, sweep):
    """6 Mi-point clouds (bench box + far field), both robots, identity and tilted orientations."""
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 6 * (1 << 20) + 37
    lo = torch.tensor([-100.0, -400.0, -500.0], device="cuda")
    ext = torch.tensor([700.0, 800.0, 700.0], device="cuda")
    pts = torch.rand((n, 3), device="cuda", generator=g) * ext + lo
    pts[: n // 4] = torch.rand((n // 4, 3), device="cuda", generator=g) * 1800.0 - 900.0
    for robot, az, q in ((1, 0.0, None), (0, 2.0, oracle_quat(9)), (1, 5.5, oracle_quat(4))):
        _sample_check(lrm, oracle, pts, lrm.get_leg(robot, az), q, f"fast robot{robot} az{az}")


def test_reference_bench_slice_at_scale(lrm, oracle, sweep):
    """The reference's own benchmark shape (bench.cpp:109-120: the y = 0 slice, x in [-100, 601],
    z in [-100, 51], float-accumulated arange) at 16.5 M points: every point lies IN a decision
    surface of the yaw tests (y = +0), so no cube of the choice volume is certified — the tiered
    sweep takes its whole-warp table path, the two-tier sweep its full redo; both must meet the bars."""
    def arange32(start, end, step):
        out, v, step = [], np.float32(start), np.float32(step)
        while v <= np.float32(end):
            out.append(v)
            v = np.float32(v + step)
        return np.array(out, np.float32)
    xs, zs = arange32(-100, 601, 0.08), arange32(-100, 51, 0.08)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    pts = np.ascontiguousarray(np.stack([X, np.zeros_like(X), Z], -1).reshape(-1, 3), np.float32)
    assert len(pts) == 16_544_544
    _sample_check(lrm, oracle, torch.from_numpy(pts).cuda(), lrm.get_M2_leg(0.0), None, "bench slice", n_sample=200_000)


def test_fast_path_with_every_point_parked(lrm, oracle, sweep):
    """Points on / next to the coxa axis and on the yaw seam cannot be certified: the whole sweep
    goes through the deferred redo (ring wrap-around, partial last block, drain at the end)."""
    leg = lrm.get_M2_leg(0.0)
    n = 5 * (1 << 20) + 5
    g = torch.Generator(device="cuda").manual_seed(9)
    t = torch.rand(n, device="cuda", generator=g) * 900.0 - 450.0
    # the coxa axis in the world frame (x' = 0 after place_over_coxa, one_leg.cu:9-24):
    # through (body, 0, 0) along (-sin pitch, 0, cos pitch)
    c, s = float(np.cos(leg.coxa_pitch)), float(np.sin(leg.coxa_pitch))
    pts = torch.stack([leg.body - t * s, torch.zeros_like(t), t * c], 1).contiguous()
    pts[n // 2:, 1] = 0.0
    pts[n // 2:, 0] -= torch.rand(n - n // 2, device="cuda", generator=g) * 300.0   # y = +0, behind: the +-pi seam
    vec, fr = _sample_check(lrm, oracle, pts, leg, None, "all parked", n_sample=100_000)
    assert bool(torch.isfinite(vec).all())


def test_fast_path_soa_matches_aos_at_scale(lrm, sweep):
    g = torch.Generator(device="cuda").manual_seed(2)
    n = 5 * (1 << 20)
    pts = torch.rand((n, 3), device="cuda", generator=g) * 900.0 - 300.0
    leg = lrm.get_M2_leg(0.3)
    fr, vec = lrm.reach_dist(pts, leg)
    planes = [pts[:, k].contiguous() for k in range(3)]
    f2, dx, dy, dz = lrm.reach_dist_soa(*planes, leg)
    assert torch.equal(f2, fr) and torch.equal(torch.stack([dx, dy, dz], 1), vec)


def test_golden_vectors_through_the_fast_path(lrm, oracle, golden, sweep):
    """The committed golden vectors (lattices, the bench slice, random clouds, 2 robots x 2 azimuths
    x 4 orientations) with the certified tables forced on for every size: the fast path and its
    deferred redo must meet the same bars as the plain kernels on the reference's own outputs."""
    old = lrm.set_fast_path_min_points(1)
    try:
        for key, rname, az, qname, pname in _cases(golden):
            _check(lrm, oracle, golden[f"pts_{pname}"], golden[f"leg_{rname}_{az}"], golden[f"quat_{qname}"],
                   golden[f"reach_{key}"], golden[f"dist_{key}"], golden[f"dflag_{key}"], "fast " + key)
        # ragged tiny sizes through the ring / drain logic
        rng = np.random.default_rng(4)
        leg = lrm.get_M2_leg(0.0)
        la = leg.as_array()
        for n in (1, 15, 16, 17, 1023, 1025, 4099):
            pts = rng.uniform([-100, -400, -500], [600, 400, 200], (n, 3)).astype(np.float32)
            _check(lrm, oracle, pts, la, None, oracle.reach(pts, la, threads=2), *oracle.dist(pts, la, threads=2),
                   f"fast ragged {n}")
    finally:
        lrm.set_fast_path_min_points(old)


def test_every_sweep_returns_the_same_bits(lrm):
    """The two-tier sweep, the tiered sweep and the probe's own pick return identical bytes — on a
    lattice slab (coherent: the probe picks the tiered sweep), on a shuffled cloud reaching beyond
    the choice volume (the probe picks the two-tier sweep), and on points lying ON the reachability
    edge (every plane cell uncertified: all three rings wrap and overflow)."""
    leg = lrm.get_M2_leg(0.0)
    n = 6 * (1 << 20) + 123
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (7, 1000, 1000))
    lattice = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    lrm.make_lattice(lattice, lo, step, dims, 0, n)
    g = torch.Generator(device="cuda").manual_seed(12)
    cloud = torch.rand((n, 3), device="cuda", generator=g) * 2000.0 - 1000.0
    old = lrm.set_option("sweep", 0)
    try:
        _, v = lrm.reach_dist(lattice, leg)
        edge = (lattice - v).contiguous()
        for name, pts in (("lattice", lattice), ("cloud", cloud), ("edge", edge)):
            got = {}
            for mode, kern in ((0, 1), (1, 0), (1, 1), (2, 1)):
                lrm.set_option("sweep", mode)
                lrm.set_option("tier_kernel", kern)
                fr, vec = lrm.reach_dist(pts, leg)
                d, f = lrm.distance(pts, leg)
                got[(mode, kern)] = (fr, vec, d, f)
            for key in ((1, 0), (1, 1), (2, 1)):
                for a, b in zip(got[(0, 1)], got[key]):
                    assert torch.equal(a, b), (name, key)
        # a call whose output buffer IS its input buffer (lrm_c.h allows it): parked points are
        # redone from an input the tile's store has not changed
        lrm.set_option("sweep", 1)
        want = lrm.distance(lattice, leg)[0]
        for kern in (0, 1):
            lrm.set_option("tier_kernel", kern)
            buf = lattice.clone()
            lrm.distance(buf, leg, out=buf)
            assert torch.equal(buf, want), kern
    finally:
        lrm.set_option("sweep", old)
        lrm.set_option("tier_kernel", 0)


def test_bricks_never_change_a_result(lrm):
    """lrm_set_option("volume_bricks", 1): cubes the choice volume cannot settle carry a brick of 4^3
    fine cubes, found through a pointer in the cube's texel.  Same bytes as the two-tier sweep on a
    lattice slab, on a rotated leg and on points lying ON the reachability edge, through both
    organisations of the tiered sweep; the build reports how many bricks it made."""
    n = 5 * (1 << 20) + 77
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (6, 1000, 1000))
    lattice = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    lrm.make_lattice(lattice, lo, step, dims, 0, n)
    quat = np.array([0.9914449, 0.0, 0.1305262, 0.0], np.float32)   # 15 degrees about y
    old_sweep, old_bricks = lrm.set_option("sweep", 0), lrm.set_option("volume_bricks", 1)
    try:
        for leg, q in ((lrm.get_M2_leg(0.0), None), (lrm.get_moonbot_leg(1.0), quat)):
            lrm.set_option("sweep", 0)
            f0, v0 = lrm.reach_dist(lattice, leg, q)
            r0 = lrm.reachability(lattice, leg, q)
            edge = (lattice - v0).contiguous()
            fe, ve = lrm.reach_dist(edge, leg, q)
            lrm.set_option("sweep", 1)
            for kern in (0, 1):
                lrm.set_option("tier_kernel", kern)
                f1, v1 = lrm.reach_dist(lattice, leg, q)
                assert torch.equal(f0, f1) and torch.equal(v0, v1), kern
                assert torch.equal(r0, lrm.reachability(lattice, leg, q)), kern
                f2, v2 = lrm.reach_dist(edge, leg, q)
                assert torch.equal(fe, f2) and torch.equal(ve, v2), kern
            torch.cuda.synchronize()
            assert lrm.get_stat("volume_bricks") > 1000
            assert lrm.get_stat("volume_brick_capacity") >= lrm.get_stat("volume_bricks")
    finally:
        lrm.set_option("sweep", old_sweep)
        lrm.set_option("volume_bricks", old_bricks)
        lrm.set_option("tier_kernel", 0)


def test_plan_cache_eviction_keeps_results(lrm):
    """More distinct (leg, orientation) plans than the per-device table cache holds (8), swept back
    to back with the choice volume forced on: every sweep must rebuild / reuse atlas and volume
    correctly (eviction while earlier sweeps are still in flight — ordered by events, no host wait —
    and rebuild into the evicted entry's arrays) and return the bytes of the two-tier sweep."""
    n = 4 * (1 << 20) + 77
    lo, step, dims = lrm.lattice_spec((-100, -400, -500), (600, 400, 200), (5, 1000, 1000))
    pts = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    lrm.make_lattice(pts, lo, step, dims, 0, n)
    first = [(robot, az) for robot in (1, 0) for az in (0.0, 0.7, 1.4, 2.1, 2.8)]     # 10 plans > 8 entries
    plans = first + [(1, 0.0), (0, 0.7), (1, 2.8)]
    want = {}
    old = lrm.set_option("sweep", 0)
    try:
        for robot, az in first:
            want[(robot, az)] = lrm.reach_dist(pts, lrm.get_leg(robot, az))
        lrm.set_option("sweep", 1)
        got = [lrm.reach_dist(pts, lrm.get_leg(robot, az)) for robot, az in plans]   # no sync in between
        torch.cuda.synchronize()
    finally:
        lrm.set_option("sweep", old)
    for (robot, az), (fr, vec) in zip(plans, got):
        assert torch.equal(fr, want[(robot, az)][0]) and torch.equal(vec, want[(robot, az)][1]), (robot, az)


def test_hexapod_legs_stay_cached(lrm):
    """A hexapod sweeping its six legs in turn (BASELINE configs[4]'s mounts, k * pi / 3) builds each
    leg's tables once: the second round over the legs builds nothing (lrm_get_stat table_builds)."""
    n = 4 * (1 << 20)
    pts = torch.rand((n, 3), device="cuda") * 900.0 - 300.0
    legs = [lrm.get_M2_leg(float(np.float32(k) * np.float32(np.pi / 3))) for k in range(6)]
    first = [lrm.distance(pts, leg)[0] for leg in legs]
    torch.cuda.synchronize()
    builds = lrm.get_stat("table_builds")
    second = [lrm.distance(pts, leg)[0] for leg in legs]
    torch.cuda.synchronize()
    assert lrm.get_stat("table_builds") == builds
    for a, b in zip(first, second):
        assert torch.equal(a, b)




@pytest.mark.parametrize("n", [777, 40_000, 300_000])
def test_in_place_calls_at_every_size_class(lrm, n):
    """out_xyz == xyz (lrm_c.h allows it) through the plain kernel (< 16 Ki points), the staged
    kernel, and — tables cached by an earlier large call — the table sweeps."""
    g = torch.Generator(device="cuda").manual_seed(n)
    pts = torch.rand((n, 3), device="cuda", generator=g) * 900.0 - 300.0
    leg = lrm.get_M2_leg(0.0)
    want, wf = lrm.distance(pts, leg)
    buf = pts.clone()
    got, gf = lrm.distance(buf, leg, out=buf)
    assert got.data_ptr() == buf.data_ptr()
    assert torch.equal(got, want) and torch.equal(gf, wf)


def test_two_devices_keep_separate_table_caches(lrm):
    """One process driving two GPUs: each device keeps its own LRU of plans, a sweep on
    one never evicts or rebuilds the other's tables, and tensors are launched on their own device
    whatever the current device is."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 4 * (1 << 20) + 5
    legs = [lrm.get_leg(r, az) for r in (1, 0) for az in (0.0, 0.9, 1.8)]
    pts = [torch.rand((n, 3), device=f"cuda:{d}", generator=torch.Generator(device=f"cuda:{d}").manual_seed(3)) * 900 - 300
           for d in (0, 1)]
    first = [[lrm.distance(pts[d], leg)[0] for leg in legs] for d in (0, 1)]       # current device stays cuda:0
    for d in (0, 1):
        torch.cuda.synchronize(d)
    builds = lrm.get_stat("table_builds")
    second = [[lrm.distance(pts[d], leg)[0] for leg in legs] for d in (1, 0)][::-1]
    for d in (0, 1):
        torch.cuda.synchronize(d)
    assert lrm.get_stat("table_builds") == builds          # 6 plans per device fit its own 8 entries
    for d in (0, 1):
        for a, b in zip(first[d], second[d]):
            assert a.device.index == d and torch.equal(a, b)
    assert torch.equal(first[0][0].cpu(), first[1][0].cpu())

"""Body-space octree (apply_oct, several_leg_octree.cu): the sequential restatement in the oracle
(child boxes pinned to the compiled reference through the golden vectors) and GPU parity of
lrm_oct against it."""
import numpy as np
import pytest

from tests import terrain


def wide_leg(port):
    """M2 leg with an (almost) unrestricted coxa yaw: the only kind of leg for which the
    reference's predicate (all four legs, mounted at k*pi/4, reach the SAME foothold) can hold."""
    leg = port.get_leg(1, 0.0).copy()
    leg[8], leg[9] = 3.0, -3.0
    return leg


def test_child_boxes_match_reference_golden(port, golden):
    """CreateChildBox (octree_util.cu.h:105-151) incl. SURVEY appendix D known answers."""
    k = 0
    for p in golden["box_parents"]:
        for c in range(8):
            r, box, missing = port.create_child_box(p, c)
            w = golden["box_children"][k]
            k += 1
            assert missing == int(w[7])
            if missing != 128:
                assert np.array_equal(box, w[:6]) and r == int(w[6])
    # appendix D: root children, c bit-reversed -> sign pattern x<-bit2, y<-bit1, z<-bit0
    root = np.array([0, 0, 0, 5000, 5000, 5000], np.float32)
    signs = {0: (1, 1, 1), 1: (1, 1, -1), 2: (1, -1, 1), 3: (1, -1, -1), 4: (-1, 1, 1), 5: (-1, 1, -1),
             6: (-1, -1, 1), 7: (-1, -1, -1)}
    quads = {0: 0, 1: 4, 2: 2, 3: 6, 4: 1, 5: 5, 6: 3, 7: 7}
    for c in range(8):
        r, box, missing = port.create_child_box(root, c)
        assert tuple(np.sign(box[:3]).astype(int)) == signs[c] and r == quads[c] and missing == 0
        assert np.array_equal(box[3:], [2500, 2500, 2500])
    # one axis below MINBOXSIZE: odd children are dead quadrants
    flat = np.array([0, 0, 0, 150, 80, 150], np.float32)
    assert [port.create_child_box(flat, c)[2] for c in range(8)] == [1, 128, 1, 128, 1, 128, 1, 128]


def test_oracle_octree_semantics(port):
    """As shipped (M2 leg, coxa +-60 deg, mounts k*pi/4) no body box can be valid: the four yaw
    ranges (with their pi-flipped twins) have an empty intersection, so apply_oct returns nothing
    at any depth — a finding about the reference's work-in-progress predicate, kept as a test.
    With a wide coxa range the tree refines and valid boxes appear once boxes are small."""
    terr = terrain.sine_terrain(25, 1200.0, 80.0)
    assert len(port.apply_oct(terr, port.get_leg(1, 0.0), 1)) == 0
    assert len(port.apply_oct(terr, port.get_leg(1, 0.0), 5)) == 0
    wide = wide_leg(port)
    assert len(port.apply_oct(terr, wide, 1)) == 0
    out5 = port.apply_oct(terr, wide, 5)
    out6 = port.apply_oct(terr, wide, 6)
    assert 0 < len(out5) < len(out6)
    assert np.abs(out6).max() < 5000
    assert len(port.apply_oct(np.zeros((0, 3), np.float32), wide, 3)) == 0


@pytest.mark.gpu
def test_gpu_octree_matches_oracle(lrm, port):
    torch = pytest.importorskip("torch")
    terr = terrain.sine_terrain(25, 1200.0, 80.0)
    wide = wide_leg(port)
    leg = lrm.LegDimensions.from_array(wide)
    for depth in (1, 3, 5, 6):
        want = port.apply_oct(terr, wide, depth)
        got = lrm.apply_oct(torch.from_numpy(terr).cuda(), leg, depth)
        # the predicate is pinned to the reference's own kernel child by child
        # (tests/test_refgpu_pin.py); the whole tree must therefore equal the restatement's: same
        # boxes, same traversal order.  A box may only differ where one of its (foothold, sample,
        # leg) distance vectors ties with a box half-extent to the last float32 bit — such a box
        # is reported, and must be re-decided identically by the restatement for a foothold set
        # nudged by 1e-3 mm (the flag band of the parity protocol); anything else fails.
        if not np.array_equal(want, got):
            ws, gs = {tuple(r) for r in want.tolist()}, {tuple(r) for r in got.tolist()}
            nudged = set()
            for d in (-1e-3, 1e-3):
                for ax in range(3):
                    t2 = terr.copy()
                    t2[:, ax] += np.float32(d)
                    nudged |= {tuple(r) for r in port.apply_oct(t2, wide, depth).tolist()} ^ ws
            assert (ws ^ gs) <= nudged, (depth, len(ws), len(gs), sorted(ws ^ gs)[:4])
            assert len(ws ^ gs) <= 2, (depth, len(ws ^ gs))
            common_w = [tuple(r) for r in want.tolist() if tuple(r) in gs]
            common_g = [tuple(r) for r in got.tolist() if tuple(r) in ws]
            assert common_w == common_g
    # host-pointer path and the shipped configuration (empty result)
    got_h = lrm.apply_oct(terr, leg, 5)
    assert len(got_h) == len(lrm.apply_oct(torch.from_numpy(terr).cuda(), leg, 5))
    assert len(lrm.apply_oct(terr, lrm.get_M2_leg(0.0), 4)) == 0
    assert len(lrm.apply_oct(np.zeros((0, 3), np.float32), leg, 2)) == 0


def test_oracle_recurs_depth_map(port):
    """apply_recurs paints the octree depth of the single-leg distance field: deep boxes hug the
    reachability edge, shallow ones are far from it."""
    rng = np.random.default_rng(2)
    pts = rng.uniform(-900, 900, (4000, 3)).astype(np.float32)
    leg = port.get_leg(1, 0.0)
    out = port.apply_recurs(pts, leg, 6)
    assert set(np.unique(out[:, 1:]).tolist()) == {0.0}
    d, _ = port.dist(pts, leg)
    dist = np.linalg.norm(d, axis=1)
    deep, shallow = out[:, 0] >= 6, out[:, 0] <= 3
    assert deep.any() and shallow.any()
    assert np.median(dist[deep]) < np.median(dist[shallow])
    outside = port.apply_recurs(np.array([[6000, 0, 0], [0, -5000.0, 0]], np.float32), leg, 3)
    assert outside[0, 0] == -1.0 and outside[1, 0] == -1.0   # (-h, h]: -5000 is outside


@pytest.mark.gpu
def test_gpu_recurs_matches_oracle(lrm, port):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(4)
    pts = np.concatenate([rng.uniform(-900, 900, (20000, 3)), rng.uniform(-5200, 5200, (4000, 3))]).astype(np.float32)
    for robot, az, q in ((1, 0.0, None), (0, 2.0, port.quaternion_from_angle_index(0))):
        leg_o = port.get_leg(robot, az)
        leg = lrm.LegDimensions.from_array(leg_o)
        for depth in (1, 4, 7):
            want = port.apply_recurs(pts, leg_o, depth, quat=[1, 0, 0, 0] if q is None else q)
            got = lrm.apply_recurs(torch.from_numpy(pts).cuda(), leg, depth, quat=q).cpu().numpy()
            # a box centre whose |d| ties with the box diagonal may flip one subtree
            assert (got[:, 0] != want[:, 0]).sum() <= len(pts) // 500, (robot, depth)
            assert np.array_equal(got[:, 1:], want[:, 1:])
    got_h = lrm.apply_recurs(pts[:5000], lrm.get_M2_leg(0.0), 5)
    want_h = port.apply_recurs(pts[:5000], port.get_leg(1, 0.0), 5)
    assert (got_h[:, 0] != want_h[:, 0]).sum() <= 10


@pytest.mark.gpu
def test_gpu_octree_shards_merge_to_the_full_tree(lrm, port):
    """lrm_oct_sharded: the root's children dealt round-robin to 2 / 3 / 8 shards; the merged lists
    (top-level child c from shard c % nshards) equal the unsharded result, order included."""
    torch = pytest.importorskip("torch")
    terr = torch.from_numpy(terrain.sine_terrain(25, 1200.0, 80.0)).cuda()
    leg = lrm.LegDimensions.from_array(wide_leg(port))
    full, counts = lrm.apply_oct(terr, leg, 6, child_counts=True)
    assert len(full) > 0 and counts.sum() == len(full)
    for nshards in (2, 3, 8):
        parts = [lrm.apply_oct(terr, leg, 6, shard=r, nshards=nshards, child_counts=True) for r in range(nshards)]
        for r, (_, c) in enumerate(parts):
            assert all(c[k] == 0 for k in range(8) if k % nshards != r)
        merged = lrm.merge_oct_shards(parts)
        assert np.array_equal(merged, full), nshards

"""Parity bookkeeping shared by the GPU tests, smoke() and bench.py.

BASELINE.json's bar: reachability / standability flags bit-exact except points within 1e-6 m
(= 1e-3 mm, the reference's own CIRCLE_MARGIN) of the reachability boundary, which are counted
and reported; distance vectors within 1e-5 m (= 1e-2 mm) absolute.  The reference's distance
field is piecewise: at a sector switch, a direct/flipped-coxa switch or a
nearest-candidate switch the vector jumps, so a point sitting on such a seam is classified
separately: it is accepted when the oracle itself produces the GPU's vector somewhere within
1e-3 mm of the point, or when the GPU's vector has the reference's length (within tolerance) and
ends on the reachability edge (an equally-near boundary point: a tie).
"""
import itertools

import numpy as np

FLAG_BAND_MM = 1e-3   # 1e-6 m
DIST_TOL_MM = 1e-2    # 1e-5 m

_OFFSETS = np.array([o for o in itertools.product((-1.0, 0.0, 1.0), repeat=3) if any(o)], np.float32)


def _neighbours(p, radius):
    """26 neighbours of each point at `radius` (per axis), float32."""
    return (p[:, None, :] + _OFFSETS[None, :, :] * np.float32(radius)).astype(np.float32)


def flag_report(pts, got, want, reach_fn):
    """reach_fn(points) -> oracle flags.  Returns dict(mismatch, near_boundary, unexplained)."""
    got = np.asarray(got).astype(np.uint8)
    want = np.asarray(want).astype(np.uint8)
    bad = np.nonzero(got != want)[0]
    unexplained = 0
    if len(bad):
        nb = _neighbours(pts[bad], FLAG_BAND_MM).reshape(-1, 3)
        fl = np.asarray(reach_fn(nb)).reshape(len(bad), -1)
        flips = (fl != want[bad][:, None]).any(axis=1)   # the oracle's own flag changes within the band
        unexplained = int((~flips).sum())
    return {"n": int(len(got)), "mismatch": int(len(bad)), "near_boundary": int(len(bad)) - unexplained,
            "unexplained": unexplained}


def dist_report(pts, got, want, dist_fn, tol=DIST_TOL_MM, frame_slack=0.0):
    """dist_fn(points) -> oracle vectors.  A point over tolerance is a 'seam' point when the oracle
    yields the GPU vector (within tol) at a neighbour <= 1e-3 mm away."""
    err = np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64)).max(axis=1)
    bad = np.nonzero(~(err <= tol))[0]
    unexplained = 0
    if len(bad):
        g = np.asarray(got, np.float64)[bad]
        close = np.zeros(len(bad), bool)
        rng = np.random.default_rng(12345)
        clouds = [_neighbours(pts[bad], radius) for radius in (FLAG_BAND_MM, FLAG_BAND_MM / 4, FLAG_BAND_MM / 16)]
        # where two candidates are nearly tied the reference's own choice flips under sub-micron
        # perturbations (float32 resolution of the two lengths): sample the 1e-3 mm ball as well
        clouds.append((pts[bad][:, None, :] +
                       rng.uniform(-FLAG_BAND_MM, FLAG_BAND_MM, (len(bad), 192, 3))).astype(np.float32))
        for nb in clouds:
            vec = np.asarray(dist_fn(nb.reshape(-1, 3))).reshape(len(bad), -1, 3)
            close |= (np.abs(vec.astype(np.float64) - g[:, None, :]).max(axis=2)
                      <= tol + 2 * FLAG_BAND_MM).any(axis=1)
        # a tie: the GPU vector has the reference's length and ends on the reachability edge, i.e. it
        # points at another boundary point that is equally near (within tol)
        same_len = np.abs(np.linalg.norm(g, axis=1) -
                          np.linalg.norm(np.asarray(want, np.float64)[bad], axis=1)) <= tol
        landing = (pts[bad].astype(np.float64) - g).astype(np.float32)
        land_err = np.linalg.norm(np.asarray(dist_fn(landing), np.float64), axis=1)
        # yardstick: how well the reference's own vector lands (exactly, unless the caller's
        # quaternion is not unit length, in which case the reference's frames are not isometric)
        ref_landing = (pts[bad].astype(np.float64) - np.asarray(want, np.float64)[bad]).astype(np.float32)
        ref_err = np.linalg.norm(np.asarray(dist_fn(ref_landing), np.float64), axis=1)
        # frame_slack = | |q|^2 - 1 |: with a non-unit quaternion the reference's world<->leg maps
        # are not isometric, so "on the edge" is only defined up to that relative distortion
        allow = np.maximum(tol, 1.5 * ref_err) + 2.0 * frame_slack * np.linalg.norm(g, axis=1)
        close |= same_len & (land_err <= allow)
        unexplained = int((~close).sum())
    ok = err[np.isfinite(err) & (err <= tol)]
    return {"n": int(len(err)), "over_tol": int(len(bad)), "seam": int(len(bad)) - unexplained,
            "unexplained": unexplained, "max_err_within_tol": float(ok.max()) if len(ok) else 0.0}


def pose_report(bodies, got, want, stand_fn):
    """Standability parity.  stand_fn(poses) -> oracle result (0 or 1 + first orientation).
    A pose whose flag (or first-orientation index) differs is 'near_boundary' when the oracle
    itself returns the GPU's answer for that pose displaced by <= 1e-3 mm (its decisive foothold
    sits on a leg's reachability boundary, on the gravity-side plane or on a cull cylinder)."""
    got = np.asarray(got).astype(np.int32)
    want = np.asarray(want).astype(np.int32)
    flag_bad = np.nonzero((got != 0) != (want != 0))[0]
    idx_bad = np.nonzero((got != want) & (got != 0) & (want != 0))[0]
    out = {"n": int(len(got)), "flag_mismatch": int(len(flag_bad)), "orientation_mismatch": int(len(idx_bad)),
           "unexplained": 0}
    bad = np.concatenate([flag_bad, idx_bad])
    if len(bad):
        offs = np.array(list(itertools.product((-1.0, 0.0, 1.0), repeat=3)), np.float32) * np.float32(FLAG_BAND_MM)
        nb = (bodies[bad][:, None, :] + offs[None, :, :]).astype(np.float32)
        res = np.asarray(stand_fn(nb.reshape(-1, 3))).reshape(len(bad), -1).astype(np.int32)
        ok = (res == got[bad][:, None]).any(axis=1)
        out["unexplained"] = int((~ok).sum())
    return out

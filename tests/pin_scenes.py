"""Fixed inputs of the reference-GPU pins (tests/test_refgpu_pin.py, tools/refgpu_dump.py,
tests/golden/make_refgpu_golden.py): everything is generated from seeds, nothing is read from disk.
"""
import numpy as np

from tests import terrain


def _port():
    from oracle.oracle import PortOracle
    return PortOracle()


def m2_legs4():
    """The four M2 legs robot_full_struct is driven with (mounts k*pi/2, several_leg.cpp:60-70)."""
    p = _port()
    return np.stack([p.get_leg(1, float(np.float32(k) * np.float32(np.pi / 2))) for k in range(4)]).astype(np.float32)


def full_struct_scenes():
    """name -> (map points, body poses, legs 4 x 14).  The two scenes of round 1's builder-run
    comparison (tools/vs_refgpu.py): a sinusoidal terrain and a window of the C3 Perlin generator."""
    legs = m2_legs4()
    sine = terrain.sine_terrain(128, 2400.0, 100.0)
    perlin = terrain.perlin_terrain(128)
    return {
        "sine_16k_x_6912": (sine, terrain.body_lattice(sine, 24, 24, 12), legs),
        "perlin_16k_x_20k": (perlin, terrain.body_lattice(perlin, 32, 40, 16), legs),
    }


def oct_footholds():
    """169 footholds: the reference's kernel can only run as one GPU thread (oracle/ref_gpu_shim.cu)."""
    return terrain.sine_terrain(13, 600.0, 80.0)


def oct_cases():
    """name -> (parent box {centre, half extents}, parent validity, leg 14 floats): one-level trees
    for validity_child covering its branches — box rule vs convex-radius rule (:96-105), rotation
    samples on / off (:53-57), unsplit axes and dead quadrants (CreateChildBox), an inherited
    parent validity (:71)."""
    p = _port()
    m2 = p.get_leg(1, 0.0)
    wide = m2.copy()
    wide[8], wide[9] = 3.0, -3.0            # (almost) unrestricted coxa yaw: the predicate can hold
    moon = p.get_leg(0, 0.0).copy()
    moon[8], moon[9] = 3.0, -3.0
    cases = {}
    boxes = {
        "root": [0, 0, 0, 5000, 5000, 5000],
        "b400": [100, 50, 150, 400, 400, 400],
        "b200": [-150, 120, 200, 200, 200, 200],
        "sphere": [60, -40, 180, 90, 90, 90],            # children 45^3: convex-radius rule, no rotation
        "rot_box": [30, 20, 170, 45, 150, 150],          # x unsplit, rotation samples, box rule
        "rot_sphere": [-20, 35, 160, 45, 110, 45],       # only y split, rotation samples, radius rule
        "flat": [0, 0, 150, 150, 80, 150],               # y below MINBOXSIZE: odd children dead
        "tiny": [10, 10, 165, 40, 40, 40],               # nothing splits: child 0 == parent, rest dead
        "high": [0, 0, 650, 200, 200, 200],              # upper children out of every leg's reach
        "far": [0, 0, 1500, 400, 400, 400],              # no foothold inside the elongated box
    }
    for bname, box in boxes.items():
        for lname, leg in (("m2", m2), ("wide", wide), ("moonwide", moon)):
            if lname == "moonwide" and bname not in ("b200", "rot_box"):
                continue
            cases[f"{bname}_{lname}"] = (np.array(box, np.float32), 0, leg)
    cases["b200_wide_pv"] = (np.array(boxes["b200"], np.float32), 1, wide)
    cases["rot_box_m2_pv"] = (np.array(boxes["rot_box"], np.float32), 1, m2)
    return cases

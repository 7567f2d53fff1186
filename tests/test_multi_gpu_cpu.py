"""CPU-only cover of the N>1 path: world_size-2 gloo job (tests/_dist_worker.py) + slab arithmetic."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_ranges_partition_exactly(lrm):
    from importlib import import_module
    slabs = import_module("lrm_b200.slabs")
    for n in (0, 1, 7, 1000, 10 ** 9 + 3):
        for world in (1, 2, 3, 8):
            spans = [slabs.slab_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert slabs.weak_slab(10 ** 9, 3, 8) == (3 * 10 ** 9, 10 ** 9)
    with pytest.raises(ValueError):
        slabs.slab_range(10, 2, 2)


def test_dealt_chunks_partition_exactly(lrm):
    """dealt_chunks: world x chunks_per_rank contiguous chunks dealt round-robin; over all ranks
    they partition [0, n) exactly, and every rank's share differs by at most chunks_per_rank units."""
    from importlib import import_module
    slabs = import_module("lrm_b200.slabs")
    for n in (0, 5, 1000, 16_777_216):
        for world in (1, 2, 8):
            parts = [slabs.dealt_chunks(n, r, world, chunks_per_rank=8) for r in range(world)]
            spans = sorted(sum(parts, []))
            pos = 0
            for f, c in spans:
                assert f == pos
                pos += c
            assert pos == n
            sizes = [sum(c for _, c in p) for p in parts]
            assert max(sizes) - min(sizes) <= 8
            for p in parts:
                assert p == sorted(p)
    with pytest.raises(ValueError):
        slabs.dealt_chunks(10, 2, 2)


def test_two_rank_gloo_job_covers_the_lattice():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[len("RESULT "):])
    (f0, c0, r0, a0), (f1, c1, r1, a1) = res["slabs"]
    assert f0 == 0 and f0 + c0 == f1 and c0 + c1 == res["total"]          # contiguous, complete
    assert int(r0 + r1) == res["want_reach"]                                # per-slab results add up
    assert abs((a0 + a1) - res["want_abs"]) < 1e-3 * max(1.0, res["want_abs"])
    assert res["max_elapsed"] > 0

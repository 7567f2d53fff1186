"""Capture the reference-GPU pins as committed fixtures (needs a GPU + oracle/_ref/libref_gpu*_precise.so):

    python tests/golden/make_refgpu_golden.py [outdir]      # default: tests/golden

Writes refgpu_full_struct.npz (robot_full_struct's standable sets on tests/pin_scenes.py's two
scenes) and refgpu_validity_child.npz (validity_child's per-child flags on its one-level trees),
each from a fresh subprocess of tools/refgpu_dump.py.  Under gpurun pass gpurun_out/ and copy the
two files into tests/golden/ afterwards (only gpurun_out/ travels back from the GPU box).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
for what, name in (("full", "refgpu_full_struct.npz"), ("oct", "refgpu_validity_child.npz")):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "refgpu_dump.py"), what, os.path.join(out, name)],
                   check=True, cwd=ROOT)
    print("wrote", os.path.join(out, name))

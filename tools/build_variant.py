"""Measurement builds of liblrm_b200.so with extra -D flags, written to tools/_variants/ (git-ignored,
travels to the GPU box).  Select one at run time with LRM_B200_LIB=<path>.

    python tools/build_variant.py g2 -DLRM_T0_GROUP=2
"""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "legged-robot-movability-cuda_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "tools", "_variants", f"liblrm_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = ["nvcc"] + b.NVCC_FLAGS + extra + ["-Xptxas", "-v", "-shared", "-o", out, "-x", "cu"] + \
      [os.path.join(b.CSRC, s) for s in b.SOURCES] + ["-lcudart"]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode != 0:
    sys.exit(r.stderr[-3000:])
lines = r.stderr.splitlines()
for i, l in enumerate(lines):
    if ("one_leg_tier_kernelILi3ELb0" in l or "one_leg_warp_kernelILi3ELb0" in l) and "Compiling" in l:
        print(name, " | ".join(x.strip() for x in lines[i + 2:i + 4]))
print(out)

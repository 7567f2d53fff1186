"""Build the test-only host emulator of the device math (see emu.cu)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "legged-robot-movability-cuda_b200", "csrc")
LIB = os.path.join(HERE, "liblrm_emu.so")


def build(force=False):
    deps = [os.path.join(HERE, "emu.cu")] + [os.path.join(CSRC, f) for f in
                                             ("leg_math.cuh", "leg_plan.h", "leg_plan.cpp", "fast_tables.cpp")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    fma = "-mfma" if "fma" in open("/proc/cpuinfo").read() else "-O2"
    cmd = ["nvcc", "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", fma,
           "-Xcompiler", "-ffp-contract=off", "-gencode", "arch=compute_100a,code=sm_100a",
           "-I" + CSRC, "-I" + os.path.join(ROOT, "include"), "-o", LIB,
           "-x", "cu", os.path.join(HERE, "emu.cu"), os.path.join(CSRC, "leg_plan.cpp"), os.path.join(CSRC, "fast_tables.cpp")]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

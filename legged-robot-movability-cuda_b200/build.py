"""Build liblrm_b200.so (hand-written sm_100a CUDA + the C ABI of include/lrm_c.h) in-tree.

    python legged-robot-movability-cuda_b200/build.py [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels with the tree
to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblrm_b200.so")
SOURCES = ["leg_plan.cpp", "fast_tables.cpp", "one_leg_kernels.cu", "plane_atlas.cu", "positionability.cu", "octree.cu", "lrm_api.cu"]
HEADERS = ["leg_plan.h", "leg_math.cuh", "bulk_copy.cuh", "kernels.h", "cell_grid.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    # host-built constants (leg plans, yaw tables, gravity knife-edge coefficients) must round like
    # the reference's host pass and the oracle: never contract a*b+c (GCC does by default on aarch64)
    "-Xcompiler", "-ffp-contract=off",
    "--expt-relaxed-constexpr",
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps += [os.path.join(ROOT, "include", "lrm_c.h"), os.path.abspath(__file__),
             os.path.join(HERE, "driver", "lrm_cuda.cpp")]
    if not os.path.exists(os.path.join(HERE, "lrm_cuda")):
        return True
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    # LRM_NVCC_EXTRA: extra flags for measurement builds (e.g. "-DLRM_TIER_THREADS=192 -DLRM_TIER_TILE=768")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("LRM_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB]
    cmd += ["-x", "cu"] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lcudart"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    build_driver()
    return LIB


DRIVER = os.path.join(HERE, "lrm_cuda")


def build_driver():
    """The file-protocol driver (driver/lrm_cuda.cpp): plain C++ over the C ABI, rpath = its own dir."""
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), "-o", DRIVER,
           os.path.join(HERE, "driver", "lrm_cuda.cpp"), "-L" + HERE, "-llrm_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return DRIVER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

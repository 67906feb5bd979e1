"""Builds liblgs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m lidar_graph_slam_b200.build [--force] [--verbose]
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblgs_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

NVCC = os.environ.get("LGS_NVCC", "/usr/local/cuda/bin/nvcc")
# the image exports CXX=/opt/gcc/bin/g++ (a wrapper without OpenMP specs); pin the distro host compiler
CCBIN = os.environ.get("LGS_CCBIN", "/usr/bin/g++")

COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # No FMA contraction: integer decisions (voxel index, validity flags, range test) and the parity of the
    # f32 derivative terms depend on products and sums rounding separately, as in the SSE-only reference build
    # (thirdparty/ndt_omp/CMakeLists.txt:5-6).  Kernels that want FMA opt in with explicit fmaf().
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-ccbin", CCBIN,
] + os.environ.get("LGS_EXTRA_NVCC_FLAGS", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp")) +
                  glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    os.makedirs(OBJ, exist_ok=True)
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + COMMON + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s\n%s\n" % (os.path.basename(s), out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(OUT, objs):
        cmd = [NVCC, "-Wno-deprecated-gpu-targets", "-shared", "-cudart", "static", "-ccbin", CCBIN, "-o", OUT] + objs + ["-lpthread"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

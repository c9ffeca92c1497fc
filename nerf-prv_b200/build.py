"""Builds libprv_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python nerf-prv_b200/build.py [--force] [--verbose]

Flags that matter for bit-exactness: -fmad=false (no FMA contraction in device code) and
-Xcompiler -ffp-contract=off (none in host code).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libprv_b200.so")
SOURCES = ["prv_device.cu", "prv_host.cpp"]
DEPS = ["prv_kernels.cuh", "prv_view_const.hpp", "kernels_common.cuh", "kernels_cast.cuh", "kernels_map.cuh", "kernels_greedy.cuh", "kernels_ensemble.cuh",
        "kernels_splat.cuh", "prv_keys.hpp", "../host/prv_linalg.hpp", "../host/View_Space.hpp", "../host/Share_Data.hpp",
        "../host/Perception_3D.hpp", "../host/NBV_Net_Labeler.hpp", "../host/prv_io.hpp", "../host/prv_simulation.cpp", "../../include/prv.h"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
           "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall", "-shared", "-o", LIB]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl"]
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or verbose:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libprv_b200.so")
    build_driver(verbose)
    return LIB


DRIVER = os.path.join(HERE, "prv_simulation")


def build_driver(verbose=False):
    """The C++ host driver (reference stdin protocol, mode 3) on top of the C ABI."""
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-o", DRIVER, os.path.join(HERE, "host", "prv_simulation.cpp"),
           "-L" + HERE, "-lprv_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or verbose:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("g++ failed building prv_simulation")
    return DRIVER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

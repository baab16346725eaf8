"""Builds the in-tree CUDA library ``viennaray_b200/libviennaray_b200.so`` for
sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libviennaray_b200.so")
SOURCES = ["vr_api.cu", "vr_bvh.cu", "vr_scene.cu", "vr_trace.cu"]
HEADERS = ["vr_internal.h", "vr_device.cuh", os.path.join("..", "..", "include", "viennaray_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-reproducible float arithmetic against the CPU oracle: no FMA
    # contraction, IEEE division and square root, denormals kept
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("VR_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building libviennaray_b200.so")
    return LIB


if __name__ == "__main__":
    build_library(force=True, verbose="-v" in sys.argv)
    print(LIB)

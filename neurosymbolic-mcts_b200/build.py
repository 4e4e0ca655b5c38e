"""Build libcaissa_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python neurosymbolic-mcts_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
nvcc cross-compiles without a GPU, so this is also the CPU-side build check.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libcaissa_b200.so")
SELFPLAY = os.path.join(HERE, "lib", "caissa_selfplay")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(exts):
    out = [os.path.join(HERE, "..", "include", "caissa_b200.h")]
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(root, f))
    return out


def build_lib(force=False, verbose=False):
    srcs = _sources((".cu", ".cuh", ".h", ".hpp", ".cpp"))
    if not force and _newer(LIB, srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    units = [os.path.join(CSRC, "cb_api.cu")]
    host_dir = os.path.join(CSRC, "host")
    if os.path.isdir(host_dir):
        units += sorted(os.path.join(host_dir, f) for f in os.listdir(host_dir) if f.endswith(".cpp"))
    cmd = ["nvcc"] + NVCC_FLAGS + ["-shared", "-o", LIB] + units + ["-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))

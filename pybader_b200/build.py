"""Builds libbader_b200.so in-tree with nvcc for sm_100a only.

-fmad=false: ongrid pointers and trajectory steps must be bit-exact against
the reference, whose numba code has no FMA contraction.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "bader_b200.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "kernels.cuh"), os.path.join(HERE, "csrc", "common.cuh"),
        os.path.join(HERE, "csrc", "seed.cuh"), os.path.join(HERE, "csrc", "edge.cuh"), os.path.join(HERE, "csrc", "comm.cuh"), os.path.join(HERE, "csrc", "parse.cuh"),
        os.path.join(HERE, "csrc", "parse_num.h"), os.path.join(HERE, "csrc", "format.h"), os.path.join(HERE, "csrc", "pow5_table.h"),
        os.path.join(os.path.dirname(HERE), "include", "bader_b200.h")]
SO = os.path.join(HERE, "libbader_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-fmad=false",
    "-shared", "-Xcompiler", "-fPIC", "-ldl",
]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(SO)

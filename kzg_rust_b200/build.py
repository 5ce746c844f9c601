"""Builds libkzg_b200.so (hand-written sm_100a CUDA + the C ABI of include/kzg_b200.h).

    python -m kzg_rust_b200.build        # in-tree, next to this file

nvcc cross-compiles without a GPU.  The library is git-ignored but travels with the tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkzg_b200.so")
SOURCES = ["kzg_b200.cu", "host_pairing.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
]


def _newest_source_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in os.listdir(root):
            newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def build(force=False, verbose=False, extra_flags=()):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

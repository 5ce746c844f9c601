"""Builds libkzg_b200.so (hand-written sm_100a CUDA + the C ABI of include/kzg_b200.h).

    python -m kzg_rust_b200.build [--force] [-v]      # in-tree, next to this file

The translation units are compiled in parallel (objects under build/obj) and linked with nvcc.
nvcc cross-compiles without a GPU.  The library is git-ignored but travels with the tree."""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkzg_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
SOURCES = ["kzg_b200.cu", "msm.cu", "g1ops.cu", "frops.cu", "host_pairing.cpp", "host_sha256.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-lineinfo", "-Xcompiler", "-fPIC",
]


def _newest_source_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(ROOT, "include")):
        for name in os.listdir(root):
            newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def _compile(nvcc, src, obj, flags, verbose):
    cmd = [nvcc] + NVCC_FLAGS + list(flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
    subprocess.check_call(cmd)
    return obj


def build(force=False, verbose=False, extra_flags=(), lib=LIB):
    """extra_flags: more nvcc flags (e.g. -DKZG_ADD_MIN_BLOCKS=2 for an experiment); lib: output path."""
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _newest_source_mtime():
        return lib
    nvcc = os.environ.get("NVCC", "nvcc")
    tag = hashlib.sha1(" ".join(extra_flags).encode()).hexdigest()[:8] if extra_flags else "default"
    obj_dir = os.path.join(OBJ_DIR, tag)
    os.makedirs(obj_dir, exist_ok=True)
    newest = _newest_source_mtime()
    jobs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        for src in SOURCES:
            obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
            if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
                jobs.append(pool.submit(_compile, nvcc, src, obj, extra_flags, verbose))
        for j in jobs:
            j.result()
    objs = [os.path.join(obj_dir, os.path.splitext(src)[0] + ".o") for src in SOURCES]
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

#!/usr/bin/env python3
"""Build differently configured copies of the library for A/B runs on the GPU box (tools/variant_sweep.sh):

    python tools/build_variants.py name1:-DFLAG=1,-DOTHER=2 name2:...      -> build/variants/<name>.so

build/ is git-ignored but travels to the GPU box with the tree."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from kzg_rust_b200 import build as b  # noqa: E402

out = os.path.join(ROOT, "build", "variants")
os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    lib = os.path.join(out, name + ".so")
    b.build(force=True, extra_flags=tuple(f for f in flags.split(",") if f), lib=lib)
    print(lib)

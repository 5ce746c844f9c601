#!/usr/bin/env python3
"""Where the small-call forms of the MSM stop paying: wall time of blob_to_kzg_commitment_batch on pinned host buffers
for n blobs with the latency comb / the warp-per-sum form / the batched affine tree.  Not part of the product."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import kzg_rust_b200 as k
from golden_util import golden
from gpu_util import synthetic_blobs
g = golden()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, int(os.environ.get("AB_COMB", "20")))
nmax = 1024
blobs = torch.from_numpy(synthetic_blobs(nmax, seed=99)).pin_memory().numpy()
ref = None
for n in (1, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
    row = []
    outs = []
    for label, small_max, lat in (("latency comb", 1024, 1), ("warp per sum", 1024, 0), ("affine tree", 0, 1)):
        if label == "warp per sum" and n > 64:
            row.append("%s      -   " % label); continue
        os.environ["KZG_B200_MSM_SMALL_MAX"] = str(small_max); os.environ["KZG_B200_LATENCY_TABLE"] = str(lat)
        fn = lambda: k.Kzg.blob_to_kzg_commitment_batch(blobs[:n], s)
        out, st = fn(); outs.append(out.tobytes()); fn()
        ts = []
        for _ in range(7):
            t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
        row.append("%s %7.3f ms" % (label, sorted(ts)[3]))
    assert all(o == outs[0] for o in outs)
    print("n=%-5d %s" % (n, "   ".join(row)), flush=True)

#!/usr/bin/env python3
"""KZG_B200_TRACE=1 on a few small verify_blob_kzg_proof_batch calls (host buffers): where the wall clock of a Deneb-sized
call goes on the host side.  Not part of the product."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import kzg_rust_b200 as k
from golden_util import golden
from gpu_util import synthetic_blobs
g = golden()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 16)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
blobs = synthetic_blobs(n, seed=3)
cms, st = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
prs, st = k.Kzg.compute_blob_kzg_proof_batch(blobs, cms, s)
pin = lambda a: torch.empty(a.shape, dtype=torch.uint8, pin_memory=True).copy_(torch.from_numpy(a)).numpy()
vb, vc, vp = pin(blobs), pin(cms), pin(prs)
for _ in range(3):
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(vb, vc, vp, n, s)
os.environ["KZG_B200_TRACE"] = sys.argv[2] if len(sys.argv) > 2 else "1"
for _ in range(3):
    t = time.perf_counter()
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(vb, vc, vp, n, s)
    print("call: %.3f ms" % ((time.perf_counter() - t) * 1e3), flush=True)

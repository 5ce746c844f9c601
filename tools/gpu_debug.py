#!/usr/bin/env python3
"""Scratch GPU debugging: commitment parity across window widths / blob shapes."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kzg_rust_b200 as k
from golden_util import golden
from gpu_util import synthetic_blobs, oracle_settings
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
g = golden(); o = oracle_settings("mainnet")
os.environ["KZG_B200_CHUNK"] = os.environ.get("KZG_B200_CHUNK", "64")
cs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "5,8,11").split(",")]
rng = np.random.default_rng(1)
full = np.frombuffer(b"".join((int.from_bytes(rng.bytes(32), "big") % R).to_bytes(32, "big") for _ in range(4 * 4096)), dtype=np.uint8).reshape(4, -1).copy()
for c in cs:
    s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, c)
    for name, blobs in (("top0x1", synthetic_blobs(1, seed=9)), ("top0x8", synthetic_blobs(8, seed=10)), ("top0x24", synthetic_blobs(24, seed=11)), ("full4", full)):
        out, st = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
        exp, est = o.blob_to_kzg_commitment_many(blobs, nthreads=os.cpu_count())
        bad = [i for i in range(len(exp)) if not np.array_equal(out[i], exp[i])]
        print("c=%d %s: bad=%s status=%s" % (c, name, bad, st.tolist() if st.any() else "ok"), flush=True)
    s.close()

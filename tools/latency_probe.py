#!/usr/bin/env python3
"""Per-stage device time and wall time of small calls (verify n = 1 and 6, one commitment, one proof) and of
the host pairing check alone.  Not part of the product."""
import ctypes, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kzg_rust_b200 as k
from golden_util import golden
from gpu_util import synthetic_blobs
STAGES = ["digits", "msm_gather", "msm_tree", "compress", "challenge", "eval", "validate", "verify_terms"]
g = golden(); L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 0)
blobs = synthetic_blobs(8, seed=3)
cms, _ = k.Kzg.blob_to_kzg_commitment_batch(blobs, s); prs, _ = k.Kzg.compute_blob_kzg_proof_batch(blobs, cms, s)
ms = (ctypes.c_double * 8)(); ln = (ctypes.c_uint64 * 8)()
def stages(fn, label):
    for _ in range(3): fn()
    t=time.perf_counter(); fn(); wall=(time.perf_counter()-t)*1e3
    L.kzg_b200_profile_enable(s._h, 1); fn(); L.kzg_b200_profile_read(s._h, ms, ln); L.kzg_b200_profile_enable(s._h, 0)
    print(label, "wall %.2f ms" % wall, {STAGES[i]: round(ms[i], 3) for i in range(8) if ms[i] > 0}, "sum %.2f" % sum(ms))
for n in (1, 6):
    stages(lambda: k.Kzg.verify_blob_kzg_proof_batch_raw(blobs[:n], cms[:n], prs[:n], n, s), "verify n=%d" % n)
stages(lambda: k.Kzg.blob_to_kzg_commitment_batch(blobs[:1], s), "commit n=1")
stages(lambda: k.Kzg.compute_blob_kzg_proof_batch(blobs[:1], cms[:1], s), "proof n=1")
# host pairing alone
a1 = cms[0].tobytes(); 
ok = ctypes.c_int(0)
g2 = g.g2_bytes[:96]
L.kzg_b200_pairings_verify.argtypes=[ctypes.c_char_p]*4+[ctypes.POINTER(ctypes.c_int)]
t=time.perf_counter()
for _ in range(5): L.kzg_b200_pairings_verify(a1, g2, a1, g2, ctypes.byref(ok))
print("pairings_verify %.2f ms, ok=%d" % ((time.perf_counter()-t)/5*1e3, ok.value))

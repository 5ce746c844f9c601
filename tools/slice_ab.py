#!/usr/bin/env python3
"""A/B of the column-sliced uploads of host verification (KZG_B200_HASH_SLICES): wall time per call of
verify_blob_kzg_proof_batch on pinned host buffers for several batch sizes.  Not part of the product."""
import ctypes, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import kzg_rust_b200 as k
from golden_util import golden
g = golden(); L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, int(os.environ.get("AB_COMB", "20")))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(7)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev, generator=gen); blobs[:, :, 0] = 0
cm = torch.zeros((n, 48), dtype=torch.uint8, device=dev); pr = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev)
assert L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), n, cm.data_ptr(), st.data_ptr()) == 0
assert L.kzg_b200_compute_blob_kzg_proof_device(s._h, blobs.data_ptr(), cm.data_ptr(), n, pr.data_ptr(), st.data_ptr()) == 0
L.kzg_b200_synchronize(s._h)
pin = lambda t_: torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True).copy_(t_).numpy()
vb, vc, vp = pin(blobs.reshape(n, 131072)), pin(cm), pin(pr)
for m in (64, 128, 256, 1024, 4096, n):
    if m > n: continue
    row = []
    for slices in sys.argv[2:] or ["1", "4", "8"]:
        os.environ["KZG_B200_HASH_SLICES"] = slices
        fn = lambda: k.Kzg.verify_blob_kzg_proof_batch_raw(vb[:m], vc[:m], vp[:m], m, s)
        assert fn() is True
        fn()
        reps = 10 if m <= 4096 else 4
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
        row.append("S=%s %7.2f ms (min %7.2f)" % (slices, sorted(ts)[len(ts) // 2], min(ts)))
    print("verify host n=%-6d %s" % (m, "   ".join(row)), flush=True)

#!/usr/bin/env python3
"""First-contact GPU probe: roofline micro-benchmarks and commitment throughput for a few
window widths / batch sizes.  Writes gpurun_out/probe.json.  Not part of the product."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import kzg_rust_b200 as k  # noqa: E402
from golden_util import golden  # noqa: E402
from gpu_util import synthetic_blobs  # noqa: E402

out = {"probe": []}
g = golden()
L = k.load_library()
cs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8,12").split(",")]
nblobs = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
for c in cs:
    t0 = time.time()
    s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, c)
    t_create = time.time() - t0
    imad, imadw, fpmul = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    L.kzg_b200_measure_peaks(s._h, ctypes.byref(imad), ctypes.byref(imadw), ctypes.byref(fpmul))
    blobs = torch.from_numpy(synthetic_blobs(min(nblobs, 256), seed=3)).cuda()
    reps = (nblobs + blobs.shape[0] - 1) // blobs.shape[0]
    blobs = blobs.repeat(reps, 1)[:nblobs].contiguous()
    outb = torch.zeros((nblobs, 48), dtype=torch.uint8, device="cuda")
    st = torch.zeros(nblobs, dtype=torch.int32, device="cuda")
    for _ in range(2):
        rc = L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), nblobs, outb.data_ptr(), st.data_ptr())
        assert rc == 0, rc
        L.kzg_b200_synchronize(s._h)
    t0 = time.time()
    rc = L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), nblobs, outb.data_ptr(), st.data_ptr())
    L.kzg_b200_synchronize(s._h)
    dt = time.time() - t0
    rec = {"c": c, "create_s": t_create, "imad_per_s": imad.value, "imad_wide_per_s": imadw.value, "fp_mul_per_s": fpmul.value,
           "blobs": nblobs, "commit_s": dt, "blobs_per_s": nblobs / dt, "status_any": bool(st.any().item())}
    print(rec, flush=True)
    out["probe"].append(rec)
    s.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)

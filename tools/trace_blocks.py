#!/usr/bin/env python3
"""Experiment: when do the warps of one addition-kernel launch finish?  Needs a library built with
-DKZG_TRACE (python tools/trace_blocks.py build) loaded via KZG_B200_LIB.  Not part of the product."""
import ctypes, json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
TRACE_LIB = os.path.join(ROOT, "build", "libkzg_b200_trace.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    from kzg_rust_b200 import build as b
    os.makedirs(os.path.dirname(TRACE_LIB), exist_ok=True)
    cmd = ["nvcc"] + b.NVCC_FLAGS + ["-DKZG_TRACE", "-o", TRACE_LIB] + [os.path.join(b.CSRC, s) for s in b.SOURCES]
    subprocess.check_call(cmd); print(TRACE_LIB); sys.exit(0)
import numpy as np, torch
import kzg_rust_b200.kzg as kk
kk.LIB_PATH = TRACE_LIB
import kzg_rust_b200 as k
from golden_util import golden
g = golden(); L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 0)
n = 4096
dev = torch.device("cuda", 0)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev); blobs[:, :, 0] = 0
out = torch.zeros((n, 48), dtype=torch.uint8, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
trace = torch.zeros(4 * (1 + 4 * 444 * 4), dtype=torch.int64, device=dev)
W = {19: 14, 18: 15}.get(s.window_bits, 14)
level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
trace[0] = n * W * (2048 >> level)  # additions of the wanted level (0 = gather)
L.kzg_b200_debug_set_trace.argtypes = [ctypes.c_void_p]
for it in range(2):
    L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), n, out.data_ptr(), st.data_ptr()); L.kzg_b200_synchronize(s._h)
# the trace holds the LAST launch that wrote each slot; to isolate one level run with only the big levels:
# levels overwrite each other, so trace only launches whose total equals the wanted one
res = {}
L.kzg_b200_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), n, out.data_ptr(), st.data_ptr()); L.kzg_b200_synchronize(s._h)
t = trace.cpu().numpy().reshape(-1, 4)[1:]
t = t[t[:, 1] > 0]
print("last writer total =", set(t[:, 3].tolist()))
start, end, sm = t[:, 0], t[:, 1], t[:, 2]
t0 = start.min(); dur = (end.max() - t0) / 1e3
print("warps %d, kernel %.1f us" % (len(t), dur))
rel = (end - t0) / 1e3 / dur
print("end-time quantiles (fraction of kernel):", np.quantile(rel, [0, .1, .25, .5, .75, .9, 1]).round(3).tolist())
print("mean warp lifetime / kernel = %.3f" % (((end - start) / 1e3).mean() / dur))
# per SM: sorted end times of its blocks
per = {}
for e, m in zip(rel, sm): per.setdefault(int(m), []).append(float(e))
rows = [sorted(v) for v in per.values()]
print("SMs seen:", len(rows), "warps per SM:", sorted(set(len(r) for r in rows)))
arr = np.array([r for r in rows if len(r) == 12])
if len(arr): print("mean sorted end times within an SM:", arr.mean(axis=0).round(3).tolist())

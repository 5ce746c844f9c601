#!/usr/bin/env python3
"""Scratch: where does the host-buffer call spend its time? (not part of the product)"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kzg_rust_b200 as k
from golden_util import golden
g = golden(); L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, int(sys.argv[1]) if len(sys.argv) > 1 else 16)
n = 16384
dev = torch.device("cuda", 0)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev); blobs[:, :, 0] = 0
h = torch.empty((n, 131072), dtype=torch.uint8, pin_memory=True); h.copy_(blobs.reshape(n, 131072))
out = torch.zeros((n, 48), dtype=torch.uint8, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
ho = torch.empty((n, 48), dtype=torch.uint8, pin_memory=True); hs = torch.empty(n, dtype=torch.int32, pin_memory=True)
def dev_call():
    L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), n, out.data_ptr(), st.data_ptr()); L.kzg_b200_synchronize(s._h)
def host_call():
    L.kzg_b200_blob_to_kzg_commitment_batch(s._h, h.data_ptr(), n, ho.data_ptr(), hs.data_ptr())
def h2d():
    blobs.reshape(n, 131072).copy_(h, non_blocking=True); torch.cuda.synchronize()
for name, fn in (("device", dev_call), ("host", host_call), ("h2d", h2d), ("device", dev_call), ("host", host_call)):
    fn()
    t = time.perf_counter(); fn(); fn(); dt = (time.perf_counter() - t) / 2
    print(name, "%.1f ms" % (dt * 1e3), flush=True)

#!/usr/bin/env python3
"""One device-resident verify_blob_kzg_proof_batch call per batch size given on the command line (for an ncu launch
list of the phase B kernels).  Not part of the product."""
import ctypes, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kzg_rust_b200 as k
from golden_util import golden
g = golden(); L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, int(os.environ.get("PROBE_WIDTH", "16")))
sizes = [int(a) for a in sys.argv[1:]] or [6, 1024, 16384]
n = max(sizes)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(7)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev, generator=gen); blobs[:, :, 0] = 0
cm = torch.zeros((n, 48), dtype=torch.uint8, device=dev); pr = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev)
assert L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), n, cm.data_ptr(), st.data_ptr()) == 0
assert L.kzg_b200_compute_blob_kzg_proof_device(s._h, blobs.data_ptr(), cm.data_ptr(), n, pr.data_ptr(), st.data_ptr()) == 0
L.kzg_b200_synchronize(s._h)
ok = ctypes.c_int(0)
for m in sizes:
    for _ in range(2):
        assert L.kzg_b200_verify_blob_kzg_proof_batch_device(s._h, blobs.data_ptr(), cm.data_ptr(), pr.data_ptr(), m, ctypes.byref(ok)) == 0
        assert ok.value == 1
    print("n=%d ok" % m, flush=True)

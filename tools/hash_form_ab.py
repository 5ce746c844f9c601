#!/usr/bin/env python3
"""A/B of the challenge-hash kernel forms (KZG_B200_CHALLENGE_G) on whole verification calls: wall time per call for
host buffers and device-resident blobs.  Not part of the product."""
import ctypes, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kzg_rust_b200 as k
from golden_util import golden
g = golden(); L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 16)
n = 4096
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
ok = ctypes.c_int(0)
for m in (6, 64, 256, 1024, 2048, 4096):
    for form in (os.environ.get("AB_FORMS", "0,1,2,4,8,16,32").split(",")):
        if int(form) * m > 32 * 2400:
            continue
        os.environ["KZG_B200_CHALLENGE_G"] = form
        res = []
        for fn in (lambda: k.Kzg.verify_blob_kzg_proof_batch_raw(vb[:m], vc[:m], vp[:m], m, s),
                   lambda: L.kzg_b200_verify_blob_kzg_proof_batch_device(s._h, blobs.data_ptr(), cm.data_ptr(), pr.data_ptr(), m, ctypes.byref(ok))):
            fn()
            ts = []
            for _ in range(8):
                t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
            res.append((min(ts), sum(ts) / len(ts)))
        print("n=%5d G=%2s  host %.2f ms (mean %.2f)  device %.2f ms (mean %.2f)" % (m, form, res[0][0], res[0][1], res[1][0], res[1][1]), flush=True)

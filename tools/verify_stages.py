#!/usr/bin/env python3
"""Where the time of verify_blob_kzg_proof_batch goes: wall time per call and per-stage device time (profiling
mode keeps the stages apart) for several batch sizes, host buffers and device-resident, plus the two host pieces
(the hash of compute_r_powers and the final pairing check).  Not part of the product."""
import ctypes, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import kzg_rust_b200 as k
from golden_util import golden
STAGES = ["digits", "msm_gather", "msm_tree", "compress", "challenge", "eval", "validate", "verify_terms"]
g = golden(); L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 0)
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
ms = (ctypes.c_double * 8)(); ln = (ctypes.c_uint64 * 8)()

def stages(fn, label, reps=3):
    fn()
    t = time.perf_counter()
    for _ in range(reps): fn()
    wall = (time.perf_counter() - t) / reps * 1e3
    L.kzg_b200_profile_enable(s._h, 1); fn(); L.kzg_b200_profile_read(s._h, ms, ln); L.kzg_b200_profile_enable(s._h, 0)
    print("%-28s wall %7.2f ms  %s  device sum %.2f" % (label, wall, {STAGES[i]: round(ms[i], 2) for i in range(8) if ms[i] > 0}, sum(ms)), flush=True)

for m in (1, 6, 64, 1024, 4096, n):
    if m <= n:
        stages(lambda: k.Kzg.verify_blob_kzg_proof_batch_raw(vb[:m], vc[:m], vp[:m], m, s), "verify host n=%d" % m)
ok = ctypes.c_int(0)
for m in sorted({min(6, n), min(1024, n), n}):
    stages(lambda: L.kzg_b200_verify_blob_kzg_proof_batch_device(s._h, blobs.data_ptr(), cm.data_ptr(), pr.data_ptr(), m, ctypes.byref(ok)), "verify device n=%d" % m)
# host pieces
zy = np.zeros((n, 64), dtype=np.uint8); r = np.zeros(32, dtype=np.uint8)
t = time.perf_counter()
for _ in range(3): L.kzg_b200_compute_r(s._h, vc.ctypes.data, zy.ctypes.data, vp.ctypes.data, n, r.ctypes.data)
print("compute_r over %d blobs: %.2f ms" % (n, (time.perf_counter() - t) / 3 * 1e3))
a1 = vc[0].tobytes(); g2 = g.g2_bytes[:96]
L.kzg_b200_pairings_verify.argtypes = [ctypes.c_char_p] * 4 + [ctypes.POINTER(ctypes.c_int)]
t = time.perf_counter()
for _ in range(5): L.kzg_b200_pairings_verify(a1, g2, a1, g2, ctypes.byref(ok))
print("pairings_verify (2 decompressions + 2 Miller loops + final exp): %.2f ms" % ((time.perf_counter() - t) / 5 * 1e3))
part = np.zeros(224, dtype=np.uint8); part[0] = 0x40; part[96] = 0x40
t = time.perf_counter()
for _ in range(5): L.kzg_b200_verify_finish(s._h, part.ctypes.data, 1, ctypes.byref(ok))
print("verify_finish on an empty partial: %.2f ms" % ((time.perf_counter() - t) / 5 * 1e3))
# single calls of the other entry points
stages(lambda: k.Kzg.blob_to_kzg_commitment_batch(vb[:1], s), "commit host n=1", 5)
stages(lambda: k.Kzg.compute_blob_kzg_proof_batch(vb[:1], vc[:1], s), "blob proof host n=1", 5)
stages(lambda: k.Kzg.blob_to_kzg_commitment_batch(vb[:64], s), "commit host n=64", 3)

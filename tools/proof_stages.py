#!/usr/bin/env python3
"""Per-stage device times of compute_blob_kzg_proof and verify (profiling mode runs chunks one at a time)."""
import ctypes, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kzg_rust_b200 as k
from golden_util import golden
STAGES = ["digits", "msm_gather", "msm_tree", "compress", "challenge", "eval", "validate", "verify_terms"]
g = golden(); L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev); blobs[:, :, 0] = 0
cm = torch.zeros((n, 48), dtype=torch.uint8, device=dev); pr = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev)
L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), n, cm.data_ptr(), st.data_ptr()); L.kzg_b200_synchronize(s._h)
def proof():
    rc = L.kzg_b200_compute_blob_kzg_proof_device(s._h, blobs.data_ptr(), cm.data_ptr(), n, pr.data_ptr(), st.data_ptr()); assert rc == 0
    L.kzg_b200_synchronize(s._h)
proof()
t = time.perf_counter(); proof(); dt = time.perf_counter() - t
print("proof: %.1f ms for %d blobs = %.0f blobs/s" % (dt * 1e3, n, n / dt))
L.kzg_b200_profile_enable(s._h, 1); proof()
ms = (ctypes.c_double * 8)(); ln = (ctypes.c_uint64 * 8)()
L.kzg_b200_profile_read(s._h, ms, ln); L.kzg_b200_profile_enable(s._h, 0)
print("proof stages ms:", {STAGES[i]: round(ms[i], 1) for i in range(8) if ms[i] > 0})
# verify
pin = lambda t_: torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True).copy_(t_).numpy()
nv = min(n, 4096)
vb, vc, vp = pin(blobs[:nv].reshape(nv, 131072)), pin(cm[:nv]), pin(pr[:nv])
assert k.Kzg.verify_blob_kzg_proof_batch_raw(vb, vc, vp, nv, s)
t = time.perf_counter(); k.Kzg.verify_blob_kzg_proof_batch_raw(vb, vc, vp, nv, s); dt = time.perf_counter() - t
print("verify: %.1f ms for %d blobs = %.0f blobs/s" % (dt * 1e3, nv, nv / dt))
L.kzg_b200_profile_enable(s._h, 1); k.Kzg.verify_blob_kzg_proof_batch_raw(vb, vc, vp, nv, s)
L.kzg_b200_profile_read(s._h, ms, ln); L.kzg_b200_profile_enable(s._h, 0)
print("verify stages ms:", {STAGES[i]: round(ms[i], 1) for i in range(8) if ms[i] > 0})

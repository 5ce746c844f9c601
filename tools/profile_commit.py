#!/usr/bin/env python3
"""Short commit run for `ncu --set full` captures (not part of the product).

    python tools/profile_commit.py [comb_width] [blobs] [calls]

Creates a context, then runs `kzg_b200_blob_to_kzg_commitment_device` `calls` times over the
same device-resident synthetic blobs.  One call over <= 4096 blobs is one chunk: 1 digit
kernel, 1 gather-level batch_add launch, 11 pair-level launches, 1 Horner/compress launch."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import kzg_rust_b200 as k  # noqa: E402
from golden_util import golden  # noqa: E402

c = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = golden()
L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, c)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(0xB200)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev, generator=gen)
blobs[:, :, 0] = 0
out = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for _ in range(calls):
    rc = L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), n, out.data_ptr(), st.data_ptr())
    assert rc == 0, rc
    L.kzg_b200_synchronize(s._h)
assert not bool(st.any().item())
print("profile_commit ok: c=%d n=%d calls=%d" % (s.comb_width, n, calls))
s.close()

#!/usr/bin/env python3
"""The reference's own `cargo bench` (benches/kzg_benches.rs:25-126), name for name, through this repo's
mirror of `impl Kzg` (host buffers in, host results out, one call per sample): median latency of a single
call, and the verify_blob_kzg_proof_batch group for 1..64 blobs.  BASELINE.json configs[0].

    python tools/criterion_bench.py [samples=100] [--cpu]     # --cpu adds the oracle port on one host core

One JSON line per benchmark."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import kzg_rust_b200 as k  # noqa: E402
from golden_util import golden  # noqa: E402
from gpu_util import synthetic_blobs  # noqa: E402

samples = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 100
with_cpu = "--cpu" in sys.argv
g = golden()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, int(os.environ.get("KZG_CRITERION_COMB_WIDTH", 0)))
max_count = 64                                    # benches/kzg_benches.rs:26
blobs = synthetic_blobs(max_count, seed=0xC817)   # :14-23, seeded
cms, st = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
proofs, st2 = k.Kzg.compute_blob_kzg_proof_batch(blobs, cms, s)
assert not st.any() and not st2.any()
rng = np.random.default_rng(5)
field = rng.integers(0, 256, size=32, dtype=np.uint8)
field[0] = 0
blob0, cm0, pr0 = blobs[0].tobytes(), cms[0].tobytes(), proofs[0].tobytes()


def bench(name, fn, n_elems=None, reps=samples):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    line = {"bench": name, "median_ms": round(med * 1e3, 4), "min_ms": round(min(ts) * 1e3, 4), "samples": reps,
            "comb_width": s.comb_width}
    if n_elems:
        line["elements_per_s"] = round(n_elems / med, 1)
    print(json.dumps(line), flush=True)
    return med


bench("blob_to_kzg_commitment", lambda: k.Kzg.blob_to_kzg_commitment(blob0, s))
bench("compute_kzg_proof", lambda: k.Kzg.compute_kzg_proof(blob0, field.tobytes(), s))
bench("compute_blob_kzg_proof", lambda: k.Kzg.compute_blob_kzg_proof(blob0, cm0, s))
# the reference passes the same random field element as z and y (benches/kzg_benches.rs:70-79): the verdict is False
bench("verify_kzg_proof", lambda: k.Kzg.verify_kzg_proof(cm0, field.tobytes(), field.tobytes(), pr0, s))
assert k.Kzg.verify_blob_kzg_proof(blob0, cm0, pr0, s) is True
bench("verify_blob_kzg_proof", lambda: k.Kzg.verify_blob_kzg_proof(blob0, cm0, pr0, s))
for count in (1, 2, 4, 8, 16, 32, 64):
    b, c, p = blobs[:count], cms[:count], proofs[:count]
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(b, c, p, count, s) is True
    bench("verify_blob_kzg_proof_batch/%d" % count, lambda: k.Kzg.verify_blob_kzg_proof_batch_raw(b, c, p, count, s), n_elems=count)

if with_cpu:
    from gpu_util import oracle_settings
    o = oracle_settings("mainnet")
    reps = 5

    def cpu(name, fn, n_elems=None):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        line = {"bench": name + " [CPU restatement, 1 core, not blst]", "median_ms": round(statistics.median(ts) * 1e3, 3), "samples": reps}
        if n_elems:
            line["elements_per_s"] = round(n_elems / statistics.median(ts), 1)
        print(json.dumps(line), flush=True)

    cpu("blob_to_kzg_commitment", lambda: o.blob_to_kzg_commitment(blob0))
    cpu("compute_blob_kzg_proof", lambda: o.compute_blob_kzg_proof(blob0, cm0))
    cpu("verify_blob_kzg_proof", lambda: o.verify_blob_kzg_proof(blob0, cm0, pr0))
    cpu("verify_blob_kzg_proof_batch/64", lambda: o.verify_blob_kzg_proof_batch([b.tobytes() for b in blobs], [c.tobytes() for c in cms], [p.tobytes() for p in proofs]), n_elems=64)
s.close()

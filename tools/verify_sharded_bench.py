#!/usr/bin/env python3
"""BASELINE.json configs[4]: verify_blob_kzg_proof_batch over a batch sharded across the GPUs of one box
(one rank per GPU, NCCL), partial-G1-sum combination and ONE host pairing check.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/verify_sharded_bench.py [total_blobs=16384]

Every rank makes its shard of synthetic blobs, commits and proves them on its GPU, then the batch is
verified through kzg_rust_b200.sharded (phase A -> r -> phase B -> finish); a second run with two proofs
swapped on the last rank must come out False on every rank.  Rank 0 prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import kzg_rust_b200 as k  # noqa: E402
from golden_util import golden  # noqa: E402
from kzg_rust_b200.sharded import CudaBackend, shard_range, verify_blob_kzg_proof_batch_sharded  # noqa: E402

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE", "ABORT"):
        os.environ["NCCL_DEBUG"] = "NONE"  # VERSION and WARN both print the banner (on stdout)
    dist.init_process_group("nccl", device_id=dev)
g = golden()
L = k.load_library()
s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, local, int(os.environ.get("KZG_VERIFY_WINDOW_BITS", 16)))
lo, hi = shard_range(n_total, rank, world)
n = hi - lo
gen = torch.Generator(device=dev)
gen.manual_seed(0xB200 + rank)
d_blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev, generator=gen)
d_blobs[:, :, 0] = 0
d_cm = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
d_pr = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
d_st = torch.zeros(n, dtype=torch.int32, device=dev)
assert L.kzg_b200_blob_to_kzg_commitment_device(s._h, d_blobs.data_ptr(), n, d_cm.data_ptr(), d_st.data_ptr()) == 0
assert L.kzg_b200_compute_blob_kzg_proof_device(s._h, d_blobs.data_ptr(), d_cm.data_ptr(), n, d_pr.data_ptr(), d_st.data_ptr()) == 0
L.kzg_b200_synchronize(s._h)
assert not bool(d_st.any().item())
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t).numpy()
blobs, cms, prs = pin(d_blobs.reshape(n, 131072)), pin(d_cm), pin(d_pr)
DEVICE = os.environ.get("SHARDS", "host") == "device"  # SHARDS=device: every rank's shard stays in its GPU's memory
if not DEVICE:
    del d_blobs
backend = CudaBackend(s)


TRACE = {}
DEV_CACHE = {}


def run(proofs):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    if DEVICE:
        from kzg_rust_b200.sharded import verify_blob_kzg_proof_batch_sharded_device
        d_proofs = DEV_CACHE.setdefault(id(proofs), torch.from_numpy(proofs).to(dev))
        ok = verify_blob_kzg_proof_batch_sharded_device(backend, d_blobs, d_cm, d_proofs, n_total, dev, trace=TRACE)
    else:
        ok = verify_blob_kzg_proof_batch_sharded(backend, blobs, cms, proofs, n_total, device=dev, trace=TRACE)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return ok, dt


ok, _ = run(prs)            # warm-up
assert ok is True
times = []
TRACE.clear()
for _ in range(3):
    ok, dt = run(prs)
    assert ok is True
    times.append(dt)
steps = dict(TRACE)
bad = prs.copy()
if rank == world - 1 and n >= 2:
    bad[[0, n - 1]] = bad[[n - 1, 0]]
ok_bad, _ = run(bad)
assert ok_bad is False
if rank == 0:
    best = min(times)
    print(json.dumps({"metric": "verify_blob_kzg_proof_batch throughput (sharded, one verdict)", "value": n_total / best,
                      "unit": "blobs/s", "n_gpus": world, "blobs": n_total, "ms_per_call": best * 1e3,
                      "all_ms": [round(t * 1e3, 2) for t in times], "negative_control_rejected": True,
                      "comb_width": s.comb_width, "rank0_step_ms": {k_: round(v / 3, 2) for k_, v in steps.items()}, "timing": "wall clock around the blocking call, max over ranks, " + ("device-resident shards" if DEVICE else "host buffers")}), flush=True)
s.close()
if world > 1:
    dist.destroy_process_group()

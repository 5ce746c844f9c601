"""Blob batches sharded over the GPUs of one box: one process (rank) and one context per GPU.

Blobs are independent, so commitments and proofs need no exchange at all: every rank runs the
batched entry point on its contiguous range.  Batch verification (reference
verify_blob_kzg_proof_batch, src/kzg.rs:637-693) has one exchange step, because the random
challenge r hashes every blob's (C_i, z_i, y_i, proof_i) (compute_r_powers,
src/utils.rs:426-474):

    phase A (local)   validate, z_i, y_i                           kzg_b200_verify_phase_a
    exchange 1        ONE all_gather: status + 160-byte records (C_i, z_i, y_i, proof_i)
    r                 one sequential SHA-256 over all records      kzg_b200_compute_r
    phase B (local)   partial sums with r^(first + i)              kzg_b200_verify_phase_b
    exchange 2        ONE all_gather: status + one 224-byte partial per rank
    finish            add the partials, one pairing check (host)   kzg_b200_verify_finish

The payloads are a few hundred bytes per rank, so the collective is plain
`torch.distributed.all_gather` (NCCL between GPUs, gloo in the CPU tests); there is no tile
stream to fuse with it.  The arithmetic is behind a small backend interface so the driver can
be exercised without a GPU (tests/test_multi_gpu.py); `CudaBackend` is the product path.
"""
import ctypes

import numpy as np

from . import kzg as _k


def shard_range(n, rank, world):
    """Contiguous, balanced ranges: the first n % world ranks get one blob more."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class CudaBackend:
    """The four calls of the two-phase protocol on one context (one GPU)."""

    def __init__(self, settings):
        self.s = settings
        self.L = _k.load_library()

    def phase_a(self, blobs, commitments, proofs):
        n = len(commitments) // 48
        zy = np.zeros((n, 64), dtype=np.uint8)
        rc = self.L.kzg_b200_verify_phase_a(self.s._h, blobs.ctypes.data, commitments.ctypes.data, proofs.ctypes.data, n,
                                            zy.ctypes.data)
        return rc, zy.reshape(-1)

    def phase_a_device(self, d_blobs, d_commitments, d_proofs, n):
        """The same for a shard that is already in this GPU's memory (device pointers, 16-byte aligned).  Returns
        (rc, zy, commitments, proofs) with the compressed points copied to the host for the exchange."""
        zy = np.zeros(64 * n, dtype=np.uint8)
        cm = np.zeros(48 * n, dtype=np.uint8)
        pr = np.zeros(48 * n, dtype=np.uint8)
        rc = self.L.kzg_b200_verify_phase_a_device(self.s._h, d_blobs, d_commitments, d_proofs, n, zy.ctypes.data, cm.ctypes.data,
                                                   pr.ctypes.data)
        return rc, zy, cm, pr

    def compute_r(self, commitments, zy, proofs):
        r = np.zeros(32, dtype=np.uint8)
        rc = self.L.kzg_b200_compute_r(self.s._h, commitments.ctypes.data, zy.ctypes.data, proofs.ctypes.data,
                                       len(commitments) // 48, r.ctypes.data)
        if rc:
            _k._raise(rc, "compute_r")
        return r

    def phase_b(self, commitments, zy, proofs, r, first_index):
        part = np.zeros(224, dtype=np.uint8)
        rc = self.L.kzg_b200_verify_phase_b(self.s._h, commitments.ctypes.data, zy.ctypes.data, proofs.ctypes.data,
                                            len(commitments) // 48, r.ctypes.data, first_index, part.ctypes.data)
        return rc, part

    def finish(self, partials):
        ok = ctypes.c_int(0)
        rc = self.L.kzg_b200_verify_finish(self.s._h, partials.ctypes.data, len(partials) // 224, ctypes.byref(ok))
        if rc:
            _k._raise(rc, "verify_finish")
        return bool(ok.value)


def _u8(x):
    return np.ascontiguousarray(np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else x.reshape(-1).view(np.uint8))


def _all_gather_bytes(local, sizes, device):
    """all_gather of variable-length byte arrays (sizes known on every rank)."""
    import torch
    import torch.distributed as dist
    width = max(max(sizes), 1)
    buf = torch.zeros(width, dtype=torch.uint8)
    buf[:local.size] = torch.from_numpy(local.copy()) if local.size else buf[:0]
    buf = buf.to(device)
    outs = [torch.empty(width, dtype=torch.uint8, device=device) for _ in sizes]
    dist.all_gather(outs, buf)
    return [o.cpu().numpy()[:sz] for o, sz in zip(outs, sizes)] if sizes else []


def verify_blob_kzg_proof_batch_sharded(backend, blobs, commitments, proofs, n_total, device="cpu", trace=None):
    """Every rank passes ITS contiguous shard (shard_range(n_total, rank, world)) and gets the
    verdict for the whole batch.  Raises BadArgs on every rank if any shard holds a malformed blob,
    commitment or proof (the reference's first-error abort, src/kzg.rs:671-683, seen from outside).

    Two collectives per verdict: one all_gather of the shards' records (status byte + 160 B per blob: C_i, z_i,
    y_i, proof_i -- exactly what compute_r_powers hashes, so every rank derives the same r without a broadcast),
    and one all_gather of the 224-byte partial sums (+ status byte)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mark = _marker(trace)
    lo, hi = shard_range(n_total, rank, world)
    n_local = hi - lo
    blobs, commitments, proofs = _u8(blobs), _u8(commitments), _u8(proofs)
    if commitments.size != 48 * n_local or proofs.size != 48 * n_local:
        raise _k.BadArgs("shard does not match shard_range(n_total, rank, world)")
    if n_total == 0:
        return True
    rc, zy = backend.phase_a(blobs, commitments, proofs) if n_local else (0, np.zeros(0, np.uint8))
    mark("phase_a")
    return _exchange_and_finish(backend, rc, zy, commitments, proofs, n_total, lo, n_local, device, mark)


def verify_blob_kzg_proof_batch_sharded_device(backend, d_blobs, d_commitments, d_proofs, n_total, device, trace=None):
    """The same for shards that are already in each rank's GPU memory: d_* are contiguous torch uint8 CUDA tensors holding
    this rank's shard_range(n_total, rank, world) blobs, commitments and proofs.  Phase A reads them where they are
    (no upload); 160 bytes per blob come to the host for the exchange, as in the host form."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mark = _marker(trace)
    lo, hi = shard_range(n_total, rank, world)
    n_local = hi - lo
    if d_commitments.numel() != 48 * n_local or d_proofs.numel() != 48 * n_local:
        raise _k.BadArgs("shard does not match shard_range(n_total, rank, world)")
    if n_total == 0:
        return True
    if n_local:
        rc, zy, commitments, proofs = backend.phase_a_device(d_blobs.data_ptr(), d_commitments.data_ptr(), d_proofs.data_ptr(), n_local)
    else:
        rc, zy, commitments, proofs = 0, np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.uint8)
    mark("phase_a")
    return _exchange_and_finish(backend, rc, zy, commitments, proofs, n_total, lo, n_local, device, mark)


def _marker(trace):
    import time
    t_last = [time.perf_counter()]

    def mark(name):  # trace: a dict that receives the wall-clock milliseconds of every step on this rank
        if trace is not None:
            now = time.perf_counter()
            trace[name] = trace.get(name, 0.0) + (now - t_last[0]) * 1e3
            t_last[0] = now
    return mark


def _exchange_and_finish(backend, rc, zy, commitments, proofs, n_total, lo, n_local, device, mark):
    """Everything after phase A: the records' all_gather, r, phase B, the partials' all_gather, the final check."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1:
        counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
        rec = np.zeros(1 + 160 * n_local, dtype=np.uint8)
        rec[0] = min(rc, 255)
        if n_local and not rc:
            body = rec[1:].reshape(n_local, 160)
            body[:, :48] = commitments.reshape(n_local, 48)
            body[:, 48:112] = zy.reshape(n_local, 64)
            body[:, 112:] = proofs.reshape(n_local, 48)
        recs = _all_gather_bytes(rec, [1 + 160 * c for c in counts], device)
        rc = max(int(r[0]) for r in recs)
        if rc:
            _k._raise(rc, "verify_blob_kzg_proof_batch (phase A)")
        body = np.concatenate([r[1:] for r in recs]).reshape(n_total, 160)
        all_c = np.ascontiguousarray(body[:, :48]).reshape(-1)
        all_zy = np.ascontiguousarray(body[:, 48:112]).reshape(-1)
        all_p = np.ascontiguousarray(body[:, 112:]).reshape(-1)
    else:
        if rc:
            _k._raise(rc, "verify_blob_kzg_proof_batch (phase A)")
        all_c, all_p, all_zy = commitments, proofs, zy
    mark("exchange_records")
    r = backend.compute_r(all_c, all_zy, all_p)       # every rank hashes the same bytes: no broadcast needed
    mark("compute_r")
    rc, part = backend.phase_b(commitments, zy, proofs, r, lo)
    mark("phase_b")
    if world > 1:
        rec = np.concatenate([np.array([min(rc, 255)], dtype=np.uint8), part])
        recs = _all_gather_bytes(rec, [225] * world, device)
        rc = max(int(x[0]) for x in recs)
        parts = np.concatenate([x[1:] for x in recs])
    else:
        parts = part
    mark("exchange_partials")
    if rc:
        _k._raise(rc, "verify_blob_kzg_proof_batch (phase B)")
    ok = backend.finish(parts)
    mark("finish")
    return ok

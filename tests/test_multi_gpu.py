"""The N > 1 path without GPUs: two gloo ranks drive kzg_rust_b200.sharded with a CPU backend
built from the oracle, and the verdict must equal the oracle's single-call
verify_blob_kzg_proof_batch.  What is under test is the driver: shard ranges, gather order,
one r for the whole batch, the per-shard first index, agreement on errors."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
G1_GEN = bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
INF = bytes([0xC0]) + bytes(47)


class OracleBackend:
    """CPU stand-in for CudaBackend: same four calls, arithmetic by the oracle.  The partial
    record is its own format (A || B compressed, s), consumed only by this class's finish()."""

    def __init__(self):
        from gpu_util import oracle_settings
        from golden_util import golden
        self.o = oracle_settings("mainnet")
        self.g2 = golden().g2_bytes

    def phase_a(self, blobs, commitments, proofs):
        from oracle import binding as ob
        n = commitments.size // 48
        zy = bytearray()
        for i in range(n):
            blob = blobs[131072 * i:131072 * (i + 1)].tobytes()
            c, p = commitments[48 * i:48 * i + 48].tobytes(), proofs[48 * i:48 * i + 48].tobytes()
            if not (ob.validate_kzg_g1(c) and ob.validate_kzg_g1(p)):
                return 1, np.zeros(64 * n, np.uint8)
            try:
                z = self.o.compute_challenge(blob, c)
                y = self.o.evaluate_polynomial(blob, z)
            except ob.OracleError:
                return 1, np.zeros(64 * n, np.uint8)
            zy += z + y
        return 0, np.frombuffer(bytes(zy), dtype=np.uint8)

    def compute_r(self, commitments, zy, proofs):
        n = commitments.size // 48
        h = hashlib.sha256(b"RCKZGBATCH___V1_" + (4096).to_bytes(8, "big") + n.to_bytes(8, "big"))
        for i in range(n):
            h.update(commitments[48 * i:48 * i + 48].tobytes() + zy[64 * i:64 * i + 64].tobytes() + proofs[48 * i:48 * i + 48].tobytes())
        return np.frombuffer((int.from_bytes(h.digest(), "big") % R).to_bytes(32, "big"), dtype=np.uint8)

    def phase_b(self, commitments, zy, proofs, r, first_index):
        from oracle import binding as ob
        n = commitments.size // 48
        rv = int.from_bytes(r.tobytes(), "big")
        cs = [commitments[48 * i:48 * i + 48].tobytes() for i in range(n)]
        ps = [proofs[48 * i:48 * i + 48].tobytes() for i in range(n)]
        zs = [int.from_bytes(zy[64 * i:64 * i + 32].tobytes(), "big") for i in range(n)]
        ys = [int.from_bytes(zy[64 * i + 32:64 * i + 64].tobytes(), "big") for i in range(n)]
        rp = [pow(rv, first_index + i, R) for i in range(n)]
        b32 = lambda v: (v % R).to_bytes(32, "big")
        A = ob.g1_lincomb(ps, [b32(x) for x in rp]) if n else INF
        B = ob.g1_lincomb(cs + ps, [b32(x) for x in rp] + [b32(x * z) for x, z in zip(rp, zs)]) if n else INF
        s = sum(x * y for x, y in zip(rp, ys)) % R
        return 0, np.frombuffer(A + B + bytes(96) + b32(s), dtype=np.uint8)

    def finish(self, partials):
        from oracle import binding as ob
        k = partials.size // 224
        recs = [partials[224 * i:224 * (i + 1)].tobytes() for i in range(k)]
        one = (1).to_bytes(32, "big")
        A = ob.g1_lincomb([r[:48] for r in recs], [one] * k)
        s = sum(int.from_bytes(r[192:224], "big") for r in recs) % R
        B = ob.g1_lincomb([r[48:96] for r in recs] + [G1_GEN], [one] * k + [((R - s) % R).to_bytes(32, "big")])
        return ob.pairings_verify(A, self.g2[96:192], B, self.g2[0:96])


def _worker(rank, world, port, n_total, tamper, bad_blob, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from kzg_rust_b200 import BadArgs
    from kzg_rust_b200.sharded import shard_range, verify_blob_kzg_proof_batch_sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        data = np.load(os.path.join(os.environ["KZG_TEST_TMP"], "batch.npz"))
        blobs, cms, proofs = data["blobs"].copy(), data["cms"].copy(), data["proofs"].copy()
        if tamper:
            proofs[[0, n_total - 1]] = proofs[[n_total - 1, 0]]
        if bad_blob:
            blobs[n_total - 1, :32] = 0xFF   # non-canonical field element in the LAST shard only
        lo, hi = shard_range(n_total, rank, world)
        try:
            out = verify_blob_kzg_proof_batch_sharded(OracleBackend(), blobs[lo:hi], cms[lo:hi], proofs[lo:hi], n_total)
        except BadArgs:
            out = "BadArgs"
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture(scope="module")
def batch(tmp_path_factory):
    from gpu_util import oracle_settings, synthetic_blobs
    o = oracle_settings("mainnet")
    n = 5
    blobs = synthetic_blobs(n, seed=2024)
    cms, st = o.blob_to_kzg_commitment_many(blobs, nthreads=os.cpu_count() or 1)
    proofs, st2 = o.compute_blob_kzg_proof_many(blobs, cms, nthreads=os.cpu_count() or 1)
    assert not st.any() and not st2.any()
    d = tmp_path_factory.mktemp("mgpu")
    np.savez(os.path.join(d, "batch.npz"), blobs=blobs, cms=cms, proofs=proofs)
    return str(d), n, (blobs, cms, proofs)


@pytest.mark.parametrize("tamper,bad_blob,expect", [(False, False, True), (True, False, False), (False, True, "BadArgs")])
def test_two_rank_sharded_verify(batch, tamper, bad_blob, expect):
    from gpu_util import oracle_settings
    d, n, (blobs, cms, proofs) = batch
    os.environ["KZG_TEST_TMP"] = d
    if expect != "BadArgs":
        # the single-call oracle agrees with what the two ranks must report
        pr = proofs.copy()
        if tamper:
            pr[[0, n - 1]] = pr[[n - 1, 0]]
        lists = lambda a: [a[i].tobytes() for i in range(len(a))]
        assert oracle_settings("mainnet").verify_blob_kzg_proof_batch(lists(blobs), lists(cms), lists(pr)) is expect
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, tamper, bad_blob, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: expect, 1: expect}


def test_shard_ranges_cover_the_batch():
    from kzg_rust_b200.sharded import shard_range
    for n in (0, 1, 5, 8, 16384, 65536):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in rs) - min(hi - lo for lo, hi in rs) <= 1

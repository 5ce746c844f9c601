"""-m gpu: compute_kzg_proof / compute_blob_kzg_proof through the C ABI (CUDA path) against the
reference's own vectors (tests/golden) and against the CPU oracle on seeded random blobs.
Bit-exact 48-byte proofs and 32-byte evaluations."""
import os

import numpy as np
import pytest

from golden_util import golden
from gpu_util import VECTOR_WIDTH_IDS, VECTOR_WIDTHS, gpu_settings, oracle_settings, synthetic_blobs

pytestmark = pytest.mark.gpu
G = golden()
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _kzg():
    import kzg_rust_b200
    return kzg_rust_b200


def _ids(fn):
    return [c["name"] for c in G.by_fn(fn)]


@pytest.mark.parametrize("width", VECTOR_WIDTHS, ids=VECTOR_WIDTH_IDS)
@pytest.mark.parametrize("case", G.by_fn("compute_kzg_proof"), ids=_ids("compute_kzg_proof"))
def test_compute_kzg_proof_vectors(case, width):
    """reference src/lib.rs:54-79 (36 valid cases, 18 of them with z inside the domain)."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    try:
        blob = k.Blob.from_bytes(G.get_bytes(case["input"]["blob"]))
        z = k.Bytes32.from_bytes(G.get_bytes(case["input"]["z"]))
    except (k.Error, ValueError):
        assert case["output"] is None
        return
    try:
        proof, y = k.Kzg.compute_kzg_proof(blob, z, s)
    except k.Error:
        assert case["output"] is None
        return
    assert ["0x" + proof.to_bytes().hex(), "0x" + y.to_bytes().hex()] == case["output"]


@pytest.mark.parametrize("width", VECTOR_WIDTHS, ids=VECTOR_WIDTH_IDS)
@pytest.mark.parametrize("case", G.by_fn("compute_blob_kzg_proof"), ids=_ids("compute_blob_kzg_proof"))
def test_compute_blob_kzg_proof_vectors(case, width):
    """reference src/lib.rs:81-105."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    try:
        blob = k.Blob.from_bytes(G.get_bytes(case["input"]["blob"]))
        c = k.Bytes48.from_bytes(G.get_bytes(case["input"]["commitment"]))
    except (k.Error, ValueError):
        assert case["output"] is None
        return
    try:
        proof = k.Kzg.compute_blob_kzg_proof(blob, c, s)
    except k.Error:
        assert case["output"] is None
        return
    assert "0x" + proof.to_bytes().hex() == case["output"]


@pytest.mark.parametrize("width", VECTOR_WIDTHS, ids=VECTOR_WIDTH_IDS)
def test_proof_vectors_as_one_batch(width):
    """All well-formed compute_blob_kzg_proof vectors in one batched call: per-blob status."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    cases = [c for c in G.by_fn("compute_blob_kzg_proof")
             if len(G.get_bytes(c["input"]["blob"])) == 131072 and len(G.get_bytes(c["input"]["commitment"])) == 48]
    blobs = b"".join(G.get_bytes(c["input"]["blob"]) for c in cases)
    cms = b"".join(G.get_bytes(c["input"]["commitment"]) for c in cases)
    out, status = k.Kzg.compute_blob_kzg_proof_batch(blobs, cms, s)
    for i, c in enumerate(cases):
        if c["output"] is None:
            assert status[i] == 1, c["name"]
        else:
            assert status[i] == 0 and "0x" + out[i].tobytes().hex() == c["output"], c["name"]


@pytest.mark.parametrize("comb_width", [6, 9, 0])
def test_config2_64_blob_batch_vs_oracle(comb_width):
    """BASELINE.json config 2: compute_blob_kzg_proof on a 64-blob synthetic batch, byte-equal
    with the CPU restatement, then verified as a batch (and rejected after one proof is swapped)."""
    k = _kzg()
    s = gpu_settings("mainnet", comb_width)
    o = oracle_settings("mainnet")
    blobs = synthetic_blobs(64, seed=0xB200)
    blobs[5, :] = 0                                           # zero polynomial: commitment and proof at infinity
    blobs[6, 64:96] = np.frombuffer((R - 1).to_bytes(32, "big"), dtype=np.uint8)
    cms, st = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
    assert not st.any()
    exp_c, _ = o.blob_to_kzg_commitment_many(blobs, nthreads=os.cpu_count() or 1)
    assert np.array_equal(cms, exp_c)
    proofs, st = k.Kzg.compute_blob_kzg_proof_batch(blobs, cms, s)
    assert not st.any()
    exp_p, est = o.compute_blob_kzg_proof_many(blobs, exp_c, nthreads=os.cpu_count() or 1)
    assert not est.any()
    assert np.array_equal(proofs, exp_p)
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, proofs, 64, s) is True
    bad = proofs.copy()
    bad[[10, 11]] = bad[[11, 10]]
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, bad, 64, s) is False


def test_compute_kzg_proof_random_z_vs_oracle():
    k = _kzg()
    s = gpu_settings("mainnet", 8)
    o = oracle_settings("mainnet")
    blobs = synthetic_blobs(12, seed=41)
    rng = np.random.default_rng(9)
    zs = [int.from_bytes(rng.bytes(32), "big") % R for _ in range(12)]
    zs[0], zs[1], zs[2] = 0, 1, R - 1      # 1 and r-1 are roots of unity of the domain
    zb = b"".join(z.to_bytes(32, "big") for z in zs)
    proofs, ys, st = k.Kzg.compute_kzg_proof_batch(blobs, zb, s)
    assert not st.any()
    for i in range(12):
        ep, ey = o.compute_kzg_proof(blobs[i].tobytes(), zs[i].to_bytes(32, "big"))
        assert proofs[i].tobytes() == ep and ys[i].tobytes() == ey, i


def test_minimal_preset_proofs_vs_oracle():
    """kzg_minimal (n = 4): parity unpinned by the reference, oracle only."""
    k = _kzg()
    s = gpu_settings("minimal", 6)
    o = oracle_settings("minimal")
    rng = np.random.default_rng(6)
    rows = [[int.from_bytes(rng.bytes(32), "big") % R for _ in range(4)] for _ in range(40)]
    rows[0] = [0, 0, 0, 0]
    blobs = b"".join(b"".join(v.to_bytes(32, "big") for v in row) for row in rows)
    cms, st = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
    assert not st.any()
    proofs, st = k.Kzg.compute_blob_kzg_proof_batch(blobs, cms.tobytes(), s)
    assert not st.any()
    for i in range(len(rows)):
        blob = blobs[128 * i:128 * i + 128]
        assert proofs[i].tobytes() == o.compute_blob_kzg_proof(blob, cms[i].tobytes()), i
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms.tobytes(), proofs.tobytes(), len(rows), s) is True
    # in-domain evaluation points of the 4-element domain
    w4 = 0x8d51ccce760304d0ec030002760300000001000000000000
    for z in (1, R - 1, w4, R - w4, 5):
        zb = z.to_bytes(32, "big")
        p, y, st = k.Kzg.compute_kzg_proof_batch(blobs[:128 * 4], zb * 4, s)
        assert not st.any()
        for i in range(4):
            ep, ey = o.compute_kzg_proof(blobs[128 * i:128 * i + 128], zb)
            assert p[i].tobytes() == ep and y[i].tobytes() == ey, (z, i)


def test_device_resident_calls_span_chunks_vs_oracle():
    """The *_device entry points (caller-owned device buffers, asynchronous on the context's stream) over
    more blobs than one workspace chunk: commitments and proofs byte-equal with the oracle.  The proof call
    hashes the Fiat-Shamir challenges of the whole call in one launch before the chunks start."""
    import torch
    k = _kzg()
    L = k.load_library()
    os.environ["KZG_B200_CHUNK"] = "24"
    try:
        s = k.KzgSettings.load_trusted_setup(G.g1_bytes, G.g2_bytes, 0, 7)
    finally:
        del os.environ["KZG_B200_CHUNK"]
    o = oracle_settings("mainnet")
    n = 64
    blobs = synthetic_blobs(n, seed=0xD3)
    blobs[7, :] = 0
    blobs[30] = blobs[29]
    dev = torch.device("cuda", 0)
    d_blobs = torch.from_numpy(blobs).to(dev)
    d_cm = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
    d_pr = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    assert L.kzg_b200_blob_to_kzg_commitment_device(s._h, d_blobs.data_ptr(), n, d_cm.data_ptr(), d_st.data_ptr()) == 0
    L.kzg_b200_synchronize(s._h)
    assert not bool(d_st.any().item())
    exp_c, est = o.blob_to_kzg_commitment_many(blobs, nthreads=os.cpu_count() or 1)
    assert not est.any() and np.array_equal(d_cm.cpu().numpy(), exp_c)
    for _ in range(2):  # twice: the second call reuses the grow-only challenge buffer
        assert L.kzg_b200_compute_blob_kzg_proof_device(s._h, d_blobs.data_ptr(), d_cm.data_ptr(), n, d_pr.data_ptr(), d_st.data_ptr()) == 0
        L.kzg_b200_synchronize(s._h)
        assert not bool(d_st.any().item())
        exp_p, est = o.compute_blob_kzg_proof_many(blobs, exp_c, nthreads=os.cpu_count() or 1)
        assert not est.any() and np.array_equal(d_pr.cpu().numpy(), exp_p)
        d_pr.zero_()
    # a commitment that is not a curve point and one outside the subgroup: only those blobs are refused
    # (the whole call's commitments are validated in one launch before the chunks start)
    bad = d_cm.clone()
    bad[5, 47] ^= 1
    bad[40, 1] ^= 0x55
    assert L.kzg_b200_compute_blob_kzg_proof_device(s._h, d_blobs.data_ptr(), bad.data_ptr(), n, d_pr.data_ptr(), d_st.data_ptr()) == 0
    L.kzg_b200_synchronize(s._h)
    st = d_st.cpu().numpy()
    from oracle.binding import validate_kzg_g1
    expect_bad = [i for i in (5, 40) if not validate_kzg_g1(bad[i].cpu().numpy().tobytes())]
    assert sorted(np.nonzero(st)[0].tolist()) == sorted(expect_bad) and len(expect_bad) >= 1
    s.close()

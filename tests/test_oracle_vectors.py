"""Pins the CPU oracle (oracle/kzg_oracle.c) against every known-answer vector the
reference ships (reference tests/*/small/*/data.yaml -> tests/golden/, 208 cases),
under the reference runner's contract (reference src/lib.rs:14-204): an input that
fails to parse, or an API `Err`, must coincide with `output: null`; otherwise the
outputs must be byte-equal."""
import os
import pytest

from golden_util import golden
from oracle.binding import OracleError, OracleSettings

G = golden()
BYTES_PER_BLOB = 131072


@pytest.fixture(scope="module")
def s():
    return OracleSettings.load_trusted_setup(G.g1_bytes, G.g2_bytes)


def _parse(case, key, want_len):
    """Mirror of the typed getters in reference src/test_formats/*.rs."""
    try:
        b = G.get_bytes(case["input"][key])
    except ValueError:
        return None
    return b if len(b) == want_len else None


def _hex(b):
    return "0x" + b.hex()


def _ids(fn):
    return [c["name"] for c in G.by_fn(fn)]


@pytest.mark.parametrize("case", G.by_fn("blob_to_kzg_commitment"), ids=_ids("blob_to_kzg_commitment"))
def test_blob_to_kzg_commitment(s, case):
    blob = _parse(case, "blob", BYTES_PER_BLOB)
    if blob is None:
        assert case["output"] is None
        return
    try:
        out = s.blob_to_kzg_commitment(blob)
    except OracleError:
        assert case["output"] is None
        return
    assert _hex(out) == case["output"]


@pytest.mark.parametrize("case", G.by_fn("compute_kzg_proof"), ids=_ids("compute_kzg_proof"))
def test_compute_kzg_proof(s, case):
    blob, z = _parse(case, "blob", BYTES_PER_BLOB), _parse(case, "z", 32)
    if blob is None or z is None:
        assert case["output"] is None
        return
    try:
        proof, y = s.compute_kzg_proof(blob, z)
    except OracleError:
        assert case["output"] is None
        return
    assert [_hex(proof), _hex(y)] == case["output"]


@pytest.mark.parametrize("case", G.by_fn("compute_blob_kzg_proof"), ids=_ids("compute_blob_kzg_proof"))
def test_compute_blob_kzg_proof(s, case):
    blob, c = _parse(case, "blob", BYTES_PER_BLOB), _parse(case, "commitment", 48)
    if blob is None or c is None:
        assert case["output"] is None
        return
    try:
        proof = s.compute_blob_kzg_proof(blob, c)
    except OracleError:
        assert case["output"] is None
        return
    assert _hex(proof) == case["output"]


@pytest.mark.parametrize("case", G.by_fn("verify_kzg_proof"), ids=_ids("verify_kzg_proof"))
def test_verify_kzg_proof(s, case):
    c, z, y, p = (_parse(case, "commitment", 48), _parse(case, "z", 32), _parse(case, "y", 32),
                  _parse(case, "proof", 48))
    if None in (c, z, y, p):
        assert case["output"] is None
        return
    try:
        ok = s.verify_kzg_proof(c, z, y, p)
    except OracleError:
        assert case["output"] is None
        return
    assert ok is case["output"]


@pytest.mark.parametrize("case", G.by_fn("verify_blob_kzg_proof"), ids=_ids("verify_blob_kzg_proof"))
def test_verify_blob_kzg_proof(s, case):
    blob, c, p = _parse(case, "blob", BYTES_PER_BLOB), _parse(case, "commitment", 48), _parse(case, "proof", 48)
    if None in (blob, c, p):
        assert case["output"] is None
        return
    try:
        ok = s.verify_blob_kzg_proof(blob, c, p)
    except OracleError:
        assert case["output"] is None
        return
    assert ok is case["output"]


@pytest.mark.parametrize("case", G.by_fn("verify_blob_kzg_proof_batch"), ids=_ids("verify_blob_kzg_proof_batch"))
def test_verify_blob_kzg_proof_batch(s, case):
    def many(key, want):
        out = []
        for v in case["input"][key]:
            try:
                b = G.get_bytes(v)
            except ValueError:
                return None
            if len(b) != want:
                return None
            out.append(b)
        return out

    blobs, cs, ps = many("blobs", BYTES_PER_BLOB), many("commitments", 48), many("proofs", 48)
    if None in (blobs, cs, ps):
        assert case["output"] is None
        return
    try:
        ok = s.verify_blob_kzg_proof_batch(blobs, cs, ps)
    except OracleError:
        assert case["output"] is None
        return
    assert ok is case["output"]


def test_fast_cpu_baseline_path_matches_the_checker():
    """bench.py times `blob_to_kzg_commitment_many_fast` (batch-affine buckets, mulx Montgomery) as the CPU baseline:
    it must give the checker's bytes -- on the reference's commitment vectors (status included), on the
    adversarially regular blobs and on random ones."""
    import numpy as np
    from gpu_util import oracle_settings, synthetic_blobs
    o = oracle_settings("mainnet")
    G = golden()
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    cases = [c for c in G.by_fn("blob_to_kzg_commitment") if len(G.get_bytes(c["input"]["blob"])) == 131072]
    vec = np.frombuffer(b"".join(G.get_bytes(c["input"]["blob"]) for c in cases), dtype=np.uint8).reshape(len(cases), 131072)
    extra = synthetic_blobs(6, seed=21)
    extra[0, :] = 0
    extra[1, :] = 0
    extra[1, 31::32] = 1
    extra[2] = np.frombuffer((R - 1).to_bytes(32, "big") * 4096, dtype=np.uint8)
    extra[3, 32 * 7:32 * 8] = 0xff
    blobs = np.concatenate([vec, extra], axis=0)
    exp, est = o.blob_to_kzg_commitment_many(blobs, nthreads=os.cpu_count() or 1)
    got, gst = o.blob_to_kzg_commitment_many_fast(blobs, nthreads=os.cpu_count() or 1)
    assert np.array_equal(est, gst)
    ok = est == 0
    assert ok.sum() >= 8 and np.array_equal(exp[ok], got[ok])
    for i, c in enumerate(cases):
        if c["output"] is None:
            assert gst[i] != 0
        else:
            assert "0x" + got[i].tobytes().hex() == c["output"]

"""-m gpu: verify_blob_kzg_proof / verify_blob_kzg_proof_batch through the C ABI against the
reference's vectors, plus the two-phase (sharded) form used for multi-GPU verification."""
import ctypes
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
@pytest.mark.parametrize("case", G.by_fn("verify_blob_kzg_proof"), ids=_ids("verify_blob_kzg_proof"))
def test_verify_blob_kzg_proof_vectors(case, width):
    """reference src/lib.rs:142-175."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    try:
        blob = k.Blob.from_bytes(G.get_bytes(case["input"]["blob"]))
        c = k.Bytes48.from_bytes(G.get_bytes(case["input"]["commitment"]))
        p = k.Bytes48.from_bytes(G.get_bytes(case["input"]["proof"]))
    except (k.Error, ValueError):
        assert case["output"] is None
        return
    try:
        ok = k.Kzg.verify_blob_kzg_proof(blob, c, p, s)
    except k.Error:
        assert case["output"] is None
        return
    assert ok is case["output"]


@pytest.mark.parametrize("width", VECTOR_WIDTHS, ids=VECTOR_WIDTH_IDS)
@pytest.mark.parametrize("case", G.by_fn("verify_blob_kzg_proof_batch"), ids=_ids("verify_blob_kzg_proof_batch"))
def test_verify_blob_kzg_proof_batch_vectors(case, width):
    """reference src/lib.rs:177-203."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    try:
        blobs = [k.Blob.from_bytes(G.get_bytes(v)) for v in case["input"]["blobs"]]
        cs = [k.Bytes48.from_bytes(G.get_bytes(v)) for v in case["input"]["commitments"]]
        ps = [k.Bytes48.from_bytes(G.get_bytes(v)) for v in case["input"]["proofs"]]
    except (k.Error, ValueError):
        assert case["output"] is None
        return
    try:
        ok = k.Kzg.verify_blob_kzg_proof_batch(blobs, cs, ps, s)
    except k.Error:
        assert case["output"] is None
        return
    assert ok is case["output"]


def _make_batch(k, s, n, seed):
    blobs = synthetic_blobs(n, seed=seed)
    cms, st = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
    assert not st.any()
    proofs, st = k.Kzg.compute_blob_kzg_proof_batch(blobs, cms, s)
    assert not st.any()
    return blobs, cms, proofs


def test_config3_batch_of_6_and_negative_control():
    """BASELINE.json config 3 at the Deneb block maximum; the oracle agrees on both outcomes."""
    k = _kzg()
    s = gpu_settings("mainnet", 8)
    o = oracle_settings("mainnet")
    blobs, cms, proofs = _make_batch(k, s, 6, 601)
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, proofs, 6, s) is True
    lists = lambda a: [a[i].tobytes() for i in range(len(a))]
    assert o.verify_blob_kzg_proof_batch(lists(blobs), lists(cms), lists(proofs)) is True
    bad = proofs.copy()
    bad[[0, 5]] = bad[[5, 0]]
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, bad, 6, s) is False
    assert o.verify_blob_kzg_proof_batch(lists(blobs), lists(cms), lists(bad)) is False
    # a single flipped blob byte changes z and y, so the proof no longer fits
    b2 = blobs.copy()
    b2[3, 100] ^= 1
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(b2, cms, proofs, 6, s) is False


def test_two_phase_sharded_verify_matches_single_call():
    """phase A per shard, one r for the whole batch, phase B per shard with the
    shard's first index, partial sums combined by verify_finish."""
    k = _kzg()
    L = k.load_library()
    s = gpu_settings("mainnet", 8)
    n = 11
    blobs, cms, proofs = _make_batch(k, s, n, 77)
    for tamper in (False, True):
        pr = proofs.copy()
        if tamper:
            pr[[2, 9]] = pr[[9, 2]]
        shards = [(0, 4), (4, 5), (5, 11)]
        zy = np.zeros((n, 64), dtype=np.uint8)
        for lo, hi in shards:
            rc = L.kzg_b200_verify_phase_a(s._h, blobs[lo:hi].ctypes.data, cms[lo:hi].ctypes.data, pr[lo:hi].ctypes.data,
                                           hi - lo, zy[lo:hi].ctypes.data)
            assert rc == 0
        r = np.zeros(32, dtype=np.uint8)
        assert L.kzg_b200_compute_r(s._h, cms.ctypes.data, zy.ctypes.data, pr.ctypes.data, n, r.ctypes.data) == 0
        partials = np.zeros((len(shards), 224), dtype=np.uint8)
        for j, (lo, hi) in enumerate(shards):
            rc = L.kzg_b200_verify_phase_b(s._h, cms[lo:hi].ctypes.data, zy[lo:hi].ctypes.data, pr[lo:hi].ctypes.data,
                                           hi - lo, r.ctypes.data, lo, partials[j].ctypes.data)
            assert rc == 0
        ok = ctypes.c_int(-1)
        assert L.kzg_b200_verify_finish(s._h, partials.ctypes.data, len(shards), ctypes.byref(ok)) == 0
        assert bool(ok.value) is (not tamper)
        assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, pr, n, s) is (not tamper)


@pytest.mark.parametrize("lanes_per_blob", [0, 32, 16, 8, 4, 2, 1])
def test_challenge_and_evaluation_match_oracle(lanes_per_blob):
    """z_i and y_i of phase A are bit-exact with compute_challenge / evaluate_polynomial_in_evaluation_form of the
    oracle, for every form of the hash kernel (G lanes per blob, csrc/frpath.cuh; 0 = the size rule, 1 = one thread
    per blob) and a batch that does not fill its last warp."""
    k = _kzg()
    L = k.load_library()
    s = gpu_settings("mainnet", 8)
    o = oracle_settings("mainnet")
    blobs, cms, proofs = _make_batch(k, s, 5, 5)
    zy = np.zeros((5, 64), dtype=np.uint8)
    if lanes_per_blob:
        os.environ["KZG_B200_CHALLENGE_G"] = str(lanes_per_blob)
    try:
        assert L.kzg_b200_verify_phase_a(s._h, blobs.ctypes.data, cms.ctypes.data, proofs.ctypes.data, 5, zy.ctypes.data) == 0
    finally:
        os.environ.pop("KZG_B200_CHALLENGE_G", None)
    for i in range(5):
        z = o.compute_challenge(blobs[i].tobytes(), cms[i].tobytes())
        assert zy[i, :32].tobytes() == z
        assert zy[i, 32:].tobytes() == o.evaluate_polynomial(blobs[i].tobytes(), z)


@pytest.mark.parametrize("slices", [8, 4, 2])
@pytest.mark.parametrize("lanes_per_blob", [0, 32, 8, 2])
def test_sliced_upload_hash_matches_oracle(lanes_per_blob, slices):
    """Host verification uploads a chunk's blobs in column slices and hashes slice k while slice k+1 is on the wire
    (csrc/proof_verify.inl: verify_chunk_a): z_i, y_i stay bit-exact with the oracle for every slice count and hash form."""
    k = _kzg()
    L = k.load_library()
    s = gpu_settings("mainnet", 8)
    o = oracle_settings("mainnet")
    blobs, cms, proofs = _make_batch(k, s, 5, 55)
    zy = np.zeros((5, 64), dtype=np.uint8)
    env = {"KZG_B200_SLICE_MIN_BLOBS": "1", "KZG_B200_HASH_SLICES": str(slices)}
    if lanes_per_blob:
        env["KZG_B200_CHALLENGE_G"] = str(lanes_per_blob)
    os.environ.update(env)
    try:
        assert L.kzg_b200_verify_phase_a(s._h, blobs.ctypes.data, cms.ctypes.data, proofs.ctypes.data, 5, zy.ctypes.data) == 0
        assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, proofs, 5, s) is True
    finally:
        for key in env:
            os.environ.pop(key, None)
    for i in range(5):
        z = o.compute_challenge(blobs[i].tobytes(), cms[i].tobytes())
        assert zy[i, :32].tobytes() == z
        assert zy[i, 32:].tobytes() == o.evaluate_polynomial(blobs[i].tobytes(), z)


def test_sliced_and_plain_uploads_agree_over_several_pieces():
    """300 blobs in pieces of 128 (two lanes, three upload slots, a short last piece): the (z, y) records of the sliced
    uploads equal those of plain copies, and the verdicts (true / one proof swapped / one blob byte flipped) too."""
    k = _kzg()
    L = k.load_library()
    s = gpu_settings("mainnet", 8)
    n = 300
    blobs, cms, proofs = _make_batch(k, s, n, 3001)
    out = {}
    for slices in (1, 8):
        env = {"KZG_B200_SLICE_MIN_BLOBS": "1", "KZG_B200_HASH_SLICES": str(slices), "KZG_B200_VERIFY_PIECE": "128"}
        os.environ.update(env)
        try:
            zy = np.zeros((n, 64), dtype=np.uint8)
            assert L.kzg_b200_verify_phase_a(s._h, blobs.ctypes.data, cms.ctypes.data, proofs.ctypes.data, n, zy.ctypes.data) == 0
            bad = proofs.copy()
            bad[[17, 290]] = bad[[290, 17]]
            b2 = blobs.copy()
            b2[299, 131071] ^= 1
            out[slices] = (zy.tobytes(), k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, proofs, n, s),
                           k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, bad, n, s),
                           k.Kzg.verify_blob_kzg_proof_batch_raw(b2, cms, proofs, n, s))
        finally:
            for key in env:
                os.environ.pop(key, None)
    assert out[1] == out[8]
    assert out[8][1:] == (True, False, False)


def test_pairing_check_matches_oracle():
    from oracle import binding as ob
    k = _kzg()
    L = k.load_library()
    g1 = [G.g1_bytes[48 * i:48 * i + 48] for i in range(4)]
    g2 = [G.g2_bytes[96 * i:96 * i + 96] for i in range(2)]
    tau_p = ob.g1_lincomb([g1[2]], [(1337).to_bytes(32, "big")])
    for a1, a2, b1, b2 in ((g1[2], g2[1], tau_p, g2[0]), (g1[1], g2[0], g1[0], g2[1]), (g1[3], g2[1], tau_p, g2[0])):
        ok = ctypes.c_int(-1)
        assert L.kzg_b200_pairings_verify(a1, a2, b1, b2, ctypes.byref(ok)) == 0
        assert bool(ok.value) is ob.pairings_verify(a1, a2, b1, b2)


def test_config5_size_16384_blobs_round_trip_properties():
    """BASELINE.json configs[3..4] size on one GPU, checked through size-independent properties: 16,384
    synthetic blobs are committed and proved on the device (default comb width, several chunks, both
    lanes), then (1) the whole batch verifies -- one pairing equation over all 16,384 (C_i, proof_i), which
    holds only if every proof opens its commitment at its own challenge, (2) the same batch with two proofs
    swapped is rejected, (3) repeated blobs give repeated commitments and proofs wherever they sit in the
    batch, (4) a sample is byte-equal with the oracle."""
    import ctypes
    import torch
    import kzg_rust_b200 as k
    L = k.load_library()
    s = gpu_settings("mainnet", 0)
    n = 16384
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0xC5)
    d_blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev, generator=gen)
    d_blobs[:, :, 0] = 0
    d_blobs[9000] = d_blobs[17]          # repeats across chunks and lanes
    d_blobs[16383] = d_blobs[4096]
    d_blobs[5000] = 0                    # the zero polynomial
    d_cm = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
    d_pr = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    assert L.kzg_b200_blob_to_kzg_commitment_device(s._h, d_blobs.data_ptr(), n, d_cm.data_ptr(), d_st.data_ptr()) == 0
    assert L.kzg_b200_compute_blob_kzg_proof_device(s._h, d_blobs.data_ptr(), d_cm.data_ptr(), n, d_pr.data_ptr(), d_st.data_ptr()) == 0
    L.kzg_b200_synchronize(s._h)
    assert not bool(d_st.any().item())
    cms, prs = d_cm.cpu().numpy(), d_pr.cpu().numpy()
    assert np.array_equal(cms[9000], cms[17]) and np.array_equal(prs[9000], prs[17])
    assert np.array_equal(cms[16383], cms[4096]) and np.array_equal(prs[16383], prs[4096])
    assert cms[5000].tobytes() == b"\xc0" + bytes(47) and prs[5000].tobytes() == b"\xc0" + bytes(47)
    blobs = torch.empty((n, 131072), dtype=torch.uint8, pin_memory=True).copy_(d_blobs.reshape(n, 131072)).numpy()
    del d_blobs
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, prs, n, s) is True
    bad = prs.copy()
    bad[[123, 15000]] = bad[[15000, 123]]
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, bad, n, s) is False
    idx = [0, 17, 4095, 4096, 5000, 8191, 12288, 16383]
    o = oracle_settings("mainnet")
    exp_c, est = o.blob_to_kzg_commitment_many(blobs[idx], nthreads=os.cpu_count() or 1)
    assert not est.any() and np.array_equal(cms[idx], exp_c)
    exp_p, est = o.compute_blob_kzg_proof_many(blobs[idx], exp_c, nthreads=os.cpu_count() or 1)
    assert not est.any() and np.array_equal(prs[idx], exp_p)


@pytest.mark.parametrize("width", VECTOR_WIDTHS, ids=VECTOR_WIDTH_IDS)
@pytest.mark.parametrize("case", G.by_fn("verify_kzg_proof"), ids=_ids("verify_kzg_proof"))
def test_verify_kzg_proof_vectors(case, width):
    """reference src/lib.rs:112-140: all 92 vectors (correct / incorrect proofs, points at infinity, z in and
    out of the domain, non-canonical z / y, commitments and proofs off the curve or outside G1) through
    kzg_b200_verify_kzg_proof -- phase B of the batch path with one term and the points subgroup-checked."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    try:
        c = k.Bytes48.from_bytes(G.get_bytes(case["input"]["commitment"]))
        z = k.Bytes32.from_bytes(G.get_bytes(case["input"]["z"]))
        y = k.Bytes32.from_bytes(G.get_bytes(case["input"]["y"]))
        p = k.Bytes48.from_bytes(G.get_bytes(case["input"]["proof"]))
    except (k.Error, ValueError):
        assert case["output"] is None
        return
    try:
        ok = k.Kzg.verify_kzg_proof(c, z, y, p, s)
    except k.Error:
        assert case["output"] is None
        return
    assert ok is case["output"]


def test_device_resident_verification_and_multi_chunk_host_calls():
    """kzg_b200_verify_blob_kzg_proof_batch_device (whole-call validation and challenges) and the host-buffer call
    over several chunks on two lanes agree: true for a good batch, false after a swap, BadArgs for a point off the
    curve / outside G1 and for a non-canonical blob element."""
    import torch
    k = _kzg()
    L = k.load_library()
    os.environ["KZG_B200_CHUNK"] = "24"
    try:
        s = k.KzgSettings.load_trusted_setup(G.g1_bytes, G.g2_bytes, 0, 7)
    finally:
        del os.environ["KZG_B200_CHUNK"]
    n = 100  # 5 chunks of 24
    blobs, cms, proofs = _make_batch(k, s, n, 4242)
    dev = torch.device("cuda", 0)
    d_b, d_c = torch.from_numpy(blobs).to(dev), torch.from_numpy(cms).to(dev)

    def device_verify(pr, bl=d_b, cm=d_c):
        d_p = torch.from_numpy(pr).to(dev)
        ok = ctypes.c_int(-1)
        rc = L.kzg_b200_verify_blob_kzg_proof_batch_device(s._h, bl.data_ptr(), cm.data_ptr(), d_p.data_ptr(), n, ctypes.byref(ok))
        return rc, bool(ok.value)

    assert device_verify(proofs) == (0, True)
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, proofs, n, s) is True
    bad = proofs.copy()
    bad[[3, 77]] = bad[[77, 3]]
    assert device_verify(bad) == (0, False)
    assert k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, bad, n, s) is False
    from oracle.binding import validate_kzg_g1
    bad = proofs.copy()
    bad[60, 47] ^= 1
    if not validate_kzg_g1(bad[60].tobytes()):
        assert device_verify(bad)[0] == 1
        with pytest.raises(k.BadArgs):
            k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, cms, bad, n, s)
    b2 = blobs.copy()
    b2[90, :32] = 0xff  # element >= r
    assert device_verify(proofs, bl=torch.from_numpy(b2).to(dev))[0] == 1
    with pytest.raises(k.BadArgs):
        k.Kzg.verify_blob_kzg_proof_batch_raw(b2, cms, proofs, n, s)
    # unaligned device pointers are refused, not faulted on
    ok = ctypes.c_int(-1)
    d_p = torch.from_numpy(proofs).to(dev)
    assert L.kzg_b200_verify_blob_kzg_proof_batch_device(s._h, d_b.data_ptr() + 4, d_c.data_ptr(), d_p.data_ptr(), n - 1, ctypes.byref(ok)) == 1
    s.close()


def test_phase_b_validates_what_it_is_given():
    """Phase B may run on another context, or on other data, than phase A: it subgroup-checks the points and range-checks
    z and y itself (a point of the curve outside G1 must not reach the ladders)."""
    k = _kzg()
    L = k.load_library()
    s = gpu_settings("mainnet", 8)
    n = 3
    blobs, cms, proofs = _make_batch(k, s, n, 99)
    zy = np.zeros((n, 64), dtype=np.uint8)
    assert L.kzg_b200_verify_phase_a(s._h, blobs.ctypes.data, cms.ctypes.data, proofs.ctypes.data, n, zy.ctypes.data) == 0
    r = np.zeros(32, dtype=np.uint8)
    assert L.kzg_b200_compute_r(s._h, cms.ctypes.data, zy.ctypes.data, proofs.ctypes.data, n, r.ctypes.data) == 0
    part = np.zeros(224, dtype=np.uint8)
    call = lambda c, z, p: L.kzg_b200_verify_phase_b(s._h, c.ctypes.data, z.ctypes.data, p.ctypes.data, n, r.ctypes.data, 0, part.ctypes.data)
    assert call(cms, zy, proofs) == 0
    # a point on the curve but not in G1: any x with x^3 + 4 a square (the cofactor is ~2^125)
    P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    from oracle.binding import validate_kzg_g1
    x = 5
    while True:
        rhs = (x ** 3 + 4) % P
        y = pow(rhs, (P + 1) // 4, P)
        if y * y % P == rhs:
            enc = bytearray(x.to_bytes(48, "big"))
            enc[0] |= 0x80 | (0x20 if y > (P - 1) // 2 else 0)
            if not validate_kzg_g1(bytes(enc)):
                break
        x += 1
    bad_point = np.frombuffer(bytes(enc), dtype=np.uint8)
    # ... which decompresses fine: with the subgroup check off it would be accepted as an operand
    pr = proofs.copy()
    pr[1] = bad_point
    assert call(cms, zy, pr) == 1
    cm = cms.copy()
    cm[2] = bad_point
    assert call(cm, zy, proofs) == 1
    z2 = zy.copy()
    z2[0, :32] = 0xff
    assert call(cms, z2, proofs) == 1
    z2 = zy.copy()
    z2[2, 32:] = np.frombuffer(R.to_bytes(32, "big"), dtype=np.uint8)
    assert call(cms, z2, proofs) == 1


def test_one_context_from_several_host_threads():
    """The reference API is re-entrant and `&KzgSettings` is shareable: commitments, proofs and
    verifications issued concurrently from four host threads on one context give the serial results."""
    import threading
    k = _kzg()
    s = gpu_settings("mainnet", 8)
    sets = [_make_batch(k, s, 5 + t, 500 + t) for t in range(4)]
    results, errors = [None] * 4, []

    def work(t):
        try:
            blobs, cms, proofs = sets[t]
            out = []
            for _ in range(3):
                c, st = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
                p, st2 = k.Kzg.compute_blob_kzg_proof_batch(blobs, c, s)
                ok = k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, c, p, len(c), s)
                bad = p.copy()
                bad[[0, 1]] = bad[[1, 0]]
                nok = k.Kzg.verify_blob_kzg_proof_batch_raw(blobs, c, bad, len(c), s)
                out.append((c.tobytes(), p.tobytes(), bool(st.any() or st2.any()), ok, nok))
            results[t] = out
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t in range(4):
        blobs, cms, proofs = sets[t]
        for c, p, st, ok, nok in results[t]:
            assert c == cms.tobytes() and p == proofs.tobytes() and not st and ok is True and nok is False


@pytest.mark.parametrize("n", [1, 2, 6, 11, 64, 300, 1024, 5000])
def test_phase_b_bucket_method_equals_the_ladders(n):
    """The bucket method (csrc/pippenger.cuh, the default) and one GLV ladder per term (KZG_B200_VERIFY_LADDER=1) give
    byte-identical 224-byte partial records for every window width the size rule picks and for forced ones; a
    shard's first index enters the powers.  Points at infinity and repeated points are in the batch."""
    k = _kzg()
    L = k.load_library()
    s = gpu_settings("mainnet", 8)
    m = min(n, 64)
    blobs, cms, proofs = _make_batch(k, s, m, 4000 + n)
    zy_small = np.zeros((m, 64), dtype=np.uint8)
    assert L.kzg_b200_verify_phase_a(s._h, blobs.ctypes.data, cms.ctypes.data, proofs.ctypes.data, m, zy_small.ctypes.data) == 0
    # phase B is a function of (C, z, y, proof) records alone: tile the valid records up to n
    rng = np.random.default_rng(n)
    pick = rng.integers(0, m, size=n) if n > m else np.arange(n)
    cm, pr, zy = cms[pick].copy(), proofs[pick].copy(), zy_small[pick].copy()
    zy[:, :] = np.frombuffer(b"".join((int.from_bytes(rng.bytes(32), "big") % R).to_bytes(32, "big") for _ in range(2 * n)),
                             dtype=np.uint8).reshape(n, 64)
    inf = np.frombuffer(b"\xc0" + bytes(47), dtype=np.uint8)
    if n >= 6:
        cm[1] = inf
        pr[2] = inf
        pr[4] = cm[4]
        zy[3, :32] = 0
    r = np.frombuffer((int.from_bytes(rng.bytes(32), "big") % R).to_bytes(32, "big"), dtype=np.uint8).copy()

    def partial(first, **env):
        out = np.zeros(224, dtype=np.uint8)
        for key, v in env.items():
            os.environ[key] = str(v)
        try:
            rc = L.kzg_b200_verify_phase_b(s._h, cm.ctypes.data, zy.ctypes.data, pr.ctypes.data, n, r.ctypes.data, first, out.ctypes.data)
        finally:
            for key in env:
                del os.environ[key]
        assert rc == 0
        return out.tobytes()

    for first in (0, 77):
        want = partial(first, KZG_B200_VERIFY_LADDER=1)
        assert partial(first) == want
        for c in (1, 2, 5, 9, 12):
            if n <= 300 or c >= 5:
                assert partial(first, KZG_B200_PIP_C=c) == want, (n, first, c)


def test_device_resident_phase_a_and_sharded_form():
    """kzg_b200_verify_phase_a_device: the same (z, y) records as the host-buffer phase A, the compressed points handed
    back for the exchange, phase B reusing what it validated; the sharded driver over device-resident shards (one rank)
    accepts the batch and rejects a tampered one; malformed input and misaligned pointers are refused."""
    import torch
    from kzg_rust_b200 import sharded
    k = _kzg()
    L = k.load_library()
    s = gpu_settings("mainnet", 8)
    n = 9
    blobs, cms, proofs = _make_batch(k, s, n, 1234)
    dev = torch.device("cuda", 0)
    d_b, d_c, d_p = (torch.from_numpy(a.copy()).to(dev) for a in (blobs, cms, proofs))
    zy_h = np.zeros((n, 64), dtype=np.uint8)
    assert L.kzg_b200_verify_phase_a(s._h, blobs.ctypes.data, cms.ctypes.data, proofs.ctypes.data, n, zy_h.ctypes.data) == 0
    be = sharded.CudaBackend(s)
    rc, zy, cm, pr = be.phase_a_device(d_b.data_ptr(), d_c.data_ptr(), d_p.data_ptr(), n)
    assert rc == 0 and zy.tobytes() == zy_h.tobytes() and cm.tobytes() == cms.tobytes() and pr.tobytes() == proofs.tobytes()
    r = be.compute_r(cm, zy, pr)
    rc, part = be.phase_b(cm, zy, pr, r, 0)
    assert rc == 0 and be.finish(part) is True
    assert sharded.verify_blob_kzg_proof_batch_sharded_device(be, d_b, d_c, d_p, n, dev) is True
    bad = d_p.clone()
    bad[[0, n - 1]] = bad[[n - 1, 0]]
    assert sharded.verify_blob_kzg_proof_batch_sharded_device(be, d_b, d_c, bad, n, dev) is False
    # a commitment that is not a point, a non-canonical blob element, a misaligned pointer
    broken = d_c.clone()
    broken[3, 5] ^= 0x55
    assert be.phase_a_device(d_b.data_ptr(), broken.data_ptr(), d_p.data_ptr(), n)[0] == 1
    nb = d_b.clone()
    nb[2, :32] = 0xff
    assert be.phase_a_device(nb.data_ptr(), d_c.data_ptr(), d_p.data_ptr(), n)[0] == 1
    assert be.phase_a_device(d_b.data_ptr() + 4, d_c.data_ptr(), d_p.data_ptr(), n - 1)[0] == 1

"""CPU tests (no GPU): the host code path of the product's own headers (the same templates
the CUDA kernels are built from), the host-side pairing, and the C-ABI surface.

  * field arithmetic of kzg_rust_b200/csrc/bigint.cuh vs Python integers
  * the whole commitment pipeline (comb table -> digits -> gather level -> tree levels -> Horner ->
    compress) walked thread by thread on the CPU for the n = 4 preset, vs the oracle
  * the signed comb recoding (sign words, bit transpose, table index) and the comb MSM vs a plain ladder
  * libkzg_b200.so loads and exports every symbol include/kzg_b200.h declares
  * the host pairing (stand-in for blst's) vs the oracle's
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from golden_util import golden
from gpu_util import minimal_setup_bytes, oracle_settings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_DIR = os.path.join(ROOT, "tests", "hostshim")
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _build_shim(name):
    src = os.path.join(SHIM_DIR, name + ".cpp")
    out = os.path.join(SHIM_DIR, "lib" + name + ".so")
    deps = [src] + [os.path.join(ROOT, "kzg_rust_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "kzg_rust_b200", "csrc"))]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)


@pytest.fixture(scope="module")
def field():
    return _build_shim("field_shim")


@pytest.fixture(scope="module")
def msm():
    return _build_shim("msm_shim")


def _limbs(v, n):
    return (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xffffffff for i in range(n)])


def _val(arr):
    return sum(int(x) << (32 * i) for i, x in enumerate(arr))


def test_fp_and_fr_arithmetic_vs_python(field):
    rng = np.random.default_rng(1)
    edge_p = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, 2 ** 380, 2 ** 381 - 1 - (2 ** 381 - 1 - (P - 1))]
    vals_p = edge_p + [int.from_bytes(rng.bytes(48), "big") % P for _ in range(200)]
    Rp = pow(2, 384, P)
    Rp_inv = pow(Rp, -1, P)
    out = (ctypes.c_uint32 * 12)()
    for i in range(len(vals_p) - 1):
        a, b = vals_p[i], vals_p[i + 1]
        field.shim_fp_mul(_limbs(a, 12), _limbs(b, 12), out)
        assert _val(out) == a * b * Rp_inv % P
        field.shim_fp_add(_limbs(a, 12), _limbs(b, 12), out)
        assert _val(out) == (a + b) % P
        field.shim_fp_sub(_limbs(a, 12), _limbs(b, 12), out)
        assert _val(out) == (a - b) % P
    for a in vals_p[1:60]:
        am = a * Rp % P
        field.shim_fp_inv(_limbs(am, 12), out)          # binary extended Euclid
        assert _val(out) == pow(a, -1, P) * Rp % P
        field.shim_fp_inv_fermat(_limbs(am, 12), out)   # a^(p-2)
        assert _val(out) == pow(a, -1, P) * Rp % P
        sq = a * a % P
        ok = field.shim_fp_sqrt(_limbs(sq * Rp % P, 12), out)
        assert ok and _val(out) * Rp_inv % P in (a, P - a)
        assert bool(field.shim_fp_large(_limbs(am, 12))) == (a > (P - 1) // 2)
    Rr = pow(2, 256, R)
    Rr_inv = pow(Rr, -1, R)
    vals_r = [0, 1, R - 1, R - 2, 2 ** 254] + [int.from_bytes(rng.bytes(32), "big") % R for _ in range(200)]
    out8 = (ctypes.c_uint32 * 8)()
    for i in range(len(vals_r) - 1):
        a, b = vals_r[i], vals_r[i + 1]
        field.shim_fr_mul(_limbs(a, 8), _limbs(b, 8), out8)
        assert _val(out8) == a * b * Rr_inv % R
        field.shim_fr_add(_limbs(a, 8), _limbs(b, 8), out8)
        assert _val(out8) == (a + b) % R
        field.shim_fr_sub(_limbs(a, 8), _limbs(b, 8), out8)
        assert _val(out8) == (a - b) % R
    for a in vals_r[1:60]:
        field.shim_fr_inv(_limbs(a * Rr % R, 8), out8)
        assert _val(out8) == pow(a, -1, R) * Rr % R
        field.shim_fr_inv_fermat(_limbs(a * Rr % R, 8), out8)
        assert _val(out8) == pow(a, -1, R) * Rr % R
    assert field.shim_fr_canonical(_limbs(R - 1, 8)) and not field.shim_fr_canonical(_limbs(R, 8))
    assert not field.shim_fr_canonical(_limbs(2 ** 256 - 1, 8))


def test_sign_word_recoding(msm):
    """Every scalar as sum_j (+-1) 2^j mod r (blobpath.cuh scalar_sign_words): the 255 signs reproduce it."""
    rng = np.random.default_rng(3)
    vals = [0, 1, 2, 3, R - 1, R - 2, 2 ** 254, 2 ** 254 + 1, 2 ** 248 - 1, (R - 1) // 2, (R + 1) // 2]
    vals += [int.from_bytes(rng.bytes(32), "big") % R for _ in range(200)]
    out = (ctypes.c_uint32 * 8)()
    for v in vals:
        msm.shim_sign_words(_limbs(v, 8), out)
        bits = _val(out)
        assert bits >> 255 == 0
        signed = sum((1 if (bits >> j) & 1 else -1) << j for j in range(255))
        assert signed % R == v, v
        assert signed in (v, v - R)


def test_bit_transpose_and_comb_digit(msm):
    rng = np.random.default_rng(4)
    a = [int(x) for x in rng.integers(0, 2 ** 32, size=32, dtype=np.uint64)]
    arr = (ctypes.c_uint32 * 32)(*a)
    msm.shim_transpose32(arr)
    for i in range(32):
        for j in range(32):
            assert (arr[i] >> j) & 1 == (a[j] >> i) & 1
    msm.shim_comb_digit.restype = ctypes.c_uint32
    for g in (1, 2, 5, 23, 24):
        for _ in range(50):
            pat = int(rng.integers(0, 2 ** g))
            d = msm.shim_comb_digit(pat, g)
            neg, idx = d >> 31, d & 0x7fffffff
            full = pat if not neg else (~pat) & ((1 << g) - 1)
            assert full & 1 == 1 and idx == full >> 1 and neg == (1 - (pat & 1))


@pytest.mark.parametrize("g,T,k", [(1, 4, 3), (2, 8, 4), (3, 32, 1), (4, 16, 64), (9, 4, 2)])
def test_commit_pipeline_on_cpu_minimal_preset(msm, g, T, k):
    """The product's own pipeline, thread by thread on the CPU, for kzg_minimal."""
    g1, _ = minimal_setup_bytes()
    o = oracle_settings("minimal")
    rng = np.random.default_rng(100 + g)
    rows = [[0, 0, 0, 0], [R - 1] * 4, [1, 0, 0, 0], [0, 0, 5, 0], [7, 7, 7, 7], [R, 0, 0, 0], [2, 2, 2, 2]]
    rows += [[int.from_bytes(rng.bytes(32), "big") % R for _ in range(4)] for _ in range(5)]
    blobs = b"".join(b"".join(v.to_bytes(32, "big") for v in row) for row in rows)
    B = len(rows)
    out = ctypes.create_string_buffer(48 * B)
    status = (ctypes.c_int * B)()
    assert msm.shim_commit(g1, 4, g, blobs, B, out, status, T, k) == 0
    from oracle.binding import OracleError
    for i in range(B):
        try:
            exp = o.blob_to_kzg_commitment(blobs[128 * i:128 * i + 128])
        except OracleError:
            assert status[i] == 1
            continue
        assert status[i] == 0 and out.raw[48 * i:48 * i + 48] == exp, i


@pytest.mark.parametrize("n,g", [(7, 3), (11, 4), (13, 5), (6, 6), (10, 7)])
def test_comb_msm_vs_plain_ladder(msm, n, g):
    """The comb over point counts that leave an odd number of groups, a padded last group and odd tree rows,
    on the first n mainnet setup points, against sum_i [s_i]P_i by double-and-add."""
    g1 = golden().g1_bytes[:48 * n]
    rng = np.random.default_rng(n * 100 + g)
    rows = [[0] * n, [1] * n, [R - 1] * n, [2] * n, [(i + 1) for i in range(n)]]
    rows += [[int.from_bytes(rng.bytes(32), "big") % R for _ in range(n)] for _ in range(2)]
    B = len(rows)
    sc = (ctypes.c_uint32 * (B * n * 8))()
    for b, row in enumerate(rows):
        for i, v in enumerate(row):
            for w in range(8):
                sc[(b * n + i) * 8 + w] = (v >> (32 * w)) & 0xffffffff
    oc, ol = ctypes.create_string_buffer(48 * B), ctypes.create_string_buffer(48 * B)
    assert msm.shim_msm_vs_ladder(g1, n, g, sc, B, oc, ol, 8, 5) == 0
    assert oc.raw == ol.raw
    assert oc.raw[:48] == bytes([0xc0]) + bytes(47)


def test_library_exports_every_declared_symbol():
    """No compute calls (no GPU here): the C-ABI library loads and exports what the header declares."""
    import kzg_rust_b200 as k
    L = k.load_library()
    with open(os.path.join(ROOT, "include", "kzg_b200.h")) as fh:
        text = fh.read()
    names = sorted(set(re.findall(r"\b(kzg_b200_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 20
    for nm in names:
        assert hasattr(L, nm), nm


def test_missing_device_is_reported_not_papered_over():
    """Without a usable CUDA device context creation fails with the CUDA error code; there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import kzg_rust_b200 as k
    g = golden()
    with pytest.raises(k.Error):
        k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 4)


def test_host_pairing_vs_oracle():
    from oracle import binding as ob
    import kzg_rust_b200 as k
    L = k.load_library()
    G = golden()
    g1 = [G.g1_bytes[48 * i:48 * i + 48] for i in range(4)]
    g2 = [G.g2_bytes[96 * i:96 * i + 96] for i in range(3)]
    tau = (1337).to_bytes(32, "big")
    tau_p = ob.g1_lincomb([g1[2]], [tau])
    inf = bytes([0xC0]) + bytes(47)
    cases = [(g1[2], g2[1], tau_p, g2[0]), (g1[1], g2[0], g1[0], g2[1]), (g1[3], g2[1], tau_p, g2[0]),
             (g1[2], g2[2], ob.g1_lincomb([tau_p], [tau]), g2[0]), (inf, g2[1], inf, g2[0]), (inf, g2[1], g1[0], g2[0])]
    for a1, a2, b1, b2 in cases:
        ok = ctypes.c_int(-1)
        assert L.kzg_b200_pairings_verify(a1, a2, b1, b2, ctypes.byref(ok)) == 0
        assert bool(ok.value) is ob.pairings_verify(a1, a2, b1, b2)
    ok = ctypes.c_int(-1)
    assert L.kzg_b200_pairings_verify(bytes(48), g2[1], g1[0], g2[0], ctypes.byref(ok)) != 0  # bit 7 clear: bad encoding


def test_host_types_mirror_reference_length_checks():
    """reference src/kzg.rs:107-117, 130-141, 160-173."""
    import kzg_rust_b200 as k
    with pytest.raises(k.Error):
        k.Blob.from_bytes(bytes(131071))
    with pytest.raises(k.Error):
        k.Bytes48.from_bytes(bytes(47))
    with pytest.raises(k.BadArgs):
        k.Bytes32.from_bytes(bytes(33))
    with pytest.raises(k.InvalidHexFormat):
        k.hex_to_bytes("0xzz")
    assert k.Bytes48.from_hex("0x" + "ab" * 48).to_bytes() == bytes([0xab]) * 48
    with pytest.raises(k.BadArgs):
        k.Kzg.verify_blob_kzg_proof_batch([k.Blob(bytes(131072))], [], [], None)
    assert k.Kzg.verify_blob_kzg_proof_batch([], [], [], None) is True


def _pack(vals, n):
    a = np.zeros((len(vals), n), dtype=np.uint32)
    for r, v in enumerate(vals):
        for i in range(n):
            a[r, i] = (v >> (32 * i)) & 0xffffffff
    return a


def _unpack(a):
    return [sum(int(x) << (32 * i) for i, x in enumerate(row)) for row in a]


def test_two_pipe_fp_multiplication_vs_python(field):
    """fp_hybrid.cuh (FP64-pipe product in 48-bit limbs + IMAD-pipe Montgomery reduction), host code
    path with fma() under FE_TOWARDZERO: the 768-bit product, the reduction on its own, and the
    product / square must equal Python integers -- and therefore fe_mul -- on edge and random values."""
    rng = np.random.default_rng(7)
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, 2 ** 380, 2 ** 381 - 1, 2 ** 48 - 1, 2 ** 48, (2 ** 48 - 1) * sum(2 ** (48 * k) for k in range(8)) % P,
            sum((2 ** 48 - 1) << (96 * k) for k in range(4)) % P, 0xffffffff, 2 ** 336 - 1, 2 ** 336]
    rand = [int.from_bytes(rng.bytes(48), "big") % P for _ in range(400)]
    av = edge + rand + [e for e in edge for _ in edge]
    bv = edge + rand[::-1] + [e for _ in edge for e in edge]
    n = len(av)
    a, b = _pack(av, 12), _pack(bv, 12)
    ptr = lambda x: x.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    Rinv = pow(pow(2, 384, P), -1, P)
    # exact products
    t = np.zeros((n, 24), dtype=np.uint32)
    field.shim_fp_product_many(ptr(a), ptr(b), ptr(t), ctypes.c_size_t(n), 0)
    assert _unpack(t) == [x * y for x, y in zip(av, bv)]
    field.shim_fp_product_many(ptr(a), ptr(a), ptr(t), ctypes.c_size_t(n), 1)
    assert _unpack(t) == [x * x for x in av]
    # the reduction alone, over its whole domain T < p * 2^384
    tv = [x * y for x, y in zip(av, bv)] + [P * 2 ** 384 - 1, (P - 1) * 2 ** 384, 2 ** 384 - 1, 2 ** 384, 0] + \
         [int.from_bytes(rng.bytes(96), "big") % (P << 384) for _ in range(300)]
    tt = _pack(tv, 24)
    r = np.zeros((len(tv), 12), dtype=np.uint32)
    field.shim_fp_redc_many(ptr(tt), ptr(r), ctypes.c_size_t(len(tv)))
    assert _unpack(r) == [v * Rinv % P for v in tv]
    # product and square against the integer-pipe fe_mul
    r1 = np.zeros((n, 12), dtype=np.uint32)
    r2 = np.zeros((n, 12), dtype=np.uint32)
    field.shim_fp_mul_hybrid_many(ptr(a), ptr(b), ptr(r1), ctypes.c_size_t(n))
    field.shim_fp_mul_many(ptr(a), ptr(b), ptr(r2), ctypes.c_size_t(n))
    assert np.array_equal(r1, r2)
    assert _unpack(r1) == [x * y * Rinv % P for x, y in zip(av, bv)]
    field.shim_fp_sqr_hybrid_many(ptr(a), ptr(r1), ctypes.c_size_t(n))
    assert _unpack(r1) == [x * x * Rinv % P for x in av]


def test_lazy_residues_vs_python(field):
    """fe_mul_lazy / fe_sub_lazy (values in [0, 2p), no final subtraction): results stay below 2p and are
    congruent to the exact ones; zero test and canonicalisation agree with Python."""
    rng = np.random.default_rng(11)
    edge = [0, 1, P - 1, P, P + 1, 2 * P - 1, 2 * P - 2, (P - 1) // 2, 2 ** 381, 2 ** 382 - 1 if 2 ** 382 - 1 < 2 * P else 2 * P - 3]
    rand = [int.from_bytes(rng.bytes(49), "big") % (2 * P) for _ in range(500)]
    av = rand + [e for e in edge for _ in edge]
    bv = rand[::-1] + [e for _ in edge for e in edge]
    n = len(av)
    a, b = _pack(av, 12), _pack(bv, 12)
    ptr = lambda x: x.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    Rinv = pow(pow(2, 384, P), -1, P)
    r = np.zeros((n, 12), dtype=np.uint32)
    field.shim_fp_mul_lazy_many(ptr(a), ptr(b), ptr(r), ctypes.c_size_t(n))
    got = _unpack(r)
    assert all(g < 2 * P for g in got)
    assert [g % P for g in got] == [x * y * Rinv % P for x, y in zip(av, bv)]
    field.shim_fp_sub_lazy_many(ptr(a), ptr(b), ptr(r), ctypes.c_size_t(n))
    got = _unpack(r)
    assert all(g < 2 * P for g in got)
    assert [g % P for g in got] == [(x - y) % P for x, y in zip(av, bv)]
    out = (ctypes.c_uint32 * 12)()
    for v in edge + rand[:50]:
        assert bool(field.shim_fp_is_zero_lazy(_limbs(v, 12))) == (v % P == 0)
        field.shim_fp_canonical(_limbs(v, 12), out)
        assert _val(out) == v % P


def test_glv_scalar_multiplication_vs_plain_ladder_and_oracle(msm):
    """g1j_mul_glv (k = k2 lambda + k1, joint 128-step ladder over P, psi(P), P + psi(P)) against the plain
    255-step ladder of the same header and against the oracle's scalar multiplication."""
    from oracle.binding import g1_lincomb
    lam = 0xd201000000010000 ** 2 - 1
    g = golden()
    pts = [g.g1_bytes[48 * i:48 * i + 48] for i in (0, 1, 77, 4095)] + [b"\xc0" + bytes(47)]
    rng = np.random.default_rng(21)
    ks = [0, 1, 2, lam - 1, lam, lam + 1, 2 * lam, lam * lam % R, R - 1, R - 2, 2 ** 128 - 1, 2 ** 128, 2 ** 254] + \
         [int.from_bytes(rng.bytes(32), "big") % R for _ in range(12)]
    a, b = (ctypes.c_uint8 * 48)(), (ctypes.c_uint8 * 48)()
    kk = (ctypes.c_uint32 * 8)()
    for pt in pts:
        for k in ks:
            assert msm.shim_g1_mul_both(pt, _limbs(k, 8), a, b, kk) == 0
            k1, k2 = _val(kk[0:4]), _val(kk[4:8])
            assert k1 == k % lam and k2 == k // lam
            assert bytes(a) == bytes(b), (k, pt[:4])
            assert bytes(b) == g1_lincomb([pt], [k.to_bytes(32, "big")])


def test_unconditional_lazy_subtraction_and_wide_products(field):
    """fe_sub_lazy4 (a - b + 2p, no condition: result in (0, 4p)), its zero test, and fe_mul_lazy with one
    factor below 4p and the other below 2p: congruent to the exact values and back below 2p."""
    rng = np.random.default_rng(13)
    edge = [0, 1, P - 1, P, P + 1, 2 * P - 1, (P - 1) // 2]
    rand = [int.from_bytes(rng.bytes(49), "big") % (2 * P) for _ in range(400)]
    av = rand + [e for e in edge for _ in edge]
    bv = rand[::-1] + [e for _ in edge for e in edge]
    n = len(av)
    a, b = _pack(av, 12), _pack(bv, 12)
    ptr = lambda x: x.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    d = np.zeros((n, 12), dtype=np.uint32)
    field.shim_fp_sub_lazy4_many(ptr(a), ptr(b), ptr(d), ctypes.c_size_t(n))
    dv = _unpack(d)
    assert all(0 < g < 4 * P for g in dv)
    assert dv == [x - y + 2 * P for x, y in zip(av, bv)]
    for g, x, y in list(zip(dv, av, bv))[:200] + list(zip(dv, av, bv))[-49:]:
        assert bool(field.shim_fp_is_zero_lazy4(_limbs(g, 12))) == ((x - y) % P == 0)
    for v in (P, 2 * P, 3 * P, P + 2 ** 32, 2 * P - 1):
        assert bool(field.shim_fp_is_zero_lazy4(_limbs(v, 12))) == (v % P == 0)
    Rinv = pow(pow(2, 384, P), -1, P)
    wide = [min(4 * P - 1, 2 * x + 1) for x in av]        # up to 4p - 1
    w = _pack(wide, 12)
    r = np.zeros((n, 12), dtype=np.uint32)
    for first, second, fv, sv in ((w, b, wide, bv), (b, w, bv, wide)):
        field.shim_fp_mul_lazy_many(ptr(first), ptr(second), ptr(r), ctypes.c_size_t(n))
        got = _unpack(r)
        assert all(g < 2 * P for g in got)
        assert [g % P for g in got] == [x * y * Rinv % P for x, y in zip(fv, sv)]


def test_trusted_setup_json_helper():
    """reference `TrustedSetup` (src/trusted_setup.rs:21-44, 138-153): the consensus-specs json form, with and
    without 0x, truncated to the preset's FIELD_ELEMENTS_PER_BLOB; malformed points are rejected."""
    import json
    import kzg_rust_b200 as k
    g = golden()
    g1 = [g.g1_bytes[48 * i:48 * i + 48] for i in range(4096)]
    g2 = [g.g2_bytes[96 * i:96 * i + 96] for i in range(65)]
    doc = {"setup_G1": ["0x" + "00" * 48], "setup_G1_lagrange": ["0x" + p.hex() for p in g1], "setup_G2": [p.hex() for p in g2]}
    ts = k.TrustedSetup.from_json(json.dumps(doc))
    assert ts.g1_len() == 4096 and ts.g2_len() == 65
    assert ts.g1_points() == g1 and ts.g2_points() == g2
    assert k.TrustedSetup.from_json(ts.to_json()).g1_points() == g1
    small = k.TrustedSetup.from_json(json.dumps(doc), field_elements_per_blob=4)
    assert small.g1_len() == 4 and small.g1_points() == g1[:4] and small.g2_len() == 65
    bad = dict(doc, setup_G2=[g2[0].hex()[:-2]] + doc["setup_G2"][1:])
    with pytest.raises(k.InvalidTrustedSetup):
        k.TrustedSetup.from_json(json.dumps(bad))
    bad = dict(doc, setup_G1_lagrange=["0xzz" + "00" * 47] + doc["setup_G1_lagrange"][1:])
    with pytest.raises(k.InvalidTrustedSetup):
        k.TrustedSetup.from_json(json.dumps(bad))
    with pytest.raises(k.InvalidTrustedSetup):
        k.TrustedSetup.from_json("{}")


def test_host_sha256_dispatch_vs_hashlib():
    """compute_r_powers' sequential hash runs on the host (csrc/host_sha256.cpp): the SHA-NI path where the CPU has it
    and the portable compression function, fed in odd-sized pieces, against hashlib."""
    import hashlib
    import kzg_rust_b200 as k
    L = k.load_library()
    L.kzg_b200_host_sha256.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p]
    rng = np.random.default_rng(8)
    for n in [0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 1000, 32 + 160 * 333]:
        d = rng.bytes(n)
        for portable in (0, 1):
            out = ctypes.create_string_buffer(32)
            L.kzg_b200_host_sha256(d, n, portable, out)
            assert out.raw == hashlib.sha256(d).digest(), (n, portable)


def test_paired_multiplication_vs_python(field):
    """fe_mul2 (bigint.cuh): two independent Montgomery products with interleaved carry chains, outputs aliasing inputs."""
    rng = np.random.default_rng(12)
    Rp_inv = pow(pow(2, 384, P), -1, P)
    edge = [0, 1, P - 1, P - 2, (P - 1) // 2, 2 ** 380]
    vals = edge + [int.from_bytes(rng.bytes(48), "big") % P for _ in range(120)]
    r1, r2 = (ctypes.c_uint32 * 12)(), (ctypes.c_uint32 * 12)()
    for i in range(0, len(vals) - 3):
        a1, b1, a2, b2 = vals[i], vals[i + 1], vals[i + 2], vals[i + 3]
        field.shim_fp_mul2(_limbs(a1, 12), _limbs(b1, 12), _limbs(a2, 12), _limbs(b2, 12), r1, r2)
        assert _val(r1) == a1 * b1 * Rp_inv % P and _val(r2) == a2 * b2 * Rp_inv % P


def test_host_verify_finish_on_records_built_from_the_public_tau():
    """kzg_b200_verify_finish (the host end of batch verification, reference src/kzg.rs:618-625): partial records
    (A, B, s) with B - [s]G1 = [tau]A must be accepted, anything else rejected.  tau = 1337 for the bundled testing setup, so
    the records can be made with the oracle's G1 arithmetic alone: A = [a]G1, B = [1337 a + s]G1, split over shards."""
    import kzg_rust_b200 as k
    from oracle import binding as ob
    L = k.load_library()
    L.kzg_b200_host_verify_finish_with_tau.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]
    g = golden()
    tau_g2 = g.g2_bytes[96:192]
    gen = bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")

    def point(scalar):  # [scalar]G1 as the 96-byte uncompressed record of the partial format
        c = ob.g1_lincomb([gen], [(scalar % R).to_bytes(32, "big")])
        if c[0] & 0x40:
            return bytes([0x40]) + bytes(95)
        x = int.from_bytes(bytes([c[0] & 0x1f]) + c[1:], "big")
        y = pow((x ** 3 + 4) % P, (P + 1) // 4, P)
        if (y > (P - 1) // 2) != bool(c[0] & 0x20):
            y = P - y
        return x.to_bytes(48, "big") + y.to_bytes(48, "big")

    rng = np.random.default_rng(77)
    for shards in (1, 3):
        recs, good = b"", True
        for _ in range(shards):
            a = int.from_bytes(rng.bytes(32), "big") % R
            s_ = int.from_bytes(rng.bytes(32), "big") % R
            recs += point(a) + point(1337 * a + s_) + s_.to_bytes(32, "big")
        ok = ctypes.c_int(-1)
        assert L.kzg_b200_host_verify_finish_with_tau(tau_g2, recs, shards, ctypes.byref(ok)) == 0 and ok.value == 1
        bad = bytearray(recs)
        bad[192 + 31] ^= 1  # another s
        assert L.kzg_b200_host_verify_finish_with_tau(tau_g2, bytes(bad), shards, ctypes.byref(ok)) == 0 and ok.value == 0
    # empty sums and s = 0: e(inf, .) == e(inf, .)
    rec = bytes([0x40]) + bytes(95) + bytes([0x40]) + bytes(95) + bytes(32)
    ok = ctypes.c_int(-1)
    assert L.kzg_b200_host_verify_finish_with_tau(tau_g2, rec, 1, ctypes.byref(ok)) == 0 and ok.value == 1
    # a scalar that is not canonical is malformed
    rec = bytes([0x40]) + bytes(95) + bytes([0x40]) + bytes(95) + b"\xff" * 32
    assert L.kzg_b200_host_verify_finish_with_tau(tau_g2, rec, 1, ctypes.byref(ok)) != 0


@pytest.mark.parametrize("n,force_c", [(1, 0), (3, 0), (6, 0), (6, 3), (13, 2), (40, 0), (40, 5), (9, 12)])
def test_bucket_method_of_phase_b_vs_ladders(msm, n, force_c):
    """pippenger.cuh walked on the CPU (GLV halves, Booth digits, counting sort, bucket sums, bit-position rows,
    weights) against one plain 255-step ladder per term, and against Python integers for the scalar sum: every
    window width, points at infinity, repeated points, z = 0, r = 1 and r^i z_i that make digits cancel."""
    g = golden()
    rng = np.random.default_rng(100 + n + force_c)
    inf = b"\xc0" + bytes(47)
    pts = [g.g1_bytes[48 * i:48 * i + 48] for i in rng.integers(0, 4096, size=2 * n)]
    cms, prs = pts[:n], pts[n:]
    if n >= 3:
        cms[1] = inf
        prs[2] = inf
        prs[0] = cms[0]           # the same point in two lanes
    if n >= 6:
        cms[4] = cms[3]           # equal points inside one lane: a bucket meets P + P
    zs = [int.from_bytes(rng.bytes(32), "big") % R for _ in range(n)]
    ys = [int.from_bytes(rng.bytes(32), "big") % R for _ in range(n)]
    zs[0] = 0
    if n >= 6:
        zs[3], zs[5] = 1, R - 1
    zy = b"".join(z.to_bytes(32, "big") + y.to_bytes(32, "big") for z, y in zip(zs, ys))
    for r, first in ((int.from_bytes(rng.bytes(32), "big") % R, 0), (1, 0), (R - 1, 5), (int.from_bytes(rng.bytes(32), "big") % R, 12345)):
        a, b = (ctypes.c_uint8 * 128)(), (ctypes.c_uint8 * 128)()
        c_used = ctypes.c_int(0)
        rc = msm.shim_verify_sums(b"".join(cms), b"".join(prs), zy, r.to_bytes(32, "big"), ctypes.c_uint64(first), n, force_c, a, b,
                                  ctypes.byref(c_used))
        assert rc == 0
        assert bytes(a) == bytes(b), (n, force_c, c_used.value, r, first)
        s = sum(pow(r, first + i, R) * ys[i] for i in range(n)) % R
        assert int.from_bytes(bytes(a)[96:128], "little") == s
        if force_c:
            assert c_used.value == force_c


def test_divstep_inversion_vs_python(field):
    """fe_inv_binary (csrc/bigint.cuh, the inversion of fp_inv / fr_inv) and fe_inv_safegcd (tools/experiments/safegcd.cuh:
    Bernstein-Yang divsteps, 30 at a time -- measured on the GPU and not kept) against Python's pow(a, -1, m), on edge values (1, 2, m - 1, powers of two, runs of ones,
    values just below the modulus), Fp inputs in [p, 2p) (the lazy residues the MSM's shared inversion hands over) and
    random values."""
    rng = np.random.default_rng(17)
    for mod, n, fn, fn_bin in ((P, 12, field.shim_fp_inv_safegcd, field.shim_fp_inv_binary),
                               (R, 8, field.shim_fr_inv_safegcd, field.shim_fr_inv_binary)):
        bits = 32 * n
        mont = pow(2, bits, mod)
        vals = [1, 2, 3, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, 2 ** 30, 2 ** 30 - 1, 2 ** 60 + 1, 2 ** (mod.bit_length() - 1),
                2 ** (mod.bit_length() - 1) - 1, 2 ** 200 - 1, (1 << 250) - (1 << 100), 0x5555555555555555555555555555555555555555555555555555555555555555 % mod,
                mod - 2 ** 64, mod - 2 ** 30]
        vals += [int.from_bytes(rng.bytes(n * 4), "big") % mod for _ in range(600)]
        vals += [pow(3, k, mod) for k in range(1, 60)]
        out, out2 = (ctypes.c_uint32 * n)(), (ctypes.c_uint32 * n)()
        for a in vals:
            if a % mod == 0:
                continue
            raw = a  # the integer handed over: a "Montgomery form" of a / 2^bits
            fn(_limbs(raw, n), out)
            want = pow(raw * pow(mont, -1, mod) % mod, -1, mod) * mont % mod
            assert _val(out) == want, (hex(mod)[:10], hex(a))
            fn_bin(_limbs(raw, n), out2)
            assert _val(out2) == want
        if n == 12:  # lazy residues: a + p < 2p
            for a in vals[:40]:
                raw = a + mod
                fn(_limbs(raw, n), out)
                assert _val(out) == pow(a * pow(mont, -1, mod) % mod, -1, mod) * mont % mod


def test_glv_split_is_exact_division(msm):
    """glv_split (csrc/g1.cuh: Barrett division by lambda, two fix-up subtractions at most) equals divmod on random scalars
    and on the values where an estimate of the quotient is most likely to be off (multiples of lambda and their neighbours,
    all-ones words, the largest scalars)."""
    lam = 0xd201000000010000 ** 2 - 1
    rng = np.random.default_rng(33)
    ks = [0, 1, lam - 1, lam, lam + 1, R - 1, R - 2, 2 ** 255 - 1, 2 ** 254, 2 ** 128, 2 ** 128 - 1, 2 ** 160 - 1, 2 ** 224 - 1]
    for m in (1, 2, 3, 2 ** 32 - 1, 2 ** 32, 2 ** 64 - 1, 2 ** 96, 2 ** 126, 2 ** 127 - 1, (R - 1) // lam):
        ks += [m * lam - 1, m * lam, m * lam + 1]
    ks += [int.from_bytes(rng.bytes(32), "big") >> 1 for _ in range(20000)]
    ks += [(int.from_bytes(rng.bytes(16), "big") * lam + d) % 2 ** 255 for d in (0, 1, lam - 1) for _ in range(2000)]
    out = (ctypes.c_uint32 * 8)()
    for k in ks:
        msm.shim_glv_split(_limbs(k, 8), out)
        assert (_val(out[0:4]), _val(out[4:8])) == (k % lam, k // lam), hex(k)

"""-m gpu: blob_to_kzg_commitment through the C ABI (CUDA path) against the reference's own
vectors (tests/golden) and against the CPU oracle on seeded random blobs.  Bit-exact."""
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


@pytest.mark.parametrize("width", VECTOR_WIDTHS, ids=VECTOR_WIDTH_IDS)
@pytest.mark.parametrize("case", G.by_fn("blob_to_kzg_commitment"), ids=[c["name"] for c in G.by_fn("blob_to_kzg_commitment")])
def test_reference_vectors(case, width):
    """reference src/lib.rs:30-52."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    try:
        blob = k.Blob.from_bytes(G.get_bytes(case["input"]["blob"]))
    except (k.Error, ValueError):
        assert case["output"] is None
        return
    try:
        out = k.Kzg.blob_to_kzg_commitment(blob, s)
    except k.Error:
        assert case["output"] is None
        return
    assert "0x" + out.to_bytes().hex() == case["output"]


@pytest.mark.parametrize("width", VECTOR_WIDTHS, ids=VECTOR_WIDTH_IDS)
def test_vectors_as_one_batch(width):
    """All well-sized vector blobs in one batched call: per-blob status, no cross-talk."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    cases = [c for c in G.by_fn("blob_to_kzg_commitment") if len(G.get_bytes(c["input"]["blob"])) == 131072]
    blobs = b"".join(G.get_bytes(c["input"]["blob"]) for c in cases)
    out, status = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
    for i, c in enumerate(cases):
        if c["output"] is None:
            assert status[i] == 1
        else:
            assert status[i] == 0 and "0x" + out[i].tobytes().hex() == c["output"]


@pytest.mark.parametrize("comb_width", [5, 8, 11])
def test_random_blobs_vs_oracle(comb_width):
    k = _kzg()
    s = gpu_settings("mainnet", comb_width)
    blobs = synthetic_blobs(24, seed=0xB200 + comb_width)
    # sprinkle edge values: 0, 1, r-1 and a duplicate blob
    blobs[1, :32] = 0
    blobs[2, 32:64] = np.frombuffer((R - 1).to_bytes(32, "big"), dtype=np.uint8)
    blobs[3] = blobs[2]
    out, status = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
    exp, est = oracle_settings("mainnet").blob_to_kzg_commitment_many(blobs, nthreads=os.cpu_count() or 1)
    assert not status.any() and not est.any()
    assert np.array_equal(out, exp)


def test_minimal_preset_vs_oracle():
    """kzg_minimal (FIELD_ELEMENTS_PER_BLOB = 4): parity unpinned by the reference, oracle only."""
    k = _kzg()
    s = gpu_settings("minimal", 6)
    assert s.field_elements_per_blob == 4
    vals = [[0, 0, 0, 0], [R - 1] * 4, [1, 0, 0, 0], [0, 0, 5, 0], [R, 0, 0, 0], [2 ** 256 - 1] * 4]
    rng = np.random.default_rng(5)
    for _ in range(58):
        vals.append([int.from_bytes(rng.bytes(32), "big") % R for _ in range(4)])
    blobs = b"".join(b"".join(v.to_bytes(32, "big") for v in row) for row in vals)
    out, status = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
    o = oracle_settings("minimal")
    from oracle.binding import OracleError
    for i in range(len(vals)):
        blob = blobs[128 * i:128 * i + 128]
        try:
            exp = o.blob_to_kzg_commitment(blob)
        except OracleError:
            assert status[i] == 1
            continue
        assert status[i] == 0 and out[i].tobytes() == exp, i


def test_large_batch_spans_chunks():
    """More blobs than one workspace chunk; every commitment checked by linearity-free identity
    C == [p(tau)]G1 is too slow in Python, so check against the oracle on a sample and check
    that identical blobs at different batch positions give identical commitments."""
    k = _kzg()
    os.environ["KZG_B200_CHUNK"] = "96"
    try:
        g = golden()
        s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 7)
    finally:
        del os.environ["KZG_B200_CHUNK"]
    base = synthetic_blobs(16, seed=77)
    blobs = np.concatenate([base] * 16, axis=0)  # 256 blobs, 3 chunks of 96
    out, status = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
    assert not status.any()
    exp, _ = oracle_settings("mainnet").blob_to_kzg_commitment_many(base, nthreads=os.cpu_count() or 1)
    for rep in range(16):
        assert np.array_equal(out[16 * rep:16 * rep + 16], exp)
    s.close()


@pytest.mark.parametrize("env", [{"KZG_B200_ADD_BLOCKS": "2"}, {"KZG_B200_ADD_BLOCKS": "4"}, {"KZG_B200_BATCH_K": "3"},
                                 {"KZG_B200_LANES": "1", "KZG_B200_GRID_BLOCKS": "1"}],
                         ids=["2-blocks-per-sm", "4-blocks-per-sm", "short-batches", "one-lane"])
def test_kernel_configurations_vs_oracle(env):
    """The launch shapes of the addition kernel (register budget, batch length, lanes) are knobs: same bytes as the
    oracle for each, special cases included, several chunks."""
    k = _kzg()
    env = dict(env, KZG_B200_CHUNK="40")
    os.environ.update(env)
    try:
        g = golden()
        s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 9)
    finally:
        for key in env:
            del os.environ[key]
    blobs = synthetic_blobs(100, seed=0xD1)
    blobs[3, :] = 0                      # every scalar is recoded as r: the sums cancel to infinity
    blobs[4] = 0
    blobs[4, 31::32] = 1                 # the constant polynomial 1: the same table entry in every group
    blobs[5] = 0
    blobs[5, 31::32] = 2                 # even scalars: recoded as r - 2 with flipped signs
    blobs[50] = blobs[49]
    out, status = k.Kzg.blob_to_kzg_commitment_batch(blobs, s)
    assert not status.any()
    exp, est = oracle_settings("mainnet").blob_to_kzg_commitment_many(blobs, nthreads=os.cpu_count() or 1)
    assert not est.any() and np.array_equal(out, exp)
    s.close()


def test_load_trusted_setup_file(tmp_path):
    """reference load_trusted_setup_file (src/kzg.rs:906-979): the text layout of trusted_setup.txt through
    kzg_b200_ctx_create_from_file, then one vector; malformed files are rejected with the reference's error kinds."""
    k = _kzg()
    path = tmp_path / "trusted_setup.txt"
    G.write_setup_text(str(path))
    s = k.Kzg.load_trusted_setup_file(str(path), 0, 6)
    assert s.field_elements_per_blob == 4096 and s.comb_width == 6
    case = next(c for c in G.by_fn("blob_to_kzg_commitment") if c["output"] is not None)
    out = k.Kzg.blob_to_kzg_commitment(k.Blob.from_bytes(G.get_bytes(case["input"]["blob"])), s)
    assert "0x" + out.to_bytes().hex() == case["output"]
    s.close()
    text = path.read_text().split("\n")
    bad = tmp_path / "bad_count.txt"
    bad.write_text("\n".join(["4095"] + text[1:]))
    with pytest.raises(k.InvalidTrustedSetup):
        k.Kzg.load_trusted_setup_file(str(bad), 0, 6)
    bad = tmp_path / "bad_hex.txt"
    bad.write_text("\n".join(text[:2] + ["zz" + text[2][2:]] + text[3:]))
    with pytest.raises(k.InvalidHexFormat):
        k.Kzg.load_trusted_setup_file(str(bad), 0, 6)
    bad = tmp_path / "short.txt"
    bad.write_text("\n".join(text[:100]))
    with pytest.raises(k.InvalidTrustedSetup):
        k.Kzg.load_trusted_setup_file(str(bad), 0, 6)
    with pytest.raises(k.InvalidTrustedSetup):
        k.Kzg.load_trusted_setup_file(str(tmp_path / "missing.txt"), 0, 6)
    bad = tmp_path / "not_a_point.txt"
    bad.write_text("\n".join(text[:2] + ["8" + "0" * 94 + "07"] + text[3:]))  # x = 7: x^3 + 4 is not a square
    with pytest.raises(k.Error):
        k.Kzg.load_trusted_setup_file(str(bad), 0, 6)


def test_every_commitment_of_a_large_batch_by_the_tau_identity():
    """C == [p(tau)] G1 on EVERY blob (the bundled setup is the public testing setup,
    tau = 1337) at the benchmarked configuration, device-resident, several chunks; a tampered commitment is flagged."""
    import ctypes
    import torch
    k = _kzg()
    L = k.load_library()
    s = gpu_settings("mainnet", 0)
    n = 2 * s.chunk_blobs + 77
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device=dev, generator=gen)
    blobs[:, :, 0] = 0
    blobs[5] = 0                                     # zero polynomial -> infinity
    blobs[6] = 0
    blobs[6, :, 31] = 9                              # constant polynomial
    out = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    assert L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs.data_ptr(), n, out.data_ptr(), st.data_ptr()) == 0
    assert L.kzg_b200_synchronize(s._h) == 0 and not bool(st.any().item())
    assert out[5].cpu().numpy().tobytes() == bytes([0xc0]) + bytes(47)
    ok = torch.zeros(n, dtype=torch.int32, device=dev)
    tau = (1337).to_bytes(32, "big")
    assert L.kzg_b200_debug_check_tau_identity(s._h, blobs.data_ptr(), out.data_ptr(), n, tau, ok.data_ptr()) == 0
    assert int(ok.sum().item()) == n
    out[n // 2, 47] ^= 1
    out[3] = out[4]
    assert L.kzg_b200_debug_check_tau_identity(s._h, blobs.data_ptr(), out.data_ptr(), n, tau, ok.data_ptr()) == 0
    bad = set(torch.nonzero(ok == 0).flatten().cpu().tolist())
    assert bad == {3, n // 2}
    # and the sample the oracle can afford
    exp, est = oracle_settings("mainnet").blob_to_kzg_commitment_many(blobs[8:8 + 32].reshape(32, -1).cpu().numpy(), nthreads=os.cpu_count() or 1)
    out[3] = 0
    got = out[8:8 + 32].cpu().numpy()
    assert not est.any() and np.array_equal(got, exp)


def test_trusted_setup_json_to_settings():
    """reference `TrustedSetup` (src/trusted_setup.rs): json -> g1_points / g2_points -> load_trusted_setup."""
    import json
    k = _kzg()
    doc = {"setup_G1_lagrange": ["0x" + G.g1_bytes[48 * i:48 * i + 48].hex() for i in range(4096)],
           "setup_G2": [G.g2_bytes[96 * i:96 * i + 96].hex() for i in range(65)]}
    ts = k.TrustedSetup.from_json(json.dumps(doc))
    s = k.Kzg.load_trusted_setup(ts.g1_points(), ts.g2_points(), 0, 5)
    case = next(c for c in G.by_fn("blob_to_kzg_commitment") if c["output"] is not None)
    out = k.Kzg.blob_to_kzg_commitment(k.Blob.from_bytes(G.get_bytes(case["input"]["blob"])), s)
    assert "0x" + out.to_bytes().hex() == case["output"]
    s.close()


def test_automatic_comb_width_follows_the_memory_rule():
    """comb_width = 0: the widest comb whose table takes at most half of the free memory and leaves room for a workspace
    and the reserve (include/kzg_b200.h); on a small or shared GPU the width steps down, it never collapses to a tiny
    window, and the bytes stay the same."""
    k = _kzg()
    case = next(c for c in G.by_fn("blob_to_kzg_commitment") if c["output"] is not None)
    widths = []
    import torch
    for budget_gb in (160, 60, 24, 9):
        # other contexts of this test session hold memory too: the rule plans with min(budget, what is free now)
        eff = min(budget_gb * 2 ** 30, torch.cuda.mem_get_info(0)[0])
        os.environ["KZG_B200_MEM_BUDGET_GB"] = str(budget_gb)
        try:
            s = k.KzgSettings.load_trusted_setup(G.g1_bytes, G.g2_bytes, 0, 0)
        finally:
            del os.environ["KZG_B200_MEM_BUDGET_GB"]
        widths.append(s.comb_width)
        assert s.table_bytes <= budget_gb * 2 ** 30 // 2 and s.chunk_blobs >= 1, (budget_gb, s.comb_width, s.table_bytes)
        # one step wider would not have fitted in half of the budget
        wider = -(-4096 // (s.comb_width + 1)) * 2 ** s.comb_width * 96
        assert s.comb_width == 24 or wider > eff // 2 or eff < 16 * 2 ** 30, (budget_gb, eff, s.comb_width)
        out = k.Kzg.blob_to_kzg_commitment(k.Blob.from_bytes(G.get_bytes(case["input"]["blob"])), s)
        assert "0x" + out.to_bytes().hex() == case["output"]
        s.close()
    # 9 GB cannot hold a table beside the fixed scratch of the addition kernel: the rule stops at its floor (8), not at 2
    assert widths == sorted(widths, reverse=True) and widths[2] >= 16 and widths[-1] >= 8, widths


@pytest.mark.parametrize("preset", ["mainnet", "minimal"])
def test_small_batch_form_equals_the_batched_affine_tree(preset):
    """Batches of up to 64 blobs sum the table entries with one warp per sum in Jacobian coordinates (csrc/msm.cu:
    k_comb_rows_warp) instead of the batched affine levels: the same 48 bytes, for commitments and proofs, including blobs
    whose sums cancel (all zero, all r - 1, two equal rows) and the vectors of the reference."""
    k = _kzg()
    s = gpu_settings(preset, 8)
    n = 4096 if preset == "mainnet" else 4
    blobs = synthetic_blobs(64, n=n, seed=0x5A11).copy()
    blobs[3] = 0
    blobs[7] = np.tile(np.frombuffer((R - 1).to_bytes(32, "big"), dtype=np.uint8), n)
    blobs[9] = np.tile(np.frombuffer((1).to_bytes(32, "big"), dtype=np.uint8), n)
    blobs[11] = blobs[10]
    if preset == "mainnet":
        good = [c for c in G.by_fn("blob_to_kzg_commitment") if c["output"] is not None]
        for i, c in enumerate(good):
            blobs[20 + i] = np.frombuffer(G.get_bytes(c["input"]["blob"]), dtype=np.uint8)
    outs = {}
    # (threshold of the small form, latency comb on / off): the latency comb (64 sums over the virtual points 2^(64 t) G_i,
    # csrc/internal.h) is the default of small mainnet calls, the 255-sum warp form its fallback, 0 = the batched affine tree
    for small_max, lat in ((64, 1), (64, 0), (0, 1)):
        os.environ["KZG_B200_MSM_SMALL_MAX"] = str(small_max)
        os.environ["KZG_B200_LATENCY_TABLE"] = str(lat)
        try:
            res = []
            for m in (1, 5, 64):
                cms, st = k.Kzg.blob_to_kzg_commitment_batch(blobs[:m], s)
                assert not st.any()
                prs, st = k.Kzg.compute_blob_kzg_proof_batch(blobs[:m], cms, s)
                assert not st.any()
                res.append((cms.tobytes(), prs.tobytes()))
            outs[(small_max, lat)] = res
        finally:
            del os.environ["KZG_B200_MSM_SMALL_MAX"]
            del os.environ["KZG_B200_LATENCY_TABLE"]
    assert outs[(64, 1)] == outs[(0, 1)] and outs[(64, 0)] == outs[(0, 1)]
    if preset == "mainnet":
        cms = np.frombuffer(outs[(64, 1)][2][0], dtype=np.uint8).reshape(64, 48)
        for i, c in enumerate(good):
            assert "0x" + cms[20 + i].tobytes().hex() == c["output"]


def test_piece_plan_of_host_calls_gives_the_same_bytes():
    """Host-buffer calls that fit one chunk are cut in two halves from 512 blobs on, larger ones start with a 1,024-blob
    chunk (csrc/kzg_b200.cu: msm_piece_plan): commitments, blob proofs and proofs at given points equal those of the plain
    chunking, on the benchmarked configuration (4,096-blob chunks)."""
    k = _kzg()
    s = gpu_settings("mainnet", 0)
    n = 5000
    blobs = synthetic_blobs(n, seed=0x91EC)
    zs = synthetic_blobs(1, n=n, seed=0x91ED).reshape(n, 32)
    outs = {}
    for plain in (0, 1):
        os.environ["KZG_B200_PLAIN_PIECES"] = str(plain)
        try:
            res = []
            for m in (600, 1025, n):
                cms, st = k.Kzg.blob_to_kzg_commitment_batch(blobs[:m], s)
                assert not st.any()
                prs, st = k.Kzg.compute_blob_kzg_proof_batch(blobs[:m], cms, s)
                assert not st.any()
                res.append((cms.tobytes(), prs.tobytes()))
            pz, ys, st = k.Kzg.compute_kzg_proof_batch(blobs[:700], zs[:700], s)
            assert not st.any()
            res.append((pz.tobytes(), ys.tobytes()))
            outs[plain] = res
        finally:
            del os.environ["KZG_B200_PLAIN_PIECES"]
    assert outs[0] == outs[1]


@pytest.mark.parametrize("width", VECTOR_WIDTHS, ids=VECTOR_WIDTH_IDS)
def test_jacobian_tail_of_mid_size_batches_gives_the_same_bytes(width):
    """Single-chunk calls of 17 .. 256 blobs end the addition tree at 12 rows and add the rest per sum in Jacobian coordinates
    (csrc/msm.cu: k_tail_rows_jac): commitments and proofs equal those of the full affine tree, including blobs whose sums
    cancel, and the first ones equal the oracle's."""
    k = _kzg()
    s = gpu_settings("mainnet", width)
    o = oracle_settings("mainnet")
    sizes = (17, 64, 128) if width != 0 else (40, 200, 256)
    blobs = synthetic_blobs(max(sizes), seed=0x7A11).copy()
    blobs[2] = 0
    blobs[5] = np.tile(np.frombuffer((R - 1).to_bytes(32, "big"), dtype=np.uint8), 4096)
    blobs[8] = blobs[7]
    outs = {}
    for tail in (1, 0):
        os.environ["KZG_B200_JAC_TAIL"] = str(tail)
        try:
            res = []
            for m in sizes:
                cms, st = k.Kzg.blob_to_kzg_commitment_batch(blobs[:m], s)
                assert not st.any()
                prs, st = k.Kzg.compute_blob_kzg_proof_batch(blobs[:m], cms, s)
                assert not st.any()
                res.append((cms.tobytes(), prs.tobytes()))
            outs[tail] = res
        finally:
            del os.environ["KZG_B200_JAC_TAIL"]
    assert outs[1] == outs[0]
    exp, est = o.blob_to_kzg_commitment_many(blobs[:12], nthreads=8)
    assert not est.any() and outs[1][0][0][:12 * 48] == exp.tobytes()

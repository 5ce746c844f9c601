// CPU walk-through of the product's commitment pipeline (host code path of the same
// headers the CUDA kernels are built from): decode setup -> bit-reverse -> window bases
// -> table levels -> digits -> gather level -> tree levels -> compress.  "Threads" are
// iterated sequentially.  Small presets only (n = 4); used by tests/test_host_logic.py.
#include <vector>
#include "../../kzg_rust_b200/csrc/msm.cuh"
#include "../../kzg_rust_b200/csrc/blobpath.cuh"
using namespace kzg;

template <class Policy>
static void run_level(const Policy &pol, uint64_t total, int T, int k) {
    std::vector<fp_t> scratch((size_t)T * k);
    for (int tid = 0; tid < T; tid++) batch_add_thread(pol, total, scratch.data(), k, (uint64_t)T, (uint64_t)tid);
}
static uint32_t bitrev(uint32_t v, int n) { uint32_t r = 0; for (int o = n; o > 1; o >>= 1) { r = (r << 1) | (v & 1); v >>= 1; } return r; }

static int build_table(std::vector<g1_affine_t> &table, const uint8_t *g1_bytes, int n, int c, int T, int k) {
    const uint32_t D = 1u << (c - 1);
    table.resize((size_t)n * D);
    for (int i = 0; i < n; i++) {
        g1_affine_t p;
        if (g1_decode_thread(p, g1_bytes + 48 * bitrev(i, n), false)) return KZG_BADARGS;
        table[(size_t)i * D] = p;
    }
    for (int L = 0; L + 1 < c; L++) {
        TableLevelPolicy pol{table.data(), D, (uint32_t)L};
        run_level(pol, (uint64_t)n << L, T, k);
    }
    return 0;
}

extern "C" int shim_commit(const uint8_t *g1_bytes, int n, int c, const uint8_t *blobs, int B, uint8_t *out,
                           int *status, int T, int k) {
    const int W = msm_num_windows(c);
    const uint32_t D = 1u << (c - 1);
    std::vector<g1_affine_t> table;
    if (build_table(table, g1_bytes, n, c, T, k)) return KZG_BADARGS;
    std::vector<int32_t> digits((size_t)B * W * n);
    for (int b = 0; b < B; b++) status[b] = 0;
    for (uint64_t e = 0; e < (uint64_t)B * n; e++) blob_digits_thread(blobs, e, n, c, W, digits.data(), status);
    // point-major copy of the digits (k_transpose_digits on the device)
    std::vector<int32_t> digits_t((size_t)B * W * n);
    for (int b = 0; b < B; b++)
        for (int j = 0; j < W; j++)
            for (int i = 0; i < n; i++) digits_t[((size_t)i * W + j) * B + b] = digits[((size_t)b * W + j) * n + i];
    const uint64_t R = (uint64_t)B * W;
    const FastDiv fd = FastDiv::make((uint32_t)R);
    uint32_t rows = n / 2;
    std::vector<g1_affine_t> a((size_t)R * rows), b2((size_t)R * rows);
    GatherPolicy gp{table.data(), digits_t.data(), a.data(), fd, D};
    run_level(gp, R * rows, T, k);
    g1_affine_t *in = a.data(), *o = b2.data();
    while (rows > 1) {
        rows /= 2;
        PairPolicy tp{in, o, fd};
        run_level(tp, R * rows, T, k);
        std::swap(in, o);
    }
    for (int b = 0; b < B; b++) {
        g1_affine_t p;
        horner_thread(p, in + b, (size_t)B, c, W);
        g1a_compress(out + 48 * b, p);
    }
    return 0;
}

// The precomputed table alone (same layout as the device table), for debugging / tests.
extern "C" int shim_table(const uint8_t *g1_bytes, int n, int c, g1_affine_t *table_out, int T, int k) {
    std::vector<g1_affine_t> table;
    if (build_table(table, g1_bytes, n, c, T, k)) return KZG_BADARGS;
    memcpy(table_out, table.data(), table.size() * sizeof(g1_affine_t));
    return 0;
}
// signed-digit recoding of one canonical scalar (8 little-endian words) for a given window
extern "C" int shim_recode(const uint32_t *scalar, int c, int32_t *digits_out) {
    fr_t s;
    memcpy(s.l, scalar, 32);
    int W = msm_num_windows(c);
    recode_signed(s, c, W, digits_out, 1);
    return W;
}
// [k]P by the plain ladder and by the GLV ladder (g1.cuh), both compressed; k = 8 little-endian words
extern "C" int shim_g1_mul_both(const uint8_t *point48, const uint32_t *k, uint8_t *out_plain, uint8_t *out_glv, uint32_t *k1k2) {
    g1_affine_t p;
    if (!g1a_uncompress(p, point48)) return KZG_BADARGS;
    g1_jac_t a, b;
    g1j_mul(a, p, k, 256);
    g1j_mul_glv(b, p, k);
    g1_affine_t aa, bb;
    g1j_to_affine(aa, a);
    g1j_to_affine(bb, b);
    g1a_compress(out_plain, aa);
    g1a_compress(out_glv, bb);
    glv_split(k1k2, k1k2 + 4, k);
    return 0;
}

// CPU walk-through of the product's commitment pipeline (host code path of the same
// headers the CUDA kernels are built from): decode setup -> bit-reverse -> comb table levels
// -> sign words -> comb digits -> gather level -> tree levels -> Horner -> compress.  "Threads" are
// iterated sequentially.  Small point counts only; used by tests/test_host_logic.py.
#include <cstring>
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

struct Comb {
    int n, g, G, n_pad;
    uint64_t E;
    std::vector<g1_affine_t> bases, table;
};
// bitrev_order = 0 keeps the points in the order given (n need not be a power of two then)
static int build_comb(Comb &cb, const uint8_t *g1_bytes, int n, int g, int T, int k, int bitrev_order) {
    if (g > n) g = n;
    cb.n = n; cb.g = g; cb.G = (n + g - 1) / g; cb.n_pad = cb.G * g; cb.E = 1ull << (g - 1);
    cb.bases.resize(cb.n_pad);
    for (int i = 0; i < cb.n_pad; i++) {
        if (i < n) {
            if (g1_decode_thread(cb.bases[i], g1_bytes + 48 * (bitrev_order ? bitrev(i, n) : (uint32_t)i), false)) return KZG_BADARGS;
        } else {
            g1a_set_inf(cb.bases[i]);
        }
    }
    cb.table.resize((size_t)cb.G * cb.E);
    for (int q = 0; q < cb.G; q++) cb.table[(size_t)q << (g - 1)] = cb.bases[(size_t)q * g];
    for (int m = 1; m < g; m++) {
        const uint64_t total = (uint64_t)cb.G << (m - 1);
        run_level(CombLevelPolicy{cb.table.data(), cb.bases.data(), (uint32_t)g, (uint32_t)m, 1u}, total, T, k);
        run_level(CombLevelPolicy{cb.table.data(), cb.bases.data(), (uint32_t)g, (uint32_t)m, 0u}, total, T, k);
    }
    return 0;
}
// sums over the comb of B scalar vectors given as sign words [b][8][n_pad] -> compressed points
static void comb_msm(const Comb &cb, const std::vector<uint32_t> &sw, int B, uint8_t *out, int T, int k) {
    const int W = KZG_COMB_WINDOWS;
    std::vector<uint32_t> digits((size_t)cb.G * W * B);
    for (int b = 0; b < B; b++)
        for (int q = 0; q < cb.G; q++) comb_index_thread(sw.data(), (uint32_t)b, (uint32_t)q, cb.g, cb.n_pad, (uint32_t)B, digits.data());
    const uint64_t R = (uint64_t)B * W;
    const FastDiv fd = FastDiv::make((uint32_t)R);
    uint32_t rows = (uint32_t)cb.G;
    std::vector<g1_affine_t> a((size_t)R * ((rows + 1) / 2)), b2((size_t)R * ((rows + 1) / 2));
    if (rows / 2) run_level(GatherPolicy{cb.table.data(), digits.data(), a.data(), fd, (uint32_t)cb.E}, R * (rows / 2), T, k);
    if (rows & 1) {  // k_gather_copy on the device
        for (uint64_t r = 0; r < R; r++) {
            uint32_t d = digits[(uint64_t)(rows - 1) * R + r];
            g1_affine_t p = cb.table[(uint64_t)(rows - 1) * cb.E + (d & 0x7fffffffu)];
            if ((d >> 31) && !g1a_is_inf(p)) fe_neg_lazy(p.y, p.y);
            a[(uint64_t)(rows / 2) * R + r] = p;
        }
    }
    rows = (rows + 1) / 2;
    g1_affine_t *in = a.data(), *o = b2.data();
    while (rows > 1) {
        run_level(PairPolicy{in, o, fd}, R * (rows / 2), T, k);
        if (rows & 1) memcpy(o + (uint64_t)(rows / 2) * R, in + (uint64_t)(rows - 1) * R, R * sizeof(g1_affine_t));
        rows = (rows + 1) / 2;
        std::swap(in, o);
    }
    for (int b = 0; b < B; b++) {
        g1_affine_t p;
        horner_thread(p, in + b, (size_t)B, 1, W);
        g1a_compress(out + 48 * b, p);
    }
}

// commitments of B blobs of n elements (n a power of two: the setup is bit-reversal permuted like src/kzg.rs:895-896)
extern "C" int shim_commit(const uint8_t *g1_bytes, int n, int g, const uint8_t *blobs, int B, uint8_t *out,
                           int *status, int T, int k) {
    Comb cb;
    if (build_comb(cb, g1_bytes, n, g, T, k, 1)) return KZG_BADARGS;
    std::vector<uint32_t> sw((size_t)B * 8 * cb.n_pad);
    for (int b = 0; b < B; b++) status[b] = 0;
    for (uint64_t e = 0; e < (uint64_t)B * cb.n_pad; e++) blob_sign_words_thread(blobs, e, n, cb.n_pad, sw.data(), status);
    comb_msm(cb, sw, B, out, T, k);
    return 0;
}
// the same MSM over ANY number of points in the order given, scalars as canonical limbs [b][n][8], against the plain
// ladder: out_comb / out_ladder = sum_i s_i P_i compressed (odd group counts, padding and every comb width get exercised)
extern "C" int shim_msm_vs_ladder(const uint8_t *g1_bytes, int n, int g, const uint32_t *scalars, int B, uint8_t *out_comb,
                                  uint8_t *out_ladder, int T, int k) {
    Comb cb;
    if (build_comb(cb, g1_bytes, n, g, T, k, 0)) return KZG_BADARGS;
    std::vector<uint32_t> sw((size_t)B * 8 * cb.n_pad);
    for (uint64_t e = 0; e < (uint64_t)B * cb.n_pad; e++)
        fr_sign_words_thread(reinterpret_cast<const fr_t *>(scalars), e, n, cb.n_pad, sw.data());
    comb_msm(cb, sw, B, out_comb, T, k);
    for (int b = 0; b < B; b++) {
        g1_jac_t acc;
        g1j_set_inf(acc);
        for (int i = 0; i < n; i++) {
            g1_jac_t t;
            g1j_mul(t, cb.bases[i], scalars + ((size_t)b * n + i) * 8, 255);
            g1j_add(acc, acc, t);
        }
        g1_affine_t p;
        g1j_to_affine(p, acc);
        g1a_compress(out_ladder + 48 * b, p);
    }
    return 0;
}
// The precomputed table alone (same layout as the device table), for debugging / tests.
extern "C" int shim_table(const uint8_t *g1_bytes, int n, int g, g1_affine_t *table_out, int T, int k) {
    Comb cb;
    if (build_comb(cb, g1_bytes, n, g, T, k, 1)) return KZG_BADARGS;
    memcpy(table_out, cb.table.data(), cb.table.size() * sizeof(g1_affine_t));
    return 0;
}
// the 255 signs of one canonical scalar (8 little-endian words in, 8 words out)
extern "C" void shim_sign_words(const uint32_t *scalar, uint32_t *out) {
    fr_t s;
    memcpy(s.l, scalar, 32);
    scalar_sign_words(out, s);
}
extern "C" uint32_t shim_comb_digit(uint32_t pattern, int g) { return comb_digit(pattern, g); }
extern "C" void shim_transpose32(uint32_t *a) { transpose32(a); }
// [k]P by the plain ladder and by the GLV ladder (g1.cuh), both compressed; k = 8 little-endian words
extern "C" void shim_glv_split(const uint32_t *k, uint32_t *k1k2) { glv_split(k1k2, k1k2 + 4, k); }
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

// Phase B of batch verification on the CPU, twice: the bucket method of pippenger.cuh walked thread by thread
// (scalars -> count -> offsets -> scatter -> bucket sums -> weighted rows -> total) and one GLV ladder per term.
// Points come compressed (48 B each); out_* = A (48 B compressed) || B (48 B) || sum r_i y_i (32 B little-endian limbs).
#include "../../kzg_rust_b200/csrc/pippenger.cuh"
extern "C" int shim_verify_sums(const uint8_t *commitments, const uint8_t *proofs, const uint8_t *zy, const uint8_t *r_be, uint64_t first,
                                int n, int force_c, uint8_t *out_pip, uint8_t *out_ladder, int *c_used) {
    std::vector<g1_affine_t> cp(n), pp(n);
    for (int i = 0; i < n; i++)
        if (!g1a_uncompress(cp[i], commitments + 48 * i) || !g1a_uncompress(pp[i], proofs + 48 * i)) return KZG_BADARGS;
    fr_t r;
    scalar_from_be32(r, r_be);
    const PipPlan pl = PipPlan::make((size_t)n, force_c);
    *c_used = pl.c;
    std::vector<uint32_t> halves(16 * (size_t)n), counts(pl.buckets, 0), offsets(pl.buckets + 1), cursor(pl.buckets), entries(pl.max_entries);
    std::vector<fr_t> sy(n);
    for (int i = 0; i < n; i++) pip_scalars_thread((uint32_t)i, zy, r, first, halves.data(), sy.data());
    for (int w = 0; w < pl.W; w++)
        for (int i = 0; i < n; i++) pip_digits_thread<false>((uint32_t)i, w, pl, halves.data(), counts.data(), nullptr);
    uint32_t run = 0;
    for (uint32_t b = 0; b < pl.buckets; b++) { offsets[b] = cursor[b] = run; run += counts[b]; }
    offsets[pl.buckets] = run;
    if (run > pl.max_entries) return KZG_INTERNAL;
    for (int w = 0; w < pl.W; w++)
        for (int i = 0; i < n; i++) pip_digits_thread<true>((uint32_t)i, w, pl, halves.data(), cursor.data(), entries.data());
    std::vector<g1_jac_t> buckets(pl.buckets);
    for (uint32_t b = 0; b < pl.buckets; b++) pip_bucket_thread(b, pl, offsets.data(), entries.data(), cp.data(), pp.data(), buckets.data());
    fr_t s_tot;
    fe_set_zero(s_tot);
    for (int i = 0; i < n; i++) fe_add(s_tot, s_tot, sy[i]);
    fe_from_mont(s_tot, s_tot);
    for (int o = 0; o < 2; o++) {
        g1_jac_t total;
        g1j_set_inf(total);
        for (uint32_t p = 0; p < pl.rows; p++) {
            const int w = (int)(p / pl.c), j = (int)(p % pl.c);
            g1_jac_t row;
            g1j_set_inf(row);
            for (uint32_t t = 0; t < pip_row_len(pl, o, w, j); t++) g1j_add(row, row, buckets[pip_row_slot(pl, o, w, j, t)]);
            pip_weight(row, p);
            g1j_add(total, total, row);
        }
        g1_affine_t a;
        g1j_to_affine(a, total);
        g1a_compress(out_pip + 48 * o, a);
    }
    memcpy(out_pip + 96, s_tot.l, 32);
    // the ladders
    fr_t rm, s2;
    fe_to_mont(rm, r);
    fe_set_zero(s2);
    g1_jac_t A, B;
    g1j_set_inf(A);
    g1j_set_inf(B);
    for (int i = 0; i < n; i++) {
        fr_t ri, y, z, k;
        fr_pow_u64(ri, rm, first + (uint64_t)i);
        scalar_from_be32(y, zy + 64 * (size_t)i + 32); fe_to_mont(y, y); fe_mul(y, ri, y); fe_add(s2, s2, y);
        scalar_from_be32(z, zy + 64 * (size_t)i); fe_to_mont(z, z); fe_mul(z, ri, z);
        g1_jac_t t;
        fe_from_mont(k, ri);
        g1j_mul(t, pp[i], k.l, 255); g1j_add(A, A, t);
        g1j_mul(t, cp[i], k.l, 255); g1j_add(B, B, t);
        fe_from_mont(k, z);
        g1j_mul(t, pp[i], k.l, 255); g1j_add(B, B, t);
    }
    g1_affine_t a;
    g1j_to_affine(a, A); g1a_compress(out_ladder, a);
    g1j_to_affine(a, B); g1a_compress(out_ladder + 48, a);
    fe_from_mont(s2, s2);
    memcpy(out_ladder + 96, s2.l, 32);
    return 0;
}

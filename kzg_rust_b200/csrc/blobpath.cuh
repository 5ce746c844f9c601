// blobpath.cuh -- per-thread bodies of the small kernels around the MSM: blob bytes ->
// scalars -> signed digits (reference `blob_to_polynomial` src/kzg.rs:282-291 and
// `bytes_to_bls_field` src/utils.rs:262-275), window bases for the table, point
// (de)compression.  Written as host+device functions so a CPU test can run them.
#pragma once
#include "g1.cuh"

namespace kzg {

enum { KZG_OK = 0, KZG_BADARGS = 1, KZG_INTERNAL = 2, KZG_INVALID_LENGTH = 3, KZG_INVALID_HEX = 4,
       KZG_INVALID_SETUP = 5, KZG_CUDA = 6 };

// 32 big-endian bytes -> canonical little-endian limbs
KZG_HD void scalar_from_be32(fr_t &r, const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    const uint4 *q = reinterpret_cast<const uint4 *>(p);  // blob elements are 32-byte aligned
    uint4 hi = q[0], lo = q[1];
    r.l[7] = __byte_perm(hi.x, 0, 0x0123); r.l[6] = __byte_perm(hi.y, 0, 0x0123);
    r.l[5] = __byte_perm(hi.z, 0, 0x0123); r.l[4] = __byte_perm(hi.w, 0, 0x0123);
    r.l[3] = __byte_perm(lo.x, 0, 0x0123); r.l[2] = __byte_perm(lo.y, 0, 0x0123);
    r.l[1] = __byte_perm(lo.z, 0, 0x0123); r.l[0] = __byte_perm(lo.w, 0, 0x0123);
#else
    for (int i = 0; i < 8; i++) {
        const uint8_t *q = p + 4 * (7 - i);
        r.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
#endif
}
KZG_HD void scalar_to_be32(uint8_t *p, const fr_t &canon) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint8_t *q = p + 4 * (7 - i);
        uint32_t v = canon.l[i];
        q[0] = (uint8_t)(v >> 24); q[1] = (uint8_t)(v >> 16); q[2] = (uint8_t)(v >> 8); q[3] = (uint8_t)v;
    }
}
// `hash_to_bls_field` (reference src/utils.rs:250-258): reduce a 256-bit value mod r.
// 2^256 < 3r, so at most two subtractions.
KZG_HD void scalar_reduce(fr_t &a) {
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = FrParams::mod(i);
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        if (!limbs_geq<8>(a.l, m)) break;
        uint32_t cc = 0;
        a.l[0] = sub_cc(a.l[0], m[0], cc);
#pragma unroll
        for (int i = 1; i < 8; i++) a.l[i] = subc_cc(a.l[i], m[i], cc);
    }
}

// ------------------------------------------------------------------ signed comb recoding of the scalars
// The MSM (msm.cuh) looks up precomputed signed sums of g setup points,
//     T[q][idx] = P_{q,0} + sum_{k=1..g-1} (bit_{k-1}(idx) ? + : -) P_{q,k},      P_{q,k} = G_{q g + k},
// one lookup per (group q, bit position j) -- so every scalar is first written with digits +-1 only:
//     s = sum_{j=0..254} e_j 2^j  (mod r),   e_j in {+1, -1}.
// For an odd s < 2^255 that is e_j = +1 iff bit j+1 of s is set (j < 254), e_254 = +1.  An even s
// (zero included) is replaced by r - s, which is odd and in [1, r], with every sign flipped.
// The 255 signs of a scalar are its "sign words" (bit j set = plus), 8 x 32 bits.
#define KZG_COMB_WINDOWS 255
#define KZG_COMB_MAX_WIDTH 24
KZG_HD void scalar_sign_words(uint32_t out[8], const fr_t &s) {
    const uint32_t flip = (s.l[0] & 1u) ? 0u : 0xffffffffu;
    uint32_t t[8], cc = 0;
    // t = flip ? r - s : s
    t[0] = sub_cc(FrParams::mod(0), s.l[0], cc);
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = subc_cc(FrParams::mod(i), s.l[i], cc);
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = flip ? t[i] : s.l[i];
    // (t - 1) / 2 + 2^254, then the flip
#pragma unroll
    for (int i = 0; i < 7; i++) out[i] = ((t[i] >> 1) | (t[i + 1] << 31)) ^ flip;
    out[7] = ((t[7] >> 1) | 0x40000000u) ^ (flip & 0x7fffffffu);
}
// g sign bits of one (group, bit position) pair, bit k = point k of the group  ->  table index and sign:
// the table only holds the half with +P_{q,0}; the other half is its negation.  Bit 31 = negate.
KZG_HD uint32_t comb_digit(uint32_t pattern, int g) {
    const uint32_t full = (1u << g) - 1u;  // g <= KZG_COMB_MAX_WIDTH
    const uint32_t neg = (~pattern) & 1u;
    if (neg) pattern = ~pattern & full;
    return ((pattern & full) >> 1) | (neg << 31);
}
// 32 x 32 bit-matrix transpose in place: afterwards bit j of a[i] is what bit i of a[j] was
// (the butterfly of Hacker's Delight 7-3, 5 rounds of 16 swaps).
KZG_HD void transpose32(uint32_t a[32]) {
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
            a[k] ^= t << j;
            a[k + j] ^= t;
        }
    }
}

// One blob element -> its sign words, with the range check of `bytes_to_bls_field` (reference
// src/utils.rs:262-275): a non-canonical element marks the blob BADARGS and counts as zero.
// sign_words layout: [blob][word w][point i], n_pad points per row (the padding points of the last group are
// points at infinity: their signs do not matter).
KZG_HD void blob_sign_words_thread(const uint8_t *blobs, uint64_t e, int n, int n_pad, uint32_t *sign_words, int *status) {
    uint64_t b = e / (uint64_t)n_pad;
    uint32_t i = (uint32_t)(e - b * n_pad);
    fr_t s;
    fe_set_zero(s);
    if (i < (uint32_t)n) {
        scalar_from_be32(s, blobs + (b * n + i) * 32);
        if (!fr_is_canonical(s)) {
#if defined(__CUDA_ARCH__)
            atomicMax(status + b, (int)KZG_BADARGS);
#else
            status[b] = KZG_BADARGS;
#endif
            fe_set_zero(s);
        }
    }
    uint32_t w[8];
    scalar_sign_words(w, s);
#pragma unroll
    for (int k = 0; k < 8; k++) sign_words[(b * 8 + k) * (uint64_t)n_pad + i] = w[k];
}
// Same from canonical little-endian limbs (the quotient polynomial of a proof), scalars[b*n + i].
KZG_HD void fr_sign_words_thread(const fr_t *scalars, uint64_t e, int n, int n_pad, uint32_t *sign_words) {
    uint64_t b = e / (uint64_t)n_pad;
    uint32_t i = (uint32_t)(e - b * n_pad);
    fr_t s;
    fe_set_zero(s);
    if (i < (uint32_t)n) s = scalars[b * n + i];
    uint32_t w[8];
    scalar_sign_words(w, s);
#pragma unroll
    for (int k = 0; k < 8; k++) sign_words[(b * 8 + k) * (uint64_t)n_pad + i] = w[k];
}
// Digits of (blob b, group q): for every bit position j the g sign bits of the group's points, as table index + sign.
//   digits[(q*255 + j)*count + b]   (group-major, what the gather level of the MSM reads, coalesced in b)
KZG_HD void comb_index_thread(const uint32_t *sign_words, uint32_t b, uint32_t q, int g, int n_pad, uint32_t count,
                              uint32_t *digits) {
#pragma unroll 1
    for (int w = 0; w < 8; w++) {
        const uint32_t *row = sign_words + ((uint64_t)b * 8 + w) * (uint64_t)n_pad + (uint64_t)q * g;
        uint32_t a[32];
#pragma unroll
        for (int k = 0; k < 32; k++) a[k] = k < g ? row[k] : 0u;
        transpose32(a);  // a[t] bit k = sign of point k at bit position 32 w + t
#pragma unroll
        for (int t = 0; t < 32; t++) {
            const int j = 32 * w + t;
            if (j < KZG_COMB_WINDOWS) digits[((uint64_t)q * KZG_COMB_WINDOWS + j) * count + b] = comb_digit(a[t], g);
        }
    }
}

// Horner pass over the W window sums of one blob: sum_j 2^(c j) S_j, S_j affine (possibly
// infinity) at sums[j*stride] (the MSM leaves them window-major: stride = blobs in the chunk).
// (W-1)(c doublings + 1 mixed addition) in Jacobian coordinates.
// The sums come out of the lazy levels of the MSM with coordinates in [0, 2p): made canonical here.
KZG_HD g1_affine_t horner_load(const g1_affine_t *p) {
    g1_affine_t s = *p;
    if (!g1a_is_inf(s)) { fe_canonical(s.x); fe_canonical(s.y); }
    return s;
}
KZG_HD void horner_thread(g1_affine_t &out, const g1_affine_t *sums, size_t stride, int c, int W) {
    g1_jac_t acc;
    g1j_from_affine(acc, horner_load(sums + (size_t)(W - 1) * stride));
#pragma unroll 1
    for (int j = W - 2; j >= 0; j--) {
#pragma unroll 1
        for (int k = 0; k < c; k++) g1j_dbl(acc, acc);
        g1_affine_t s = horner_load(sums + (size_t)j * stride);
        if (!g1a_is_inf(s)) g1j_add_affine(acc, acc, s.x, s.y);
    }
    g1j_to_affine(out, acc);
}

// Setup / validation: one compressed point -> affine Montgomery (+ optional subgroup check,
// reference `validate_kzg_g1` src/utils.rs:282-315; infinity is accepted).
KZG_HD int g1_decode_thread(g1_affine_t &out, const uint8_t *in, bool check_subgroup) {
    if (!g1a_uncompress(out, in)) { g1a_set_inf(out); return KZG_BADARGS; }
    if (check_subgroup && !g1a_is_inf(out) && !g1a_in_subgroup(out)) return KZG_BADARGS;
    return KZG_OK;
}

}  // namespace kzg

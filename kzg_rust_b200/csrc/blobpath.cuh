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

// Number of signed c-bit windows needed for any scalar s < r: ceil(255/c), plus one when
// the top window of r-1 (plus an incoming carry) could exceed 2^(c-1).  c = 15 -> 17.
inline int msm_num_windows(int c) {
    const uint32_t rm[8] = {FR_R_LIMBS};
    int W = (255 + c - 1) / c;
    int bit = c * (W - 1);
    uint64_t top = 0;  // (r-1) >> bit; r is odd so r-1 only clears bit 0, irrelevant unless bit == 0
    for (int k = 0; k < 32 && bit + k < 256; k++) top |= (uint64_t)((rm[(bit + k) >> 5] >> ((bit + k) & 31)) & 1) << k;
    if (top + 1 > (1ull << (c - 1))) W++;
    return W;
}

// Signed c-bit digits of a canonical scalar s < r:  s = sum_j d_j 2^(c j), |d_j| <= 2^(c-1),
// W = msm_num_windows(c) so the top window never carries out.  Digit j goes to out[j*stride].
KZG_HD void recode_signed(const fr_t &s, int c, int W, int32_t *out, uint64_t stride) {
    const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
    uint32_t carry = 0;
#pragma unroll 1
    for (int j = 0; j < W; j++) {
        int bit = j * c;
        int w = bit >> 5, sh = bit & 31;
        uint64_t lo = w < 8 ? s.l[w] : 0u, hi = (w + 1) < 8 ? s.l[w + 1] : 0u;
        uint32_t v = (uint32_t)(((hi << 32) | lo) >> sh) & mask;
        v += carry;
        int d;
        if (v > half) { d = (int)v - (int)(1u << c); carry = 1; } else { d = (int)v; carry = 0; }
        out[(uint64_t)j * stride] = d;
    }
}

// One blob element: range check + digits.  Non-canonical elements mark the blob BADARGS
// (reference src/utils.rs:266-270) and contribute zero digits.
KZG_HD void blob_digits_thread(const uint8_t *blobs, uint64_t e, int n, int c, int W, int32_t *digits, int *status) {
    uint64_t b = e / (uint64_t)n;
    uint32_t i = (uint32_t)(e - b * n);
    fr_t s;
    scalar_from_be32(s, blobs + e * 32);
    if (!fr_is_canonical(s)) {
#if defined(__CUDA_ARCH__)
        atomicMax(status + b, (int)KZG_BADARGS);
#else
        status[b] = KZG_BADARGS;
#endif
        fe_set_zero(s);
    }
    recode_signed(s, c, W, digits + (b * W) * (uint64_t)n + i, (uint64_t)n);
}
// Same, from an Fr element in Montgomery form (the quotient polynomial of a proof).
KZG_HD void fr_digits_thread(const fr_t *evals, uint64_t e, int n, int c, int W, int32_t *digits) {
    uint64_t b = e / (uint64_t)n;
    uint32_t i = (uint32_t)(e - b * n);
    fr_t s;
    fe_from_mont(s, evals[e]);
    recode_signed(s, c, W, digits + (b * W) * (uint64_t)n + i, (uint64_t)n);
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

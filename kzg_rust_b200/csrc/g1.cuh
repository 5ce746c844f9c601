// g1.cuh -- BLS12-381 G1 on the device (and host): affine points for the batched
// additions of the MSM, Jacobian points for the few per-point scalar multiplications, and
// the ZCash (de)compression that blst_p1_compress / blst_p1_uncompress implement
// (reference src/utils.rs:221-227, 282-315).
#pragma once
#include "fields.cuh"

namespace kzg {

// Affine point, Montgomery coordinates, 96 B.  Infinity is encoded by an x whose top limb
// is 0xffffffff (no field element looks like that: p < 2^381).
struct alignas(16) g1_affine_t {
    fp_t x, y;
};
KZG_HD bool g1a_is_inf(const g1_affine_t &p) { return p.x.l[11] == 0xffffffffu; }
KZG_HD bool fp_is_inf_marker(const fp_t &x) { return x.l[11] == 0xffffffffu; }
KZG_HD void g1a_set_inf(g1_affine_t &p) {
#pragma unroll
    for (int i = 0; i < 12; i++) { p.x.l[i] = 0xffffffffu; p.y.l[i] = 0; }
}

// ------------------------------------------------------------------ one step of a batched affine addition
// R = P1 + P2 is computed in two passes around one shared inversion (Montgomery's trick):
//   pass 1: den = add_denominator(...)       -> multiplied into the running product
//   pass 2: add_finish(..., 1/den)           -> the sum
// Every case of the group law is exact:
//   P1 or P2 infinity -> den = 1, R = the other operand
//   x1 != x2          -> den = x2 - x1,  lambda = (y2 - y1)/den
//   P1 == P2          -> den = 2 y1,     lambda = 3 x1^2/den      (y1 != 0: the curve has odd order)
//   P1 == -P2         -> den = 1, R = infinity
enum { ADD_GENERIC = 0, ADD_TAKE_P2 = 1, ADD_TAKE_P1 = 2, ADD_DOUBLE = 3, ADD_CANCEL = 4 };

// LAZY = true: coordinates, denominators and inverses are residues in [0, 2p) and stay so (fe_mul_lazy /
// fe_sub_lazy, bigint.cuh); only the rare equal-x cases make their operands canonical.
template <bool LAZY> KZG_HD void fpx_mul(fp_t &r, const fp_t &a, const fp_t &b) {
    if (LAZY) fe_mul_lazy(r, a, b); else fe_mul(r, a, b);
}
template <bool LAZY> KZG_HD void fpx_sub(fp_t &r, const fp_t &a, const fp_t &b) {
    if (LAZY) fe_sub_lazy(r, a, b); else fe_sub(r, a, b);
}
// y1/y2 are only read when x1 == x2 (the caller passes loaders so the common case does
// not touch them in pass 1).
template <bool LAZY, class LoadY1, class LoadY2>
KZG_HD int add_denominator(fp_t &den, const fp_t &x1, const fp_t &x2, LoadY1 load_y1, LoadY2 load_y2) {
    if (fp_is_inf_marker(x1)) { den = fe_one<FpParams>(); return ADD_TAKE_P2; }
    if (fp_is_inf_marker(x2)) { den = fe_one<FpParams>(); return ADD_TAKE_P1; }
    if (LAZY) fe_sub_lazy4(den, x2, x1);  // in (0, 4p): only ever multiplied by values below 2p
    else fe_sub(den, x2, x1);
    if (!(LAZY ? fe_is_zero_lazy4(den) : fe_is_zero(den))) return ADD_GENERIC;
    fp_t y1, y2;
    load_y1(y1);
    load_y2(y2);
    if (LAZY) { fe_canonical(y1); fe_canonical(y2); }
    if (fe_eq(y1, y2) && !fe_is_zero(y1)) { fe_dbl(den, y1); return ADD_DOUBLE; }
    den = fe_one<FpParams>();
    return ADD_CANCEL;
}
// inv = 1/den for this addition
template <bool LAZY>
KZG_HD void add_finish(g1_affine_t &r, int kind, const g1_affine_t &p1, const g1_affine_t &p2, const fp_t &inv) {
    if (kind == ADD_TAKE_P2) { r = p2; return; }
    if (kind == ADD_TAKE_P1) { r = p1; return; }
    if (kind == ADD_CANCEL) { g1a_set_inf(r); return; }
    fp_t num, lam, t;
    if (kind == ADD_DOUBLE) {
        fp_t x = p1.x;
        if (LAZY) fe_canonical(x);
        fe_sqr(t, x);
        fe_dbl(num, t);
        fe_add(num, num, t);
    } else {
        if (LAZY) fe_sub_lazy4(num, p2.y, p1.y);  // in (0, 4p), times inv < 2p
        else fe_sub(num, p2.y, p1.y);
    }
    fpx_mul<LAZY>(lam, num, inv);
    fpx_mul<LAZY>(t, lam, lam);
    fpx_sub<LAZY>(t, t, p1.x);
    fpx_sub<LAZY>(t, t, p2.x);  // x3 (p2.x == p1.x mod p when doubling)
    fp_t u;
    if (LAZY) fe_sub_lazy4(u, p1.x, t);  // in (0, 4p), times lam < 2p
    else fe_sub(u, p1.x, t);
    fpx_mul<LAZY>(u, lam, u);
    fpx_sub<LAZY>(r.y, u, p1.y);
    r.x = t;
}

// ------------------------------------------------------------------ Jacobian (Z == 0 is infinity)
struct g1_jac_t {
    fp_t x, y, z;
};
KZG_HD bool g1j_is_inf(const g1_jac_t &p) { return fe_is_zero(p.z); }
KZG_HD void g1j_set_inf(g1_jac_t &p) { fe_set_zero(p.x); fe_set_zero(p.y); fe_set_zero(p.z); }
KZG_HD void g1j_from_affine(g1_jac_t &r, const g1_affine_t &a) {
    if (g1a_is_inf(a)) { g1j_set_inf(r); return; }
    r.x = a.x; r.y = a.y; r.z = fe_one<FpParams>();
}
// The Jacobian formulas below run where one thread owns a long dependent chain (the Horner pass, the scalar
// ladders, the subgroup checks, the sums of batch verification): their products are issued in independent PAIRS
// (fe_mul2, bigint.cuh) so that a latency-bound thread keeps two carry chains in flight.
KZG_HD void g1j_dbl(g1_jac_t &r, const g1_jac_t &p) {
    // a = 0 doubling; infinity maps to infinity because Z3 = 2 Y Z
    fp_t A, B, C, D, E, F, t, z3;
    fe_mul2(A, p.x, p.x, B, p.y, p.y);
    fe_add(t, p.x, B);
    fe_mul2(C, B, B, t, t, t);
    fe_sub(t, t, A); fe_sub(t, t, C); fe_dbl(D, t);
    fe_dbl(E, A); fe_add(E, E, A);
    fe_mul2(F, E, E, z3, p.y, p.z);
    fe_dbl(z3, z3);
    fe_dbl(t, D); fe_sub(r.x, F, t);
    fe_sub(t, D, r.x); fe_mul(t, E, t);
    fe_dbl(C, C); fe_dbl(C, C); fe_dbl(C, C);
    fe_sub(r.y, t, C);
    r.z = z3;
}
#if defined(__CUDACC__)
// n doublings in a row by THREE lanes of a warp.  A chain of doublings on one thread is pure latency (seven products in
// four dependent steps of fe_mul2 plus fourteen additions, ~7 us); the seven products of a doubling are three levels of
// independent ones, so three lanes take one product per level and pass the results round by shuffles:
//     lane `leader`     : Y Z           (X + B)^2        E (D - X3)       and every sum / difference
//     lane `leader + 1` : A = X^2       F = (3 A)^2
//     lane `leader + 2` : B = Y^2       C = B^2
// -- three dependent products per doubling instead of four pairs, ~3.3 us.  Same formulas as g1j_dbl, canonical residues,
// so the result is the same Jacobian triple.  The three lanes call together (uniform n, `mask` names all lanes of the warp
// that take part, possibly of several groups); p is read and written on the leader only, role = lane - leader.
KZG_D fp_t fp_shfl(uint32_t mask, const fp_t &v, int src) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(mask, v.l[i], src);
    return r;
}
KZG_D void g1j_dbl_n_coop3(g1_jac_t &p, uint32_t n, uint32_t mask, int leader, int role) {
    fp_t X = fp_shfl(mask, p.x, leader), Y = fp_shfl(mask, p.y, leader), Z = p.z;
#pragma unroll 1
    for (uint32_t t = 0; t < n; t++) {
        fp_t m1, m2, m3, e, tx;
        {
            const fp_t a = role == 1 ? X : Y, b = role == 0 ? Z : (role == 1 ? X : Y);
            fe_mul(m1, a, b);  // 0: Y Z   1: A   2: B
        }
        const fp_t Bq = fp_shfl(mask, m1, leader + 2);
        fe_dbl(e, m1); fe_add(e, e, m1);  // lane 1: E = 3 A
        fe_add(tx, X, Bq);                // lane 0: X + B
        {
            const fp_t a = role == 0 ? tx : (role == 1 ? e : m1);
            fe_mul(m2, a, a);  // 0: (X + B)^2   1: F = E^2   2: C = B^2
        }
        const fp_t Aq = fp_shfl(mask, m1, leader + 1), Eq = fp_shfl(mask, e, leader + 1), Fq = fp_shfl(mask, m2, leader + 1);
        fp_t Cq = fp_shfl(mask, m2, leader + 2);
        fp_t D, X3, u, Y3;
        fe_sub(D, m2, Aq); fe_sub(D, D, Cq); fe_dbl(D, D);
        fe_dbl(u, D); fe_sub(X3, Fq, u);
        fe_sub(u, D, X3);
        fe_mul(m3, Eq, u);
        fe_dbl(Cq, Cq); fe_dbl(Cq, Cq); fe_dbl(Cq, Cq);
        fe_sub(Y3, m3, Cq);
        fe_dbl(Z, m1);
        X = fp_shfl(mask, X3, leader);
        Y = fp_shfl(mask, Y3, leader);
    }
    if (role == 0) { p.x = X; p.y = Y; p.z = Z; }
}
#endif
// r = p + q (q affine, not infinity), complete
KZG_HD void g1j_add_affine(g1_jac_t &r, const g1_jac_t &p, const fp_t &qx, const fp_t &qy) {
    if (g1j_is_inf(p)) { r.x = qx; r.y = qy; r.z = fe_one<FpParams>(); return; }
    fp_t z1z1, u2, s2, h, hh, i, j, rr, v, t, tz;
    fe_mul2(z1z1, p.z, p.z, s2, qy, p.z);
    fe_mul2(u2, qx, z1z1, s2, s2, z1z1);
    if (fe_eq(p.x, u2)) {
        if (fe_eq(p.y, s2)) { g1j_dbl(r, p); return; }
        g1j_set_inf(r);
        return;
    }
    fe_sub(h, u2, p.x);
    fe_add(tz, p.z, h);
    fe_mul2(hh, h, h, tz, tz, tz);
    fe_dbl(i, hh); fe_dbl(i, i);
    fe_mul2(j, h, i, v, p.x, i);
    fe_sub(rr, s2, p.y); fe_dbl(rr, rr);
    fp_t x3, y3, z3;
    fe_mul2(x3, rr, rr, s2, p.y, j);
    fe_sub(x3, x3, j); fe_dbl(t, v); fe_sub(x3, x3, t);
    fe_sub(t, v, x3); fe_mul(t, rr, t);
    fe_dbl(s2, s2);
    fe_sub(y3, t, s2);
    fe_sub(tz, tz, z1z1); fe_sub(z3, tz, hh);
    r.x = x3; r.y = y3; r.z = z3;
}
// r = p + q, both Jacobian, complete
KZG_HD void g1j_add(g1_jac_t &r, const g1_jac_t &p, const g1_jac_t &q) {
    if (g1j_is_inf(p)) { r = q; return; }
    if (g1j_is_inf(q)) { r = p; return; }
    fp_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, tz;
    fe_mul2(z1z1, p.z, p.z, z2z2, q.z, q.z);
    fe_mul2(u1, p.x, z2z2, u2, q.x, z1z1);
    fe_mul2(s1, p.y, q.z, s2, q.y, p.z);
    fe_mul2(s1, s1, z2z2, s2, s2, z1z1);
    if (fe_eq(u1, u2)) {
        if (fe_eq(s1, s2)) { g1j_dbl(r, p); return; }
        g1j_set_inf(r);
        return;
    }
    fe_sub(h, u2, u1);
    fe_dbl(i, h);
    fe_add(tz, p.z, q.z);
    fe_mul2(i, i, i, tz, tz, tz);
    fe_mul2(j, h, i, v, u1, i);
    fe_sub(rr, s2, s1); fe_dbl(rr, rr);
    fp_t x3, y3, z3;
    fe_mul2(x3, rr, rr, s1, s1, j);
    fe_sub(x3, x3, j); fe_dbl(t, v); fe_sub(x3, x3, t);
    fe_sub(t, v, x3);
    fe_sub(tz, tz, z1z1); fe_sub(tz, tz, z2z2);
    fe_mul2(t, rr, t, z3, tz, h);
    fe_dbl(s1, s1);
    fe_sub(y3, t, s1);
    r.x = x3; r.y = y3; r.z = z3;
}
KZG_HD void g1j_to_affine(g1_affine_t &r, const g1_jac_t &p) {
    if (g1j_is_inf(p)) { g1a_set_inf(r); return; }
    fp_t zi, zi2, zi3;
    fp_inv(zi, p.z);
    fe_sqr(zi2, zi);
    fe_mul(zi3, zi2, zi);
    fe_mul(r.x, p.x, zi2);
    fe_mul(r.y, p.y, zi3);
}
// [k]P for a canonical little-endian scalar of `nbits` bits (double-and-add, P affine)
KZG_HD void g1j_mul(g1_jac_t &r, const g1_affine_t &p, const uint32_t *k, int nbits) {
    g1_jac_t acc;
    g1j_set_inf(acc);
    if (g1a_is_inf(p)) { r = acc; return; }
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        g1j_dbl(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1) g1j_add_affine(acc, acc, p.x, p.y);
    }
    r = acc;
}

// [k]P for a full-width scalar with the G1 endomorphism psi(x, y) = (beta^2 x, y) = [lambda]P,
// lambda = x^2 - 1 (x the curve parameter; lambda^2 + lambda + 1 = 0 mod r, 128 bits):
//   k = k2 lambda + k1  (plain integer division, both halves below 2^128)
//   [k]P = [k1]P + [k2]psi(P), one joint ladder of 128 doublings instead of 255,
// and the third table entry is free: P + psi(P) = -psi^2(P) = (beta x, -y).
// Used by the r-power terms of batch verification, whose latency is one such ladder per blob.
KZG_HD void glv_split(uint32_t k1[4], uint32_t k2[4], const uint32_t k[8]) {
    // Barrett division by lambda = BLS_X_SQUARED - 1 (HAC 14.42, b = 2^32, four-limb modulus): mu = floor(2^256 / lambda),
    // q = floor(floor(k / 2^96) mu / 2^160) is floor(k / lambda) or up to 2 less; the remainder is fixed up by subtraction.
    // (A bit-serial long division stood here: 256 dependent rounds, most of k_pip_scalars' 80 us.)
    constexpr uint32_t xs[4] = {BLS_X_SQUARED_LIMBS};
    const uint32_t lam[5] = {xs[0] - 1u, xs[1] - (xs[0] == 0 ? 1u : 0u), xs[2], xs[3], 0u};  // low limb of x^2 is 0: borrow into limb 1
    constexpr uint32_t mu[5] = {0xf6cfee30u, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x00000001u};
    uint32_t q2[10];
#pragma unroll
    for (int i = 0; i < 10; i++) q2[i] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {  // q2 = (k >> 96) * mu
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            c += (uint64_t)k[3 + i] * mu[j] + q2[i + j];
            q2[i + j] = (uint32_t)c;
            c >>= 32;
        }
        q2[i + 5] = (uint32_t)c;
    }
    uint32_t q[5], r[5];
#pragma unroll
    for (int i = 0; i < 5; i++) { q[i] = q2[5 + i]; r[i] = 0; }
#pragma unroll
    for (int i = 0; i < 5; i++) {  // r = (q * lambda) mod 2^160
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j + i < 5; j++) {
            c += (uint64_t)q[i] * lam[j] + r[i + j];
            r[i + j] = (uint32_t)c;
            c >>= 32;
        }
    }
    {  // r = k - q lambda (mod 2^160): in [0, 3 lambda)
        uint32_t cc = 0;
        r[0] = sub_cc(k[0], r[0], cc);
#pragma unroll
        for (int i = 1; i < 5; i++) r[i] = subc_cc(k[i], r[i], cc);
    }
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        uint32_t d[5], cc = 0;
        d[0] = sub_cc(r[0], lam[0], cc);
#pragma unroll
        for (int i = 1; i < 5; i++) d[i] = subc_cc(r[i], lam[i], cc);
        const uint32_t borrow = subc(0, 0, cc);
        if (borrow != 0) break;
#pragma unroll
        for (int i = 0; i < 5; i++) r[i] = d[i];
        uint32_t c2 = 0;
        q[0] = add_cc(q[0], 1u, c2);
#pragma unroll
        for (int i = 1; i < 5; i++) q[i] = addc_cc(q[i], 0u, c2);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { k1[i] = r[i]; k2[i] = q[i]; }  // q < 2^128 for k < 2^255 (r / lambda < 2^128)
}
KZG_HD void g1j_mul_glv(g1_jac_t &r, const g1_affine_t &p, const uint32_t *k) {
    g1_jac_t acc;
    g1j_set_inf(acc);
    if (g1a_is_inf(p)) { r = acc; return; }
    uint32_t k1[4], k2[4];
    glv_split(k1, k2, k);
    fp_t beta, x2, x3, yn;
    {
        constexpr uint32_t bm[12] = {FP_BETA_MONT_LIMBS};
#pragma unroll
        for (int i = 0; i < 12; i++) beta.l[i] = bm[i];
    }
    fe_mul(x3, p.x, beta);   // P + psi(P) = (beta x, -y)
    fe_mul(x2, x3, beta);    // psi(P)     = (beta^2 x, y)
    fe_neg(yn, p.y);
#pragma unroll 1
    for (int i = 127; i >= 0; i--) {
        g1j_dbl(acc, acc);
        uint32_t sel = ((k1[i >> 5] >> (i & 31)) & 1u) | (((k2[i >> 5] >> (i & 31)) & 1u) << 1);
        if (sel) {
            fp_t qx = sel == 1 ? p.x : (sel == 2 ? x2 : x3);
            fp_t qy = sel == 3 ? yn : p.y;
            g1j_add_affine(acc, acc, qx, qy);
        }
    }
    r = acc;
}

// ------------------------------------------------------------------ serialisation
// 48 big-endian bytes <-> 12 canonical limbs
KZG_HD void fp_from_be48(fp_t &r, const uint8_t *b) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint8_t *q = b + 4 * (11 - i);
        r.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
}
KZG_HD void fp_to_be48(uint8_t *b, const fp_t &canon) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint8_t *q = b + 4 * (11 - i);
        uint32_t v = canon.l[i];
        q[0] = (uint8_t)(v >> 24); q[1] = (uint8_t)(v >> 16); q[2] = (uint8_t)(v >> 8); q[3] = (uint8_t)v;
    }
}
// blst_p1_compress (reference src/utils.rs:221-227)
KZG_HD void g1a_compress(uint8_t out[48], const g1_affine_t &p) {
    if (g1a_is_inf(p)) {
#pragma unroll
        for (int i = 0; i < 48; i++) out[i] = 0;
        out[0] = 0xC0;
        return;
    }
    fp_t cx;
    fe_from_mont(cx, p.x);
    fp_to_be48(out, cx);
    out[0] |= 0x80;
    if (fp_is_lexicographically_largest(p.y)) out[0] |= 0x20;
}
// blst_p1_uncompress: false on any malformed encoding (bit 7 clear, bad infinity, x >= p,
// x^3 + 4 not a square).  No subgroup check here.
KZG_HD bool g1a_uncompress(g1_affine_t &out, const uint8_t in[48]) {
    uint8_t b0 = in[0];
    if (!(b0 & 0x80)) return false;
    if (b0 & 0x40) {
        uint32_t o = b0 & 0x3F;
#pragma unroll 1
        for (int i = 1; i < 48; i++) o |= in[i];
        if (o) return false;
        g1a_set_inf(out);
        return true;
    }
    uint8_t tmp[48];
#pragma unroll
    for (int i = 0; i < 48; i++) tmp[i] = in[i];
    tmp[0] &= 0x1F;
    fp_t x;
    fp_from_be48(x, tmp);
    uint32_t m[12];
#pragma unroll
    for (int i = 0; i < 12; i++) m[i] = FpParams::mod(i);
    if (limbs_geq<12>(x.l, m)) return false;
    fe_to_mont(x, x);
    fp_t rhs, y;
    fe_sqr(rhs, x);
    fe_mul(rhs, rhs, x);
    fe_add(rhs, rhs, fp_const_b());
    if (!fp_sqrt(y, rhs)) return false;
    if (fp_is_lexicographically_largest(y) != ((b0 & 0x20) != 0)) fe_neg(y, y);
    out.x = x;
    out.y = y;
    return true;
}
// blst_p1_in_g1 stand-in.  For BLS12 curves P lies in G1 exactly when phi(P) = -[x^2]P, with
// phi(x, y) = (beta x, y) the endomorphism acting on G1 as multiplication by -x^2 and x the
// curve parameter (M. Scott, "A note on group membership tests for G1, G2 and GT on BLS
// pairing-friendly curves", 2021): one 128-bit multiplication instead of the 255-bit [r]P.
// P must be on the curve and not infinity.
KZG_HD bool g1a_in_subgroup(const g1_affine_t &p) {
    uint32_t k[4] = {BLS_X_SQUARED_LIMBS};
    g1_jac_t q;
    g1j_mul(q, p, k, 128);
    if (g1j_is_inf(q)) return false;
    fp_t beta, z2, z3, t;
    {
        constexpr uint32_t bm[12] = {FP_BETA_MONT_LIMBS};
#pragma unroll
        for (int i = 0; i < 12; i++) beta.l[i] = bm[i];
    }
    fe_sqr(z2, q.z);
    fe_mul(z3, z2, q.z);
    fe_mul(t, p.x, beta);
    fe_mul(t, t, z2);
    if (!fe_eq(t, q.x)) return false;
    fe_mul(t, p.y, z3);
    fe_neg(t, t);
    return fe_eq(t, q.y);
}
#if defined(__CUDACC__)
// The same test by three lanes per point (leader, leader + 1, leader + 2; p on the leader): [x^2]P as [|x|]([|x|]P) --
// |x| = 0xd201000000010000 has six set bits, so the two 64-bit multiplications are 126 doublings (shared by the three
// lanes, g1j_dbl_n_coop3) and 10 additions where the 128-bit double-and-add over x^2 takes 127 and ~30.
KZG_D void g1j_mul_x_coop3(g1_jac_t &r, const g1_jac_t &p, uint32_t mask, int leader, int role) {
    g1_jac_t acc = p;
    const uint32_t runs[6] = {1, 2, 3, 9, 32, 16};  // gaps between the set bits 63, 62, 60, 57, 48, 16, 0
#pragma unroll 1
    for (int i = 0; i < 6; i++) {
        g1j_dbl_n_coop3(acc, runs[i], mask, leader, role);
        if (i < 5 && role == 0) g1j_add(acc, acc, p);
    }
    r = acc;
}
KZG_D bool g1a_in_subgroup_coop3(const g1_affine_t &p, uint32_t mask, int leader, int role) {
    g1_jac_t b, q;
    b.x = p.x; b.y = p.y; b.z = fe_one<FpParams>();
    g1j_mul_x_coop3(q, b, mask, leader, role);
    g1j_mul_x_coop3(q, q, mask, leader, role);
    if (role != 0 || g1j_is_inf(q)) return false;
    fp_t beta, z2, z3, t;
    {
        constexpr uint32_t bm[12] = {FP_BETA_MONT_LIMBS};
#pragma unroll
        for (int i = 0; i < 12; i++) beta.l[i] = bm[i];
    }
    fe_sqr(z2, q.z);
    fe_mul(z3, z2, q.z);
    fe_mul(t, p.x, beta);
    fe_mul(t, t, z2);
    if (!fe_eq(t, q.x)) return false;
    fe_mul(t, p.y, z3);
    fe_neg(t, t);
    return fe_eq(t, q.y);
}
#endif
// the defining test, [r]P == infinity (kept for cross-checks)
KZG_HD bool g1a_in_subgroup_by_order(const g1_affine_t &p) {
    uint32_t k[8];
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = FrParams::mod(i);
    g1_jac_t t;
    g1j_mul(t, p, k, 255);
    return g1j_is_inf(t);
}

}  // namespace kzg

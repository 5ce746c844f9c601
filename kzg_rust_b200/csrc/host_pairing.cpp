// host_pairing.cpp -- the host-side final check of the verification path.
//
// The reference ends `verify_kzg_proof_batch` with one call of `pairings_verify`
// (reference src/kzg.rs:625 -> src/utils.rs:189-214: two Miller loops, one final
// exponentiation, all inside blst) and checks the trusted setup with another one
// (src/kzg.rs:802-830).  The north star keeps that single check on the host; blst is not
// available in this environment, so this file is a self-contained BLS12-381 optimal-ate
// pairing in portable C++ (64-bit limbs, unsigned __int128):
//
//   Fp  : 6 x 64-bit Montgomery (radix 2^384, same constants as the device code)
//   Fp2 : Fp[u]/(u^2+1);  Fp6 : Fp2[v]/(v^3 - (1+u));  Fp12 : Fp6[w]/(w^2 - v)
//   G2 arguments are fixed per context ([tau]G2 and the generator), so their Miller-loop
//   lines (slope, intercept) are computed once ("prepared") and a verification only
//   evaluates them at the two G1 points: per step one Fp12 squaring + two sparse products.
//   Final exponentiation: easy part by conjugation/inversion/Frobenius, hard part by the
//   (x-1)^2 (x+p) (x^2+p^2-1) + 3 chain (yields the cube of the canonical value, which is
//   1 exactly when the canonical value is 1).
#include "host_pairing.h"

#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <mutex>
#include <vector>

#include "consts_gen.cuh"

namespace {

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ------------------------------------------------------------------ Fp
struct Fp { u64 l[6]; };

static u64 P64[6], R1_64[6], R2_64[6], N0_64;
static u64 EXP_P_PLUS_1_DIV_4[6], EXP_P_MINUS_3_DIV_4[6], EXP_P_MINUS_1_DIV_2[6], EXP_P_MINUS_2[6], EXP_P_MINUS_1_DIV_6[6];

static void pack(u64 *out, const uint32_t *in, int n64) {
    for (int i = 0; i < n64; i++) out[i] = (u64)in[2 * i] | ((u64)in[2 * i + 1] << 32);
}

static inline bool fp_is_zero(const Fp &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5]) == 0; }
static inline bool fp_eq(const Fp &a, const Fp &b) {
    u64 o = 0;
    for (int i = 0; i < 6; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}
static inline bool limbs_geq(const u64 *a, const u64 *b) {
    for (int i = 5; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
static inline void fp_add(Fp &r, const Fp &a, const Fp &b) {
    u128 c = 0;
    Fp t;
    for (int i = 0; i < 6; i++) { c += (u128)a.l[i] + b.l[i]; t.l[i] = (u64)c; c >>= 64; }
    if (c || limbs_geq(t.l, P64)) {
        u128 bw = 0;
        for (int i = 0; i < 6; i++) { u128 d = (u128)t.l[i] - P64[i] - (u64)bw; t.l[i] = (u64)d; bw = (d >> 64) & 1; }
    }
    r = t;
}
static inline void fp_sub(Fp &r, const Fp &a, const Fp &b) {
    u128 bw = 0;
    Fp t;
    for (int i = 0; i < 6; i++) { u128 d = (u128)a.l[i] - b.l[i] - (u64)bw; t.l[i] = (u64)d; bw = (d >> 64) & 1; }
    if (bw) {
        u128 c = 0;
        for (int i = 0; i < 6; i++) { c += (u128)t.l[i] + P64[i]; t.l[i] = (u64)c; c >>= 64; }
    }
    r = t;
}
static inline void fp_neg(Fp &r, const Fp &a) { Fp z = {}; fp_sub(r, z, a); }
static inline void fp_dbl(Fp &r, const Fp &a) { fp_add(r, a, a); }
// coarsely integrated operand scanning, portable form
static void fp_mul_portable(Fp &r, const Fp &a, const Fp &b) {
    u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
        u128 c = 0;
        for (int j = 0; j < 6; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[6]; t[6] = (u64)c; t[7] = (u64)(c >> 64);
        u64 m = t[0] * N0_64;
        c = (u128)m * P64[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 6; j++) { c += (u128)m * P64[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[6]; t[5] = (u64)c; t[6] = t[7] + (u64)(c >> 64);
    }
    Fp o;
    memcpy(o.l, t, sizeof o.l);
    if (t[6] || limbs_geq(o.l, P64)) {
        u128 bw = 0;
        for (int i = 0; i < 6; i++) { u128 d = (u128)o.l[i] - P64[i] - (u64)bw; o.l[i] = (u64)d; bw = (d >> 64) & 1; }
    }
    r = o;
}
#if defined(__x86_64__)
// the same product on mulx with two add-with-carry chains per row (the instruction mix of blst's mulx_mont_384);
// about 1.5x the portable form, chosen at start-up when the CPU has BMI2
#define KZG_ADDROW(T, X, m) do { \
        unsigned long long lo0, lo1, lo2, lo3, lo4, lo5, hi0, hi1, hi2, hi3, hi4, hi5; unsigned char c_; \
        lo0 = _mulx_u64((X)[0], m, &hi0); lo1 = _mulx_u64((X)[1], m, &hi1); lo2 = _mulx_u64((X)[2], m, &hi2); \
        lo3 = _mulx_u64((X)[3], m, &hi3); lo4 = _mulx_u64((X)[4], m, &hi4); lo5 = _mulx_u64((X)[5], m, &hi5); \
        c_ = _addcarry_u64(0, T##0, lo0, &T##0); c_ = _addcarry_u64(c_, T##1, lo1, &T##1); c_ = _addcarry_u64(c_, T##2, lo2, &T##2); \
        c_ = _addcarry_u64(c_, T##3, lo3, &T##3); c_ = _addcarry_u64(c_, T##4, lo4, &T##4); c_ = _addcarry_u64(c_, T##5, lo5, &T##5); \
        c_ = _addcarry_u64(c_, T##6, 0, &T##6); T##7 += c_; \
        c_ = _addcarry_u64(0, T##1, hi0, &T##1); c_ = _addcarry_u64(c_, T##2, hi1, &T##2); c_ = _addcarry_u64(c_, T##3, hi2, &T##3); \
        c_ = _addcarry_u64(c_, T##4, hi3, &T##4); c_ = _addcarry_u64(c_, T##5, hi4, &T##5); c_ = _addcarry_u64(c_, T##6, hi5, &T##6); \
        T##7 += c_; \
    } while (0)
__attribute__((target("bmi2"))) static void fp_mul_bmi2(Fp &r, const Fp &a, const Fp &b) {
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0;
#define KZG_STEP(i) do { \
        KZG_ADDROW(t, a.l, b.l[i]); \
        unsigned long long q_ = t0 * N0_64; \
        KZG_ADDROW(t, P64, q_); \
        t0 = t1; t1 = t2; t2 = t3; t3 = t4; t4 = t5; t5 = t6; t6 = t7; t7 = 0; \
    } while (0)
    KZG_STEP(0); KZG_STEP(1); KZG_STEP(2); KZG_STEP(3); KZG_STEP(4); KZG_STEP(5);
#undef KZG_STEP
    Fp o = {{t0, t1, t2, t3, t4, t5}};
    if (t6 || limbs_geq(o.l, P64)) {
        u128 bw = 0;
        for (int i = 0; i < 6; i++) { u128 d = (u128)o.l[i] - P64[i] - (u64)bw; o.l[i] = (u64)d; bw = (d >> 64) & 1; }
    }
    r = o;
}
#undef KZG_ADDROW
static void (*const fp_mul_impl)(Fp &, const Fp &, const Fp &) = __builtin_cpu_supports("bmi2") ? fp_mul_bmi2 : fp_mul_portable;
static inline void fp_mul(Fp &r, const Fp &a, const Fp &b) { fp_mul_impl(r, a, b); }
#else
static inline void fp_mul(Fp &r, const Fp &a, const Fp &b) { fp_mul_portable(r, a, b); }
#endif
static inline void fp_sqr(Fp &r, const Fp &a) { fp_mul(r, a, a); }
static Fp fp_one() { Fp r; memcpy(r.l, R1_64, sizeof r.l); return r; }
static void fp_to_mont(Fp &r, const Fp &a) { Fp r2; memcpy(r2.l, R2_64, sizeof r2.l); fp_mul(r, a, r2); }
static void fp_from_mont(Fp &r, const Fp &a) { Fp one = {{1, 0, 0, 0, 0, 0}}; fp_mul(r, a, one); }
static void fp_pow(Fp &r, const Fp &a, const u64 *e, int nlimbs) {
    Fp acc = fp_one();
    bool started = false;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        if (started) fp_sqr(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) {
            if (started) fp_mul(acc, acc, a); else { acc = a; started = true; }
        }
    }
    r = acc;
}
static void fp_inv(Fp &r, const Fp &a) { fp_pow(r, a, EXP_P_MINUS_2, 6); }
static bool fp_sqrt(Fp &r, const Fp &a) {
    Fp s, s2;
    fp_pow(s, a, EXP_P_PLUS_1_DIV_4, 6);
    fp_sqr(s2, s);
    r = s;
    return fp_eq(s2, a);
}
// canonical value > (p-1)/2 ?
static bool fp_is_large(const Fp &mont) {
    Fp c;
    fp_from_mont(c, mont);
    for (int i = 5; i >= 0; i--) {
        if (c.l[i] != EXP_P_MINUS_1_DIV_2[i]) return c.l[i] > EXP_P_MINUS_1_DIV_2[i];
    }
    return false;
}
// 48 big-endian bytes -> Montgomery; false when the value is >= p
static bool fp_from_be48(Fp &r, const uint8_t *b) {
    Fp t;
    for (int i = 0; i < 6; i++) {
        u64 v = 0;
        for (int k = 0; k < 8; k++) v = (v << 8) | b[8 * (5 - i) + k];
        t.l[i] = v;
    }
    if (limbs_geq(t.l, P64)) return false;
    fp_to_mont(r, t);
    return true;
}

// ------------------------------------------------------------------ Fp2
struct Fp2 { Fp c0, c1; };
static inline void fp2_add(Fp2 &r, const Fp2 &a, const Fp2 &b) { fp_add(r.c0, a.c0, b.c0); fp_add(r.c1, a.c1, b.c1); }
static inline void fp2_sub(Fp2 &r, const Fp2 &a, const Fp2 &b) { fp_sub(r.c0, a.c0, b.c0); fp_sub(r.c1, a.c1, b.c1); }
static inline void fp2_neg(Fp2 &r, const Fp2 &a) { fp_neg(r.c0, a.c0); fp_neg(r.c1, a.c1); }
static inline void fp2_dbl(Fp2 &r, const Fp2 &a) { fp_dbl(r.c0, a.c0); fp_dbl(r.c1, a.c1); }
static inline void fp2_conj(Fp2 &r, const Fp2 &a) { r.c0 = a.c0; fp_neg(r.c1, a.c1); }
static inline bool fp2_is_zero(const Fp2 &a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
static inline bool fp2_eq(const Fp2 &a, const Fp2 &b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
static Fp2 fp2_zero() { Fp2 r; memset(&r, 0, sizeof r); return r; }
static Fp2 fp2_one() { Fp2 r = fp2_zero(); r.c0 = fp_one(); return r; }
static void fp2_mul(Fp2 &r, const Fp2 &a, const Fp2 &b) {
    Fp t0, t1, t2, s0, s1;
    fp_mul(t0, a.c0, b.c0);
    fp_mul(t1, a.c1, b.c1);
    fp_add(s0, a.c0, a.c1);
    fp_add(s1, b.c0, b.c1);
    fp_mul(t2, s0, s1);
    fp_sub(r.c0, t0, t1);
    fp_sub(t2, t2, t0);
    fp_sub(r.c1, t2, t1);
}
static void fp2_sqr(Fp2 &r, const Fp2 &a) {
    Fp s, d, m;
    fp_add(s, a.c0, a.c1);
    fp_sub(d, a.c0, a.c1);
    fp_mul(m, a.c0, a.c1);
    fp_mul(r.c0, s, d);
    fp_dbl(r.c1, m);
}
static inline void fp2_mul_fp(Fp2 &r, const Fp2 &a, const Fp &b) { fp_mul(r.c0, a.c0, b); fp_mul(r.c1, a.c1, b); }
// times xi = 1 + u
static inline void fp2_mul_xi(Fp2 &r, const Fp2 &a) {
    Fp t0, t1;
    fp_sub(t0, a.c0, a.c1);
    fp_add(t1, a.c0, a.c1);
    r.c0 = t0; r.c1 = t1;
}
static void fp2_inv(Fp2 &r, const Fp2 &a) {
    Fp t0, t1;
    fp_sqr(t0, a.c0); fp_sqr(t1, a.c1); fp_add(t0, t0, t1);
    fp_inv(t0, t0);
    fp_mul(r.c0, a.c0, t0);
    fp_mul(t1, a.c1, t0);
    fp_neg(r.c1, t1);
}
static void fp2_pow(Fp2 &r, const Fp2 &a, const u64 *e, int nlimbs) {
    Fp2 acc = fp2_one();
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        fp2_sqr(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fp2_mul(acc, acc, a);
    }
    r = acc;
}
// square root for p = 3 (mod 4) (Adj & Rodriguez-Henriquez, algorithm 9); false if none
static bool fp2_sqrt(Fp2 &r, const Fp2 &a) {
    if (fp2_is_zero(a)) { r = a; return true; }
    Fp2 a1, alpha, x0, cand, chk;
    fp2_pow(a1, a, EXP_P_MINUS_3_DIV_4, 6);
    fp2_sqr(alpha, a1); fp2_mul(alpha, alpha, a);
    fp2_mul(x0, a1, a);
    Fp2 minus_one = fp2_one();
    fp_neg(minus_one.c0, minus_one.c0);
    if (fp2_eq(alpha, minus_one)) {
        // u * x0
        fp_neg(cand.c0, x0.c1);
        cand.c1 = x0.c0;
    } else {
        Fp2 b = fp2_one();
        fp2_add(b, b, alpha);
        fp2_pow(b, b, EXP_P_MINUS_1_DIV_2, 6);
        fp2_mul(cand, b, x0);
    }
    fp2_sqr(chk, cand);
    if (!fp2_eq(chk, a)) return false;
    r = cand;
    return true;
}

// ------------------------------------------------------------------ Fp6 = Fp2[v]/(v^3 - xi)
struct Fp6 { Fp2 c0, c1, c2; };
static inline void fp6_add(Fp6 &r, const Fp6 &a, const Fp6 &b) { fp2_add(r.c0, a.c0, b.c0); fp2_add(r.c1, a.c1, b.c1); fp2_add(r.c2, a.c2, b.c2); }
static inline void fp6_sub(Fp6 &r, const Fp6 &a, const Fp6 &b) { fp2_sub(r.c0, a.c0, b.c0); fp2_sub(r.c1, a.c1, b.c1); fp2_sub(r.c2, a.c2, b.c2); }
static inline void fp6_neg(Fp6 &r, const Fp6 &a) { fp2_neg(r.c0, a.c0); fp2_neg(r.c1, a.c1); fp2_neg(r.c2, a.c2); }
static void fp6_mul(Fp6 &r, const Fp6 &a, const Fp6 &b) {
    Fp2 v0, v1, v2, t0, t1, t2, s;
    fp2_mul(v0, a.c0, b.c0);
    fp2_mul(v1, a.c1, b.c1);
    fp2_mul(v2, a.c2, b.c2);
    // c0 = v0 + xi((a1+a2)(b1+b2) - v1 - v2)
    fp2_add(t0, a.c1, a.c2); fp2_add(s, b.c1, b.c2); fp2_mul(t0, t0, s); fp2_sub(t0, t0, v1); fp2_sub(t0, t0, v2);
    fp2_mul_xi(t0, t0); fp2_add(t0, t0, v0);
    // c1 = (a0+a1)(b0+b1) - v0 - v1 + xi v2
    fp2_add(t1, a.c0, a.c1); fp2_add(s, b.c0, b.c1); fp2_mul(t1, t1, s); fp2_sub(t1, t1, v0); fp2_sub(t1, t1, v1);
    fp2_mul_xi(s, v2); fp2_add(t1, t1, s);
    // c2 = (a0+a2)(b0+b2) - v0 - v2 + v1
    fp2_add(t2, a.c0, a.c2); fp2_add(s, b.c0, b.c2); fp2_mul(t2, t2, s); fp2_sub(t2, t2, v0); fp2_sub(t2, t2, v2);
    fp2_add(t2, t2, v1);
    r.c0 = t0; r.c1 = t1; r.c2 = t2;
}
static inline void fp6_sqr(Fp6 &r, const Fp6 &a) { fp6_mul(r, a, a); }
// times v
static inline void fp6_mul_v(Fp6 &r, const Fp6 &a) {
    Fp2 t;
    fp2_mul_xi(t, a.c2);
    r.c2 = a.c1; r.c1 = a.c0; r.c0 = t;
}
static void fp6_inv(Fp6 &r, const Fp6 &a) {
    Fp2 A, B, C, t, F;
    // A = a0^2 - xi a1 a2 ; B = xi a2^2 - a0 a1 ; C = a1^2 - a0 a2
    fp2_sqr(A, a.c0); fp2_mul(t, a.c1, a.c2); fp2_mul_xi(t, t); fp2_sub(A, A, t);
    fp2_sqr(B, a.c2); fp2_mul_xi(B, B); fp2_mul(t, a.c0, a.c1); fp2_sub(B, B, t);
    fp2_sqr(C, a.c1); fp2_mul(t, a.c0, a.c2); fp2_sub(C, C, t);
    // F = a0 A + xi (a2 B + a1 C)
    fp2_mul(F, a.c2, B); fp2_mul(t, a.c1, C); fp2_add(F, F, t); fp2_mul_xi(F, F);
    fp2_mul(t, a.c0, A); fp2_add(F, F, t);
    fp2_inv(F, F);
    fp2_mul(r.c0, A, F); fp2_mul(r.c1, B, F); fp2_mul(r.c2, C, F);
}
static Fp6 fp6_zero() { Fp6 r; memset(&r, 0, sizeof r); return r; }

// ------------------------------------------------------------------ Fp12 = Fp6[w]/(w^2 - v)
struct Fp12 { Fp6 a, b; };
static Fp12 fp12_one() { Fp12 r; memset(&r, 0, sizeof r); r.a.c0.c0 = fp_one(); return r; }
static void fp12_mul(Fp12 &r, const Fp12 &x, const Fp12 &y) {
    Fp6 aa, bb, s, t, o;
    fp6_mul(aa, x.a, y.a);
    fp6_mul(bb, x.b, y.b);
    fp6_add(s, x.a, x.b); fp6_add(t, y.a, y.b); fp6_mul(o, s, t); fp6_sub(o, o, aa); fp6_sub(o, o, bb);
    fp6_mul_v(t, bb);
    fp6_add(r.a, aa, t);
    r.b = o;
}
static void fp12_sqr(Fp12 &r, const Fp12 &x) {
    // (a + bw)^2 = (a^2 + v b^2) + 2ab w, with a^2 + v b^2 = (a+b)(a+vb) - ab - v ab
    Fp6 ab, s, t, vb;
    fp6_mul(ab, x.a, x.b);
    fp6_add(s, x.a, x.b);
    fp6_mul_v(vb, x.b); fp6_add(t, x.a, vb);
    fp6_mul(s, s, t);
    fp6_sub(s, s, ab);
    fp6_mul_v(t, ab); fp6_sub(s, s, t);
    r.a = s;
    fp6_add(r.b, ab, ab);
}
static inline void fp12_conj(Fp12 &r, const Fp12 &x) { r.a = x.a; fp6_neg(r.b, x.b); }
static void fp12_inv(Fp12 &r, const Fp12 &x) {
    Fp6 t0, t1;
    fp6_sqr(t0, x.a); fp6_sqr(t1, x.b); fp6_mul_v(t1, t1); fp6_sub(t0, t0, t1);
    fp6_inv(t0, t0);
    fp6_mul(r.a, x.a, t0);
    fp6_mul(t1, x.b, t0);
    fp6_neg(r.b, t1);
}
static bool fp12_is_one(const Fp12 &x) {
    Fp12 o = fp12_one();
    return memcmp(&o, &x, sizeof o) == 0;  // limbs are fully reduced, so equality is bitwise
}
// Frobenius x -> x^(p^k), k = 1, 2.  In the basis w^i (i = 0..5, coefficient i lives at
// (i odd ? b : a).c[i/2]) the map is c_i -> conj^k(c_i) * gamma_k^i with gamma_1 =
// xi^((p-1)/6) and gamma_2 = gamma_1 * conj(gamma_1).
static Fp2 FROB1[6], FROB2[6];
static inline Fp2 *coef(Fp12 &x, int i) { Fp6 &h = (i & 1) ? x.b : x.a; return i / 2 == 0 ? &h.c0 : (i / 2 == 1 ? &h.c1 : &h.c2); }
static void fp12_frob(Fp12 &r, const Fp12 &x, int k) {
    Fp12 t = x;
    for (int i = 0; i < 6; i++) {
        Fp2 *c = coef(t, i);
        if (k == 1) fp2_conj(*c, *c);
        fp2_mul(*c, *c, k == 1 ? FROB1[i] : FROB2[i]);
    }
    r = t;
}
// f * (l0 + l2 w^2 + l3 w^3) with l3 in Fp, i.e. a-part (l0, l2, 0), b-part (0, l3, 0)
static void fp12_mul_line(Fp12 &f, const Fp2 &l0, const Fp2 &l2, const Fp &l3) {
    auto mul_la = [&](Fp6 &r, const Fp6 &c) {
        Fp2 t0, t1, u;
        Fp6 o;
        fp2_mul(t0, c.c0, l0); fp2_mul(t1, c.c2, l2); fp2_mul_xi(t1, t1); fp2_add(o.c0, t0, t1);
        fp2_mul(t0, c.c0, l2); fp2_mul(t1, c.c1, l0); fp2_add(o.c1, t0, t1);
        fp2_mul(t0, c.c1, l2); fp2_mul(u, c.c2, l0); fp2_add(o.c2, t0, u);
        r = o;
    };
    auto mul_lb = [&](Fp6 &r, const Fp6 &c) {  // c * (l3 v)
        Fp6 o;
        Fp2 t;
        fp2_mul_fp(t, c.c2, l3); fp2_mul_xi(o.c0, t);
        fp2_mul_fp(o.c1, c.c0, l3);
        fp2_mul_fp(o.c2, c.c1, l3);
        r = o;
    };
    Fp6 aa, bb, ab, ba, t;
    mul_la(aa, f.a);
    mul_lb(bb, f.b);
    mul_lb(ab, f.a);
    mul_la(ba, f.b);
    fp6_mul_v(t, bb);
    fp6_add(f.a, aa, t);
    fp6_add(f.b, ab, ba);
}

// ------------------------------------------------------------------ G2 (affine) and prepared lines
struct G2 { Fp2 x, y; bool inf; };
struct Line { Fp2 lam, c; };  // y = lam x + c on the twist:  c = yT - lam xT
struct G2Prepared { std::vector<Line> lines; bool inf = true; };

static const u64 BLS_X = BLS_X_ABS;

static bool g2_uncompress(G2 &out, const uint8_t in[96]) {
    uint8_t b0 = in[0];
    if (!(b0 & 0x80)) return false;
    if (b0 & 0x40) {
        if (b0 & 0x3F) return false;
        for (int i = 1; i < 96; i++) if (in[i]) return false;
        memset(&out, 0, sizeof out);
        out.inf = true;
        return true;
    }
    uint8_t tmp[48];
    memcpy(tmp, in, 48);
    tmp[0] &= 0x1F;
    Fp2 x;
    if (!fp_from_be48(x.c1, tmp)) return false;
    if (!fp_from_be48(x.c0, in + 48)) return false;
    Fp2 rhs, y, b;
    fp2_sqr(rhs, x); fp2_mul(rhs, rhs, x);
    Fp four = fp_one(); fp_dbl(four, four); fp_dbl(four, four);
    b.c0 = four; b.c1 = four;  // 4(1 + u)
    fp2_add(rhs, rhs, b);
    if (!fp2_sqrt(y, rhs)) return false;
    bool large = fp_is_zero(y.c1) ? fp_is_large(y.c0) : fp_is_large(y.c1);
    if (large != ((b0 & 0x20) != 0)) fp2_neg(y, y);
    out.x = x; out.y = y; out.inf = false;
    return true;
}
// the line sequence of f_{|x|,Q}: 63 doubling steps, an addition step after each set bit
static void g2_prepare(G2Prepared &out, const G2 &Q) {
    out.lines.clear();
    out.inf = Q.inf;
    if (Q.inf) return;
    Fp2 tx = Q.x, ty = Q.y;
    for (int i = 62; i >= 0; i--) {
        Fp2 lam, t, u, nx, ny;
        fp2_sqr(u, tx); fp2_dbl(lam, u); fp2_add(lam, lam, u);
        fp2_dbl(t, ty); fp2_inv(t, t);
        fp2_mul(lam, lam, t);
        Line l; l.lam = lam; fp2_mul(t, lam, tx); fp2_sub(l.c, ty, t);
        out.lines.push_back(l);
        fp2_sqr(nx, lam); fp2_sub(nx, nx, tx); fp2_sub(nx, nx, tx);
        fp2_sub(t, tx, nx); fp2_mul(t, lam, t); fp2_sub(ny, t, ty);
        tx = nx; ty = ny;
        if ((BLS_X >> i) & 1) {
            fp2_sub(t, Q.x, tx); fp2_inv(t, t);
            fp2_sub(u, Q.y, ty);
            fp2_mul(lam, u, t);
            l.lam = lam; fp2_mul(t, lam, tx); fp2_sub(l.c, ty, t);
            out.lines.push_back(l);
            fp2_sqr(nx, lam); fp2_sub(nx, nx, tx); fp2_sub(nx, nx, Q.x);
            fp2_sub(t, tx, nx); fp2_mul(t, lam, t); fp2_sub(ny, t, ty);
            tx = nx; ty = ny;
        }
    }
}

struct G1 { Fp x, y; bool inf; };
// 96-byte uncompressed affine (x || y big-endian; byte 0 bit 6 = infinity), as phase B writes it
static bool g1_from_uncompressed(G1 &out, const uint8_t in[96]) {
    if (in[0] & 0x40) { memset(&out, 0, sizeof out); out.inf = true; return true; }
    if (!fp_from_be48(out.x, in) || !fp_from_be48(out.y, in + 48)) return false;
    out.inf = false;
    return true;
}
static bool g1_uncompress(G1 &out, const uint8_t in[48]) {
    uint8_t b0 = in[0];
    if (!(b0 & 0x80)) return false;
    if (b0 & 0x40) {
        if (b0 & 0x3F) return false;
        for (int i = 1; i < 48; i++) if (in[i]) return false;
        memset(&out, 0, sizeof out);
        out.inf = true;
        return true;
    }
    uint8_t tmp[48];
    memcpy(tmp, in, 48);
    tmp[0] &= 0x1F;
    Fp x, rhs, y, four;
    if (!fp_from_be48(x, tmp)) return false;
    fp_sqr(rhs, x); fp_mul(rhs, rhs, x);
    four = fp_one(); fp_dbl(four, four); fp_dbl(four, four);
    fp_add(rhs, rhs, four);
    if (!fp_sqrt(y, rhs)) return false;
    if (fp_is_large(y) != ((b0 & 0x20) != 0)) fp_neg(y, y);
    out.x = x; out.y = y; out.inf = false;
    return true;
}

// f_{|x|,Q1}(P1) * f_{|x|,Q2}(P2) with one shared squaring per step.  The line through T
// with slope lam, evaluated at P and scaled by w^3, is (c) + (lam xP) w^2 - yP w^3... with the
// sign convention below (c = yT - lam xT).
static void miller_pair(Fp12 &f, const G2Prepared &q1, const G1 &p1, const G2Prepared &q2, const G1 &p2) {
    f = fp12_one();
    const bool use1 = !q1.inf && !p1.inf, use2 = !q2.inf && !p2.inf;
    Fp ny1, ny2;
    if (use1) fp_neg(ny1, p1.y);
    if (use2) fp_neg(ny2, p2.y);
    size_t k = 0;
    auto step = [&](size_t idx) {
        Fp2 l2;
        if (use1) { fp2_mul_fp(l2, q1.lines[idx].lam, p1.x); fp12_mul_line(f, q1.lines[idx].c, l2, ny1); }
        if (use2) { fp2_mul_fp(l2, q2.lines[idx].lam, p2.x); fp12_mul_line(f, q2.lines[idx].c, l2, ny2); }
    };
    for (int i = 62; i >= 0; i--) {
        fp12_sqr(f, f);
        step(k++);
        if ((BLS_X >> i) & 1) step(k++);
    }
}
// Squaring in the cyclotomic subgroup (where everything lives after the easy part of the final exponentiation):
// Granger-Scott, "Faster squaring in the cyclotomic subgroup of sixth degree extensions" -- three squarings in
// Fp4 = Fp2[y]/(y^2 - xi) instead of a full Fp12 squaring (18 base-field products instead of 36).
static inline void fp4_sqr(Fp2 &r0, Fp2 &r1, const Fp2 &a, const Fp2 &b) {  // (a + b y)^2 = (a^2 + xi b^2) + 2ab y
    Fp2 ab, s, t, xb;
    fp2_mul(ab, a, b);
    fp2_add(s, a, b);
    fp2_mul_xi(xb, b); fp2_add(t, a, xb);
    fp2_mul(s, s, t);
    fp2_sub(s, s, ab);
    fp2_mul_xi(t, ab); fp2_sub(r0, s, t);
    fp2_dbl(r1, ab);
}
static void fp12_cyclotomic_sqr(Fp12 &r, const Fp12 &x) {
    Fp2 z0 = x.a.c0, z4 = x.a.c1, z3 = x.a.c2, z2 = x.b.c0, z1 = x.b.c1, z5 = x.b.c2;
    Fp2 t0, t1, t2, t3, t4, t5, tmp;
    fp4_sqr(t0, t1, z0, z1);
    fp4_sqr(t2, t3, z2, z3);
    fp4_sqr(t4, t5, z4, z5);
    auto three_minus_two = [](Fp2 &z, const Fp2 &t) { fp2_sub(z, t, z); fp2_dbl(z, z); fp2_add(z, z, t); };  // 3t - 2z
    auto three_plus_two = [](Fp2 &z, const Fp2 &t) { fp2_add(z, t, z); fp2_dbl(z, z); fp2_add(z, z, t); };    // 3t + 2z
    three_minus_two(z0, t0);
    three_plus_two(z1, t1);
    fp2_mul_xi(tmp, t5);
    three_plus_two(z2, tmp);
    three_minus_two(z3, t4);
    three_minus_two(z4, t2);
    three_plus_two(z5, t3);
    r.a.c0 = z0; r.a.c1 = z4; r.a.c2 = z3; r.b.c0 = z2; r.b.c1 = z1; r.b.c2 = z5;
}
static void fp12_pow_x(Fp12 &r, const Fp12 &a) {  // a^x for the (negative) curve parameter; a in the cyclotomic subgroup
    Fp12 acc = a;
    for (int i = 62; i >= 0; i--) {
        fp12_cyclotomic_sqr(acc, acc);
        if ((BLS_X >> i) & 1) fp12_mul(acc, acc, a);
    }
    fp12_conj(r, acc);
}
static bool final_exp_is_one(const Fp12 &f) {
    Fp12 g, t0, t1, t2, t3, u;
    // easy part: g = f^((p^6-1)(p^2+1))
    fp12_inv(t0, f);
    fp12_conj(g, f);
    fp12_mul(g, g, t0);
    fp12_frob(t0, g, 2);
    fp12_mul(g, g, t0);
    // hard part (cubed): g^((x-1)^2 (x+p) (x^2+p^2-1)) * g^3
    fp12_pow_x(t0, g); fp12_conj(u, g); fp12_mul(t0, t0, u);          // g^(x-1)
    fp12_pow_x(t1, t0); fp12_conj(u, t0); fp12_mul(t1, t1, u);        // ^(x-1)
    fp12_pow_x(t2, t1); fp12_frob(u, t1, 1); fp12_mul(t2, t2, u);     // ^(x+p)
    fp12_pow_x(t3, t2); fp12_pow_x(t3, t3); fp12_frob(u, t2, 2); fp12_mul(t3, t3, u);
    fp12_conj(u, t2); fp12_mul(t3, t3, u);                            // ^(x^2+p^2-1)
    fp12_sqr(u, g); fp12_mul(u, u, g);
    fp12_mul(t3, t3, u);
    return fp12_is_one(t3);
}

static G2Prepared GEN_PREPARED;
static std::once_flag init_flag;
static void init_consts() {
    const uint32_t p[12] = {FP_P_LIMBS}, r1[12] = {FP_R1_LIMBS}, r2[12] = {FP_R2_LIMBS};
    const uint32_t e1[12] = {FP_P_PLUS_1_DIV_4_LIMBS}, e2[12] = {FP_P_MINUS_3_DIV_4_LIMBS}, e3[12] = {FP_P_MINUS_1_DIV_2_LIMBS},
                   e4[12] = {FP_P_MINUS_2_LIMBS};
    pack(P64, p, 6); pack(R1_64, r1, 6); pack(R2_64, r2, 6);
    pack(EXP_P_PLUS_1_DIV_4, e1, 6); pack(EXP_P_MINUS_3_DIV_4, e2, 6); pack(EXP_P_MINUS_1_DIV_2, e3, 6); pack(EXP_P_MINUS_2, e4, 6);
    u64 inv = 1;  // Newton: inv = p^-1 mod 2^64
    for (int i = 0; i < 6; i++) inv *= 2 - P64[0] * inv;
    N0_64 = (u64)0 - inv;
    // (p - 1) / 6 by long division
    u64 pm1[6];
    memcpy(pm1, P64, sizeof pm1);
    pm1[0] -= 1;
    u128 rem = 0;
    for (int i = 5; i >= 0; i--) { u128 cur = (rem << 64) | pm1[i]; EXP_P_MINUS_1_DIV_6[i] = (u64)(cur / 6); rem = cur % 6; }
    Fp2 xi = fp2_one(), g1;
    xi.c1 = fp_one();
    fp2_pow(g1, xi, EXP_P_MINUS_1_DIV_6, 6);
    Fp2 g2, cj;
    fp2_conj(cj, g1);
    fp2_mul(g2, g1, cj);
    FROB1[0] = FROB2[0] = fp2_one();
    for (int i = 1; i < 6; i++) { fp2_mul(FROB1[i], FROB1[i - 1], g1); fp2_mul(FROB2[i], FROB2[i - 1], g2); }
    // generator of G2
    const uint32_t x0[12] = {G2_GEN_X0_MONT_LIMBS}, x1[12] = {G2_GEN_X1_MONT_LIMBS}, y0[12] = {G2_GEN_Y0_MONT_LIMBS},
                   y1[12] = {G2_GEN_Y1_MONT_LIMBS};
    G2 gen;
    pack(gen.x.c0.l, x0, 6); pack(gen.x.c1.l, x1, 6); pack(gen.y.c0.l, y0, 6); pack(gen.y.c1.l, y1, 6);
    gen.inf = false;
    g2_prepare(GEN_PREPARED, gen);
}

static bool pairing_eq(const G1 &a1, const G2Prepared &a2, const G1 &b1, const G2Prepared &b2) {
    G1 na = a1;
    if (!na.inf) fp_neg(na.y, na.y);
    Fp12 f;
    miller_pair(f, a2, na, b2, b1);
    return final_exp_is_one(f);
}

}  // namespace

struct host_g2_prepared { G2Prepared p; };

// ------------------------------------------------------------------ G1 on the host (Jacobian; only for the few
// additions and the one generator multiplication that close a batch verification)
namespace {
struct G1J { Fp x, y, z; };  // z == 0: infinity
static bool g1j_is_inf(const G1J &p) { return fp_is_zero(p.z); }
static void g1j_from_affine(G1J &r, const G1 &a) {
    if (a.inf) { memset(&r, 0, sizeof r); return; }
    r.x = a.x; r.y = a.y; r.z = fp_one();
}
static void g1j_dbl(G1J &r, const G1J &p) {
    if (g1j_is_inf(p)) { r = p; return; }
    Fp A, B, C, D, E, F, t, z3;
    fp_sqr(A, p.x); fp_sqr(B, p.y); fp_sqr(C, B);
    fp_add(t, p.x, B); fp_sqr(t, t); fp_sub(t, t, A); fp_sub(t, t, C); fp_dbl(D, t);
    fp_dbl(E, A); fp_add(E, E, A);
    fp_sqr(F, E);
    fp_mul(z3, p.y, p.z); fp_dbl(z3, z3);
    fp_dbl(t, D); fp_sub(r.x, F, t);
    fp_sub(t, D, r.x); fp_mul(t, E, t);
    fp_dbl(C, C); fp_dbl(C, C); fp_dbl(C, C);
    fp_sub(r.y, t, C);
    r.z = z3;
}
static void g1j_add(G1J &r, const G1J &p, const G1J &q) {
    if (g1j_is_inf(p)) { r = q; return; }
    if (g1j_is_inf(q)) { r = p; return; }
    Fp z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    fp_sqr(z1z1, p.z); fp_sqr(z2z2, q.z);
    fp_mul(u1, p.x, z2z2); fp_mul(u2, q.x, z1z1);
    fp_mul(s1, p.y, q.z); fp_mul(s1, s1, z2z2);
    fp_mul(s2, q.y, p.z); fp_mul(s2, s2, z1z1);
    if (fp_eq(u1, u2)) {
        if (fp_eq(s1, s2)) { g1j_dbl(r, p); return; }
        memset(&r, 0, sizeof r);
        return;
    }
    fp_sub(h, u2, u1);
    fp_dbl(i, h); fp_sqr(i, i);
    fp_mul(j, h, i);
    fp_sub(rr, s2, s1); fp_dbl(rr, rr);
    fp_mul(v, u1, i);
    G1J o;
    fp_sqr(o.x, rr); fp_sub(o.x, o.x, j); fp_dbl(t, v); fp_sub(o.x, o.x, t);
    fp_sub(t, v, o.x); fp_mul(t, rr, t);
    fp_mul(s1, s1, j); fp_dbl(s1, s1);
    fp_sub(o.y, t, s1);
    fp_add(t, p.z, q.z); fp_sqr(t, t); fp_sub(t, t, z1z1); fp_sub(t, t, z2z2);
    fp_mul(o.z, t, h);
    r = o;
}
static void g1j_to_affine(G1 &r, const G1J &p) {
    if (g1j_is_inf(p)) { memset(&r, 0, sizeof r); r.inf = true; return; }
    Fp zi, zi2, zi3;
    fp_inv(zi, p.z);
    fp_sqr(zi2, zi);
    fp_mul(zi3, zi2, zi);
    fp_mul(r.x, p.x, zi2);
    fp_mul(r.y, p.y, zi3);
    r.inf = false;
}
}  // namespace

int host_verify_finish(const uint8_t *partials, size_t k, const host_g2_prepared *tau, int *ok) {
    std::call_once(init_flag, init_consts);
    if (!tau) return 1;
    G1J A, B;
    memset(&A, 0, sizeof A);
    memset(&B, 0, sizeof B);
    // s = sum of the shard scalars mod r (4 x 64-bit, canonical)
    const uint32_t r32[8] = {FR_R_LIMBS};
    u64 rmod[4], s[4] = {0, 0, 0, 0};
    pack(rmod, r32, 4);
    for (size_t i = 0; i < k; i++) {
        const uint8_t *rec = partials + 224 * i;
        G1 a, b;
        if (!g1_from_uncompressed(a, rec) || !g1_from_uncompressed(b, rec + 96)) return 1;
        G1J t;
        g1j_from_affine(t, a); g1j_add(A, A, t);
        g1j_from_affine(t, b); g1j_add(B, B, t);
        u64 v[4];
        for (int w = 0; w < 4; w++) {
            u64 x = 0;
            for (int b8 = 0; b8 < 8; b8++) x = (x << 8) | rec[192 + 8 * (3 - w) + b8];
            v[w] = x;
        }
        auto geq = [&](const u64 *x) { for (int w = 3; w >= 0; w--) if (x[w] != rmod[w]) return x[w] > rmod[w]; return true; };
        if (geq(v)) return 1;
        u128 c = 0;
        for (int w = 0; w < 4; w++) { c += (u128)s[w] + v[w]; s[w] = (u64)c; c >>= 64; }
        if (c || geq(s)) {
            u128 bw = 0;
            for (int w = 0; w < 4; w++) { u128 d = (u128)s[w] - rmod[w] - (u64)bw; s[w] = (u64)d; bw = (d >> 64) & 1; }
        }
    }
    // B - [s]G1: fixed-base, 4-bit windows over a table of k 2^(4w) G1 (64 x 15 Jacobian points, built once per process):
    // 64 additions instead of a 255-step double-and-add
    static std::once_flag gen_flag;
    static std::vector<G1J> gen_table;
    std::call_once(gen_flag, [] {
        const uint32_t gx[12] = {G1_GEN_X_MONT_LIMBS}, gy[12] = {G1_GEN_Y_MONT_LIMBS};
        G1J base;
        pack(base.x.l, gx, 6); pack(base.y.l, gy, 6); base.z = fp_one();
        gen_table.resize(64 * 15);
        for (int w = 0; w < 64; w++) {
            gen_table[15 * w] = base;
            for (int k = 1; k < 15; k++) g1j_add(gen_table[15 * w + k], gen_table[15 * w + k - 1], base);
            for (int d = 0; d < 4; d++) g1j_dbl(base, base);
        }
    });
    G1J acc;
    memset(&acc, 0, sizeof acc);
    for (int w = 0; w < 64; w++) {
        const unsigned d = (unsigned)(s[w >> 4] >> (4 * (w & 15))) & 15u;
        if (d) g1j_add(acc, acc, gen_table[15 * w + d - 1]);
    }
    fp_neg(acc.y, acc.y);
    g1j_add(B, B, acc);
    G1 a, b;
    g1j_to_affine(a, A);
    g1j_to_affine(b, B);
    *ok = pairing_eq(a, tau->p, b, GEN_PREPARED) ? 1 : 0;
    return 0;
}


host_g2_prepared *host_g2_prepare(const uint8_t g2_compressed[96]) {
    std::call_once(init_flag, init_consts);
    G2 q;
    if (!g2_uncompress(q, g2_compressed)) return nullptr;
    host_g2_prepared *h = new host_g2_prepared();
    g2_prepare(h->p, q);
    return h;
}
void host_g2_prepared_free(host_g2_prepared *h) { delete h; }

int host_check_setup(const uint8_t *g1_bytes, const uint8_t *g2_bytes, size_t n2) {
    std::call_once(init_flag, init_consts);
    if (n2 < 2) return 1;
    G2 q0, q1, q;
    for (size_t i = 0; i < n2; i++) {
        if (!g2_uncompress(q, g2_bytes + 96 * i)) return 1;
        if (i == 0) q0 = q;
        if (i == 1) q1 = q;
    }
    G1 p0, p1;
    if (!g1_uncompress(p0, g1_bytes) || !g1_uncompress(p1, g1_bytes + 48)) return 1;
    G2Prepared a, b;
    g2_prepare(a, q0);
    g2_prepare(b, q1);
    // monomial form <=> e(g1[1], g2[0]) == e(g1[0], g2[1])  (reference src/kzg.rs:817-829)
    if (pairing_eq(p1, a, p0, b)) return 1;
    return 0;
}

int host_pairing_check_uncompressed(const uint8_t a1[96], const host_g2_prepared *a2, const uint8_t b1[96],
                                    const host_g2_prepared *b2_or_null_for_generator, int *ok) {
    std::call_once(init_flag, init_consts);
    G1 pa, pb;
    if (!a2 || !g1_from_uncompressed(pa, a1) || !g1_from_uncompressed(pb, b1)) return 1;
    const G2Prepared &qb = b2_or_null_for_generator ? b2_or_null_for_generator->p : GEN_PREPARED;
    *ok = pairing_eq(pa, a2->p, pb, qb) ? 1 : 0;
    return 0;
}

int host_pairings_verify(const uint8_t a1[48], const uint8_t a2[96], const uint8_t b1[48], const uint8_t b2[96], int *ok) {
    std::call_once(init_flag, init_consts);
    G1 pa, pb;
    G2 qa, qb;
    if (!g1_uncompress(pa, a1) || !g1_uncompress(pb, b1) || !g2_uncompress(qa, a2) || !g2_uncompress(qb, b2)) return 1;
    G2Prepared ra, rb;
    g2_prepare(ra, qa);
    g2_prepare(rb, qb);
    *ok = pairing_eq(pa, ra, pb, rb) ? 1 : 0;
    return 0;
}

// test hook (tests/test_host_logic.py): host_verify_finish for a given [tau]G2, no context (and so no GPU) needed
extern "C" int kzg_b200_host_verify_finish_with_tau(const uint8_t tau_g2[96], const uint8_t *partials, size_t k, int *ok) {
    host_g2_prepared *h = host_g2_prepare(tau_g2);
    if (!h) return 1;
    int rc = host_verify_finish(partials, k, h, ok);
    host_g2_prepared_free(h);
    return rc;
}

// bigint.cuh -- Montgomery field arithmetic on 32-bit limbs for sm_100a (and, for the
// host-side parts of the library and for CPU unit tests, the same code on the host).
//
// Replaces, for the blob path, the blst field calls listed in SURVEY.md section 2b:
// blst_fr_add/sub/mul/sqr, blst_fr_eucl_inverse, blst_fr_from_scalar/blst_scalar_from_fr
// (reference src/utils.rs:13-123, 230-275) and the Fp arithmetic underneath every
// blst_p1_* call (reference src/utils.rs:126-140, 221-227, 282-410).
//
// Device path: carry chains of mad.lo.cc / madc.hi.cc (IMAD on the FMA pipe) written as
// inline PTX.  A 32x32->64 product is two IMAD issue slots; one Fp multiplication is
// 2*(12*12) + 2*(12*12) + 12 = 588 of them, one Fr multiplication 2*64 + 2*64 + 8 = 264.
// Products are accumulated into two interleaved accumulators ("even"/"odd" 64-bit
// columns) so that no carry ever has to ripple further than one chain.
//
// Host path: each PTX primitive has a C++ stand-in with an explicit carry variable, so the
// algorithms are byte-for-byte the same and can be checked on a CPU.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define KZG_HD __host__ __device__ __forceinline__
#define KZG_D __device__ __forceinline__
#else
#define KZG_HD inline
#define KZG_D inline
#endif

namespace kzg {

// ------------------------------------------------------------------ carry primitives
// `cc` is the emulated carry flag on the host; on the device the hardware flag is used and
// the argument is dead.
#if defined(__CUDA_ARCH__)
KZG_HD uint32_t add_cc(uint32_t a, uint32_t b, uint32_t &) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_HD uint32_t addc_cc(uint32_t a, uint32_t b, uint32_t &) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_HD uint32_t addc(uint32_t a, uint32_t b, uint32_t &) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_HD uint32_t sub_cc(uint32_t a, uint32_t b, uint32_t &) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_HD uint32_t subc_cc(uint32_t a, uint32_t b, uint32_t &) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_HD uint32_t subc(uint32_t a, uint32_t b, uint32_t &) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t &) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KZG_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t &) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KZG_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t &) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KZG_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
#else
KZG_HD uint32_t add_cc(uint32_t a, uint32_t b, uint32_t &cc) { uint64_t s = (uint64_t)a + b; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
KZG_HD uint32_t addc_cc(uint32_t a, uint32_t b, uint32_t &cc) { uint64_t s = (uint64_t)a + b + cc; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
KZG_HD uint32_t addc(uint32_t a, uint32_t b, uint32_t &cc) { return a + b + cc; }
KZG_HD uint32_t sub_cc(uint32_t a, uint32_t b, uint32_t &cc) { uint64_t s = (uint64_t)a - b; cc = (uint32_t)(s >> 63); return (uint32_t)s; }
KZG_HD uint32_t subc_cc(uint32_t a, uint32_t b, uint32_t &cc) { uint64_t s = (uint64_t)a - b - cc; cc = (uint32_t)(s >> 63); return (uint32_t)s; }
KZG_HD uint32_t subc(uint32_t a, uint32_t b, uint32_t &cc) { return a - b - cc; }
KZG_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t &cc) { uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
KZG_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t &cc) { uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c + cc; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
KZG_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t &cc) { uint64_t s = (((uint64_t)a * b) >> 32) + c + cc; cc = (uint32_t)(s >> 32); return (uint32_t)s; }
KZG_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
#endif

KZG_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c, uint32_t &cc) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return (uint32_t)((((uint64_t)a * b) >> 32) + c + cc);
#endif
}
KZG_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// 64-bit columns.  x is read with stride 2 (x[0], x[2], ...): pass `a` for the even limbs
// of an operand and `a + 1` for the odd ones.  N = limbs of the accumulator (even).
//   mul_row : acc[2k+1:2k]  = x[2k] * m
//   cmad_row: acc[2k+1:2k] += x[2k] * m  in one carry chain; the carry out is left in the flag
template <int N> KZG_HD void mul_row(uint32_t *acc, const uint32_t *x, uint32_t m) {
#pragma unroll
    for (int k = 0; k < N; k += 2) { acc[k] = mul_lo(x[k], m); acc[k + 1] = mul_hi(x[k], m); }
}
template <int N> KZG_HD void cmad_row(uint32_t *acc, const uint32_t *x, uint32_t m, uint32_t &cc) {
    acc[0] = mad_lo_cc(x[0], m, acc[0], cc);
    acc[1] = madc_hi_cc(x[0], m, acc[1], cc);
#pragma unroll
    for (int k = 2; k < N; k += 2) {
        acc[k] = madc_lo_cc(x[k], m, acc[k], cc);
        acc[k + 1] = madc_hi_cc(x[k], m, acc[k + 1], cc);
    }
}

// ------------------------------------------------------------------ the field template
// P supplies: N (limbs, even), mod(i), n0, r1(i) (Montgomery one), r2(i) (R^2 mod m).
template <class P>
struct alignas(16) Fe {
    static constexpr int N = P::N;
    uint32_t l[N];
};

template <class P> KZG_HD bool fe_is_zero(const Fe<P> &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < P::N; i++) o |= a.l[i];
    return o == 0;
}
template <class P> KZG_HD bool fe_eq(const Fe<P> &a, const Fe<P> &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < P::N; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}
template <class P> KZG_HD void fe_set_zero(Fe<P> &a) {
#pragma unroll
    for (int i = 0; i < P::N; i++) a.l[i] = 0;
}
template <class P> KZG_HD Fe<P> fe_one() {
    Fe<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.l[i] = P::r1(i);
    return r;
}
// r = a - mod if a >= mod (a < 2*mod, plus an optional incoming carry bit)
template <class P> KZG_HD void fe_reduce_once(Fe<P> &a, uint32_t top) {
    uint32_t d[P::N], cc = 0;
    d[0] = sub_cc(a.l[0], P::mod(0), cc);
#pragma unroll
    for (int i = 1; i < P::N; i++) d[i] = subc_cc(a.l[i], P::mod(i), cc);
    uint32_t borrow = subc(top, 0, cc);  // top - borrow: 0xffffffff iff (top:a) < mod
    bool keep = (borrow != 0) && (top == 0);
    // keep a when a < mod: that is borrow out with top == 0
#pragma unroll
    for (int i = 0; i < P::N; i++) a.l[i] = keep ? a.l[i] : d[i];
}
template <class P> KZG_HD void fe_add(Fe<P> &r, const Fe<P> &a, const Fe<P> &b) {
    uint32_t cc = 0;
    Fe<P> t;
    t.l[0] = add_cc(a.l[0], b.l[0], cc);
#pragma unroll
    for (int i = 1; i < P::N; i++) t.l[i] = addc_cc(a.l[i], b.l[i], cc);
    uint32_t top = addc(0, 0, cc);
    fe_reduce_once(t, top);
    r = t;
}
template <class P> KZG_HD void fe_sub(Fe<P> &r, const Fe<P> &a, const Fe<P> &b) {
    uint32_t cc = 0;
    Fe<P> t;
    t.l[0] = sub_cc(a.l[0], b.l[0], cc);
#pragma unroll
    for (int i = 1; i < P::N; i++) t.l[i] = subc_cc(a.l[i], b.l[i], cc);
    uint32_t borrow = subc(0, 0, cc);  // 0 or 0xffffffff
    uint32_t c2 = 0;
    r.l[0] = add_cc(t.l[0], P::mod(0) & borrow, c2);
#pragma unroll
    for (int i = 1; i < P::N; i++) r.l[i] = addc_cc(t.l[i], P::mod(i) & borrow, c2);
}
template <class P> KZG_HD void fe_neg(Fe<P> &r, const Fe<P> &a) {
    Fe<P> z;
    fe_set_zero(z);
    fe_sub(r, z, a);
}
template <class P> KZG_HD void fe_dbl(Fe<P> &r, const Fe<P> &a) { fe_add(r, a, a); }

// Montgomery product, operand scanning with the reduction interleaved (CIOS), arranged for
// the 64-bit multiply-add of the hardware (IMAD.WIDE): the running value is kept as
//     V = E + 2^32 * O,     E = sum ev[k] 2^(32k),  O = sum od[k] 2^(32k)      (N limbs each)
// so that x[even]*m lands on E's 64-bit columns and x[odd]*m on O's.  Each step adds
// a*b_i and m*mod (m chosen so that V becomes divisible by 2^32) and divides by 2^32:
// O becomes the new E, E >> 64 the new O, and the stray limb E[1] is added at the bottom
// of the new E with its carry entering the new O's chain.
// Bounds (mod < 2^(32N-2)): V < 2 mod before a step and < 2 mod 2^32 inside it, hence
// O < 2^(32N) always -- no chain on O ever carries out; a chain on E may, and that carry
// (weight 2^(32N)) is added to O's top limb.  Inputs < mod, output < mod.
// REDUCE = false skips the final conditional subtraction ("lazy" form, see fe_mul_lazy below).
template <class P, bool REDUCE> KZG_HD void fe_mul_core(Fe<P> &r, const Fe<P> &a, const Fe<P> &b) {
    constexpr int N = P::N;
    uint32_t mp[N];
#pragma unroll
    for (int i = 0; i < N; i++) mp[i] = P::mod(i);
    uint32_t x[N], y[N];  // x plays E and y plays O in even steps; swapped in odd steps
    uint32_t cc = 0;
    // step 0
    mul_row<N>(x, a.l, b.l[0]);
    mul_row<N>(y, a.l + 1, b.l[0]);
    {
        uint32_t m = mul_lo(x[0], P::n0);
        cmad_row<N>(y, mp + 1, m, cc);
        cmad_row<N>(x, mp, m, cc);
        y[N - 1] = addc(y[N - 1], 0, cc);
    }
#pragma unroll
    for (int i = 1; i < N; i++) {
        uint32_t *E = (i & 1) ? y : x;  // the accumulator that is E during this step
        uint32_t *O = (i & 1) ? x : y;  // old E, becomes O after the shift by 64 bits
        const uint32_t bi = b.l[i];
        E[0] = add_cc(E[0], O[1], cc);
        // O = (O >> 64) + a_odd * b_i + carry
#pragma unroll
        for (int k = 0; k < N - 2; k += 2) {
            O[k] = madc_lo_cc(a.l[k + 1], bi, O[k + 2], cc);
            O[k + 1] = madc_hi_cc(a.l[k + 1], bi, O[k + 3], cc);
        }
        O[N - 2] = madc_lo_cc(a.l[N - 1], bi, 0, cc);
        O[N - 1] = madc_hi(a.l[N - 1], bi, 0, cc);
        cmad_row<N>(E, a.l, bi, cc);
        O[N - 1] = addc(O[N - 1], 0, cc);
        uint32_t m = mul_lo(E[0], P::n0);
        cmad_row<N>(O, mp + 1, m, cc);
        cmad_row<N>(E, mp, m, cc);
        O[N - 1] = addc(O[N - 1], 0, cc);
    }
    // result = O + (E >> 32), E[0] == 0.  After step N-1 (odd): E = y, O = x.
    uint32_t *E = ((N - 1) & 1) ? y : x;
    uint32_t *O = ((N - 1) & 1) ? x : y;
    Fe<P> t;
    t.l[0] = add_cc(O[0], E[1], cc);
#pragma unroll
    for (int k = 1; k < N - 1; k++) t.l[k] = addc_cc(O[k], E[k + 1], cc);
    t.l[N - 1] = addc(O[N - 1], 0, cc);
    if (REDUCE) fe_reduce_once(t, 0);
    r = t;
}
template <class P> KZG_HD void fe_mul(Fe<P> &r, const Fe<P> &a, const Fe<P> &b) { fe_mul_core<P, true>(r, a, b); }
template <class P> KZG_HD void fe_sqr(Fe<P> &r, const Fe<P> &a) { fe_mul(r, a, a); }

// Two independent products at once: r1 = a1 b1, r2 = a2 b2, the carry chains of the two interleaved chain by chain.
// One product is a dependent sequence of ~290 multiply-adds, and a lone warp issues it at half the rate of the
// multiplier (the next instruction of a chain waits for the previous one); a second, independent chain in the gaps
// nearly doubles what a latency-bound thread gets done -- the Jacobian formulas of the Horner pass, the scalar
// ladders and the subgroup checks pair their products through this (g1.cuh).  Same steps as fe_mul_core.
template <class P> KZG_HD void mul2_step_a(uint32_t *E, uint32_t *O, const uint32_t *a, uint32_t bi) {
    constexpr int N = P::N;
    uint32_t cc = 0;
    E[0] = add_cc(E[0], O[1], cc);
#pragma unroll
    for (int k = 0; k < N - 2; k += 2) {
        O[k] = madc_lo_cc(a[k + 1], bi, O[k + 2], cc);
        O[k + 1] = madc_hi_cc(a[k + 1], bi, O[k + 3], cc);
    }
    O[N - 2] = madc_lo_cc(a[N - 1], bi, 0, cc);
    O[N - 1] = madc_hi(a[N - 1], bi, 0, cc);
}
template <class P> KZG_HD void mul2_step_b(uint32_t *E, uint32_t *O, const uint32_t *a, uint32_t bi) {
    uint32_t cc = 0;
    cmad_row<P::N>(E, a, bi, cc);
    O[P::N - 1] = addc(O[P::N - 1], 0, cc);
}
template <class P> KZG_HD uint32_t mul2_step_c(uint32_t *E, uint32_t *O, const uint32_t *mp) {
    uint32_t cc = 0;
    uint32_t m = mul_lo(E[0], P::n0);
    cmad_row<P::N>(O, mp + 1, m, cc);
    return m;
}
template <class P> KZG_HD void mul2_step_d(uint32_t *E, uint32_t *O, const uint32_t *mp, uint32_t m) {
    uint32_t cc = 0;
    cmad_row<P::N>(E, mp, m, cc);
    O[P::N - 1] = addc(O[P::N - 1], 0, cc);
}
template <class P> KZG_HD void mul2_finish(Fe<P> &r, const uint32_t *E, const uint32_t *O) {
    constexpr int N = P::N;
    uint32_t cc = 0;
    Fe<P> t;
    t.l[0] = add_cc(O[0], E[1], cc);
#pragma unroll
    for (int k = 1; k < N - 1; k++) t.l[k] = addc_cc(O[k], E[k + 1], cc);
    t.l[N - 1] = addc(O[N - 1], 0, cc);
    fe_reduce_once(t, 0);
    r = t;
}
template <class P> KZG_HD void fe_mul2(Fe<P> &r1, const Fe<P> &a1, const Fe<P> &b1, Fe<P> &r2, const Fe<P> &a2, const Fe<P> &b2) {
    constexpr int N = P::N;
    uint32_t mp[N];
#pragma unroll
    for (int i = 0; i < N; i++) mp[i] = P::mod(i);
    uint32_t x1[N], y1[N], x2[N], y2[N];
    // step 0
    mul_row<N>(x1, a1.l, b1.l[0]);
    mul_row<N>(x2, a2.l, b2.l[0]);
    mul_row<N>(y1, a1.l + 1, b1.l[0]);
    mul_row<N>(y2, a2.l + 1, b2.l[0]);
    {
        uint32_t m1 = mul2_step_c<P>(x1, y1, mp);
        uint32_t m2 = mul2_step_c<P>(x2, y2, mp);
        mul2_step_d<P>(x1, y1, mp, m1);
        mul2_step_d<P>(x2, y2, mp, m2);
    }
#pragma unroll
    for (int i = 1; i < N; i++) {
        uint32_t *E1 = (i & 1) ? y1 : x1, *O1 = (i & 1) ? x1 : y1;
        uint32_t *E2 = (i & 1) ? y2 : x2, *O2 = (i & 1) ? x2 : y2;
        mul2_step_a<P>(E1, O1, a1.l, b1.l[i]);
        mul2_step_a<P>(E2, O2, a2.l, b2.l[i]);
        mul2_step_b<P>(E1, O1, a1.l, b1.l[i]);
        mul2_step_b<P>(E2, O2, a2.l, b2.l[i]);
        uint32_t m1 = mul2_step_c<P>(E1, O1, mp);
        uint32_t m2 = mul2_step_c<P>(E2, O2, mp);
        mul2_step_d<P>(E1, O1, mp, m1);
        mul2_step_d<P>(E2, O2, mp, m2);
    }
    Fe<P> t1, t2;  // the outputs may alias the inputs
    mul2_finish<P>(t1, ((N - 1) & 1) ? y1 : x1, ((N - 1) & 1) ? x1 : y1);
    mul2_finish<P>(t2, ((N - 1) & 1) ? y2 : x2, ((N - 1) & 1) ? x2 : y2);
    r1 = t1;
    r2 = t2;
}

// ------------------------------------------------------------------ lazy residues: values in [0, 2 mod)
// Every instruction beside the multiplier costs issue slots, so the hot loop of
// the MSM keeps its values only partly reduced.  With R = 2^(32N) > 4 mod (Fp: R / mod = 9.8) the
// Montgomery product of a, b < 2 mod is (a b + M mod) / R < 4 mod^2 / R + mod < 2 mod without any final
// subtraction; inside the loop V stays below 3 mod 2^32, so the accumulator bounds of fe_mul_core hold.
// Subtraction adds 2 mod back on a borrow.  Values are made canonical where they leave the loop.
template <class P> KZG_HD void fe_mul_lazy(Fe<P> &r, const Fe<P> &a, const Fe<P> &b) { fe_mul_core<P, false>(r, a, b); }
template <class P> KZG_HD void fe_sub_lazy(Fe<P> &r, const Fe<P> &a, const Fe<P> &b) {
    uint32_t cc = 0;
    Fe<P> t;
    t.l[0] = sub_cc(a.l[0], b.l[0], cc);
#pragma unroll
    for (int i = 1; i < P::N; i++) t.l[i] = subc_cc(a.l[i], b.l[i], cc);
    uint32_t borrow = subc(0, 0, cc);  // 0 or 0xffffffff
    uint32_t c2 = 0;
    r.l[0] = add_cc(t.l[0], P::mod2(0) & borrow, c2);
#pragma unroll
    for (int i = 1; i < P::N; i++) r.l[i] = addc_cc(t.l[i], P::mod2(i) & borrow, c2);
}
// r = a - b + 2 mod with no condition at all: a, b in [0, 2 mod) give r in (0, 4 mod).  For differences that
// only feed a product whose other factor is below 2 mod -- (4 mod)(2 mod)/R + mod < 2 mod for R > 8.2 mod (Fp:
// R / mod = 9.8), and the accumulator of fe_mul_core stays below 5 mod 2^32 -- so the masked add-back of
// fe_sub_lazy (12 of its 36 instructions) is not needed.  The difference may wrap below zero in between.
template <class P> KZG_HD void fe_sub_lazy4(Fe<P> &r, const Fe<P> &a, const Fe<P> &b) {
    uint32_t cc = 0;
    Fe<P> t;
    t.l[0] = sub_cc(a.l[0], b.l[0], cc);
#pragma unroll
    for (int i = 1; i < P::N; i++) t.l[i] = subc_cc(a.l[i], b.l[i], cc);
    uint32_t c2 = 0;
    r.l[0] = add_cc(t.l[0], P::mod2(0), c2);
#pragma unroll
    for (int i = 1; i < P::N; i++) r.l[i] = addc_cc(t.l[i], P::mod2(i), c2);
}
// a == 0 (mod m) for a in (0, 4 mod): a is mod, 2 mod or 3 mod; the low limb filters almost everything
template <class P> KZG_HD bool fe_is_zero_lazy4(const Fe<P> &a) {
    const uint32_t m0 = P::mod(0);
    if (a.l[0] != m0 && a.l[0] != 2u * m0 && a.l[0] != 3u * m0) return false;
    Fe<P> c = a;
    // c in (0, 4 mod): bring it below 2 mod, then below mod
    {
        uint32_t d[P::N], cc = 0;
        d[0] = sub_cc(c.l[0], P::mod2(0), cc);
#pragma unroll
        for (int i = 1; i < P::N; i++) d[i] = subc_cc(c.l[i], P::mod2(i), cc);
        uint32_t borrow = subc(0, 0, cc);
        if (borrow == 0) {
#pragma unroll
            for (int i = 0; i < P::N; i++) c.l[i] = d[i];
        }
    }
    fe_canonical(c);
    return fe_is_zero(c);
}
// r = 2 mod - a for a in (0, 2 mod]: one carry chain where the canonical fe_neg takes three
template <class P> KZG_HD void fe_neg_lazy(Fe<P> &r, const Fe<P> &a) {
    uint32_t cc = 0;
    r.l[0] = sub_cc(P::mod2(0), a.l[0], cc);
#pragma unroll
    for (int i = 1; i < P::N; i++) r.l[i] = subc_cc(P::mod2(i), a.l[i], cc);
}
// [0, 2 mod) -> [0, mod)
template <class P> KZG_HD void fe_canonical(Fe<P> &a) { fe_reduce_once(a, 0); }
// a == 0 (mod m) for a in [0, 2 mod): a is 0 or mod.  The low limb filters out all but 2^-31 of the values.
template <class P> KZG_HD bool fe_is_zero_lazy(const Fe<P> &a) {
    if (a.l[0] != 0 && a.l[0] != P::mod(0)) return false;
    Fe<P> c = a;
    fe_canonical(c);
    return fe_is_zero(c);
}

template <class P> KZG_HD void fe_to_mont(Fe<P> &r, const Fe<P> &a) {
    Fe<P> r2;
#pragma unroll
    for (int i = 0; i < P::N; i++) r2.l[i] = P::r2(i);
    fe_mul(r, a, r2);
}
template <class P> KZG_HD void fe_from_mont(Fe<P> &r, const Fe<P> &a) {
    Fe<P> one;
    fe_set_zero(one);
    one.l[0] = 1;
    fe_mul(r, a, one);
}
// r = a^e, e given as NE little-endian 32-bit words (not unrolled: code size)
template <class P> KZG_HD void fe_pow(Fe<P> &r, const Fe<P> &a, const uint32_t *e, int ne) {
    Fe<P> acc = fe_one<P>();
    bool started = false;
#pragma unroll 1
    for (int i = ne * 32 - 1; i >= 0; i--) {
        if (started) fe_sqr(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) {
            if (started) fe_mul(acc, acc, a);
            else { acc = a; started = true; }
        }
    }
    r = acc;
}
// ------------------------------------------------------------------ inversion by the binary extended Euclid
// Fermat's a^(m-2) is ~1.5 N*32 dependent multiplications (570 for Fp); the binary algorithm
// (HAC 14.61) needs ~2 N*32 rounds of shifts, additions and subtractions on N limbs -- a few
// times fewer instructions, none of them on the multiplier.  Invariants: b x = u, c x = v
// (mod m); u, v odd after the halvings; gcd(x, m) = 1 so one of them reaches 1.
// Montgomery in, Montgomery out: the integer inverse of aR is a^-1 R^-1; two products with
// R^2 bring it back to a^-1 R.  a must not be zero.
template <int N> KZG_HD void limbs_shr1(uint32_t *a, uint32_t top) {
#pragma unroll
    for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[N - 1] = (a[N - 1] >> 1) | (top << 31);
}
template <class P> KZG_HD void halve_mod(uint32_t *b) {  // b = b / 2 mod m, b < m
    uint32_t mask = 0u - (b[0] & 1u), cc = 0, t[P::N];
    t[0] = add_cc(b[0], P::mod(0) & mask, cc);
#pragma unroll
    for (int i = 1; i < P::N; i++) t[i] = addc_cc(b[i], P::mod(i) & mask, cc);
    uint32_t top = addc(0, 0, cc);
#pragma unroll
    for (int i = 0; i < P::N; i++) b[i] = t[i];
    limbs_shr1<P::N>(b, top);
}
template <class P> KZG_HD void fe_inv_binary(Fe<P> &r, const Fe<P> &a) {
    constexpr int N = P::N;
    uint32_t u[N], v[N];
    Fe<P> b, c;
    fe_set_zero(b);
    fe_set_zero(c);
    b.l[0] = 1;
#pragma unroll
    for (int i = 0; i < N; i++) { u[i] = a.l[i]; v[i] = P::mod(i); }
    auto is_one = [](const uint32_t *x) {
        uint32_t o = x[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; i++) o |= x[i];
        return o == 0;
    };
#pragma unroll 1
    while (!is_one(u) && !is_one(v)) {
#pragma unroll 1
        while (!(u[0] & 1u)) { limbs_shr1<N>(u, 0); halve_mod<P>(b.l); }
#pragma unroll 1
        while (!(v[0] & 1u)) { limbs_shr1<N>(v, 0); halve_mod<P>(c.l); }
        // d = u - v; keep it when u >= v
        uint32_t d[N], cc = 0;
        d[0] = sub_cc(u[0], v[0], cc);
#pragma unroll
        for (int i = 1; i < N; i++) d[i] = subc_cc(u[i], v[i], cc);
        uint32_t borrow = subc(0, 0, cc);
        if (borrow == 0) {
#pragma unroll
            for (int i = 0; i < N; i++) u[i] = d[i];
            fe_sub(b, b, c);
        } else {
            cc = 0;
            v[0] = sub_cc(v[0], u[0], cc);
#pragma unroll
            for (int i = 1; i < N; i++) v[i] = subc_cc(v[i], u[i], cc);
            fe_sub(c, c, b);
        }
    }
    Fe<P> t = is_one(u) ? b : c, r2;
#pragma unroll
    for (int i = 0; i < N; i++) r2.l[i] = P::r2(i);
    fe_mul(t, t, r2);
    fe_mul(r, t, r2);
}

// canonical (non-Montgomery) a >= b ?
template <int N> KZG_HD bool limbs_geq(const uint32_t *a, const uint32_t *b) {
    uint32_t cc = 0;
    (void)sub_cc(a[0], b[0], cc);
#pragma unroll
    for (int i = 1; i < N; i++) (void)subc_cc(a[i], b[i], cc);
    uint32_t borrow = subc(0, 0, cc);
    return borrow == 0;
}

}  // namespace kzg

// fp_hybrid.cuh -- Fp multiplication split over TWO multiplier pipes of the B200 SM.
//
// The IMAD pipe (16 lanes per SM sub-partition, one 32x32+64 multiply-add per 4 clocks per
// warp) is the roofline of the all-integer fe_mul (bigint.cuh): 288 wide multiply-adds per
// product.  The B200 also has a full-rate FP64 pipe (tools/ubench.cu: 1.85e13 DFMA/s, and it
// co-issues with IMAD at 1.77e13 pairs/s) that the blob path otherwise leaves idle.  Here
//
//   * the 768-bit product a*b is formed on the FP64 pipe: operands as 8 limbs of 48 bits held
//     exactly in doubles; each 48x48 partial product is split exactly into a high part (bits
//     >= 2^48, accumulated for free in an FMA chain that rounds toward zero inside the binade
//     [2^100, 2^101), ulp 2^48) and a low part (< 2^48, one more FMA that cancels the high
//     part, exact), as in Emmart/Luitjens/Weems, "Optimizing modular multiplication for NVIDIA's
//     Maxwell GPUs" (ARITH 2016) and Emmart/Zheng/Weems, "Faster modular exponentiation using
//     double precision floating point arithmetic on the GPU" (ARITH 2018);
//   * the Montgomery reduction (by the fixed modulus, R = 2^384 as everywhere else, so values
//     are interchangeable with fe_mul's) stays on the IMAD pipe: 144 wide multiply-adds.
//
// Per product: 156 IMAD-pipe + ~270 FP64-pipe + ~200 ALU instructions instead of 312 + 0 + ~40:
// the multiplier work is spread over two pipes and the result is bit-identical to fe_mul
// (canonical residue < p).  Squaring needs 36 partial products instead of 64.
//
// Host build: the same code with fma() under FE_TOWARDZERO (tests/hostshim), so the limb
// algebra is checked against Python integers on a CPU.
#pragma once
#include "../../kzg_rust_b200/csrc/fields.cuh"

#if !defined(__CUDA_ARCH__)
#include <cfenv>
#include <cmath>
#endif

namespace kzg {

#if defined(__CUDA_ARCH__)
KZG_D double hy_make(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
KZG_D uint32_t hy_hi(double d) { return (uint32_t)__double2hiint(d); }
KZG_D uint32_t hy_lo(double d) { return (uint32_t)__double2loint(d); }
KZG_D double hy_fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
KZG_D double hy_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
KZG_D double hy_sub(double a, double b) { return __dsub_rn(a, b); }
KZG_D double hy_add(double a, double b) { return __dadd_rn(a, b); }
#else
inline double hy_make(uint32_t hi, uint32_t lo) { uint64_t v = ((uint64_t)hi << 32) | lo; double d; memcpy(&d, &v, 8); return d; }
inline uint32_t hy_hi(double d) { uint64_t v; memcpy(&v, &d, 8); return (uint32_t)(v >> 32); }
inline uint32_t hy_lo(double d) { uint64_t v; memcpy(&v, &d, 8); return (uint32_t)v; }
inline double hy_fma_rz(double a, double b, double c) {
    volatile double va = a, vb = b, vc = c;
    int old = fegetround();
    fesetround(FE_TOWARDZERO);
    volatile double r = std::fma(va, vb, vc);
    fesetround(old);
    return r;
}
inline double hy_fma(double a, double b, double c) { volatile double va = a, vb = b, vc = c; volatile double r = std::fma(va, vb, vc); return r; }
inline double hy_sub(double a, double b) { volatile double va = a, vb = b; volatile double r = va - vb; return r; }
inline double hy_add(double a, double b) { volatile double va = a, vb = b; volatile double r = va + vb; return r; }
#endif

// 12 x 32-bit limbs -> 8 x 48-bit limbs, each an exact double
KZG_HD void hy_limbs48(double *v, const fp_t &a) {
    const double two52 = 4503599627370496.0;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        uint32_t w0 = a.l[3 * m], w1 = a.l[3 * m + 1], w2 = a.l[3 * m + 2];
        v[2 * m] = hy_sub(hy_make((w1 & 0xffffu) | 0x43300000u, w0), two52);
        v[2 * m + 1] = hy_sub(hy_make((w2 >> 16) | 0x43300000u, (w1 >> 16) | (w2 << 16)), two52);
    }
}

// One column of the schoolbook product: sum over the listed (x, y) pairs, split exactly as
//   sum = 2^48 * h + l,   h = sum of floor(x y / 2^48) < 2^51,   l = sum of (x y mod 2^48) < 2^51
// A runs in [2^100, 2^101): fma_rz drops exactly the bits below 2^48 of each product (the
// running value is a multiple of the ulp), so A - 2^100 accumulates the high parts with no
// extra instruction; A_old - A_new = -high part, and fma(x, y, that) is the low part, exact.
struct HyColumn {
    double A, L;
    KZG_HD void start() { A = 1267650600228229401496703205376.0 /* 2^100 */; L = 4503599627370496.0 /* 2^52 */; }
    KZG_HD void mac(double x, double y) {
        double An = hy_fma_rz(x, y, A);
        double D = hy_sub(A, An);
        double lo = hy_fma(x, y, D);
        L = hy_add(L, lo);
        A = An;
    }
};

// columns -> 24 x 32-bit limbs.  V_k = l_k + h_(k-1) < 2^52 sits at bit 48 k: even k are
// word-aligned (3 words apart), odd k are 16 bits off; the two interleaved halves are added once.
KZG_HD void hy_assemble(uint32_t *T, const uint32_t *vlo, const uint32_t *vhi /* 16 values, hi < 2^20 */) {
    uint32_t e[24], o[24];
#pragma unroll
    for (int m = 0; m < 8; m++) {
        e[3 * m] = vlo[2 * m];
        e[3 * m + 1] = vhi[2 * m];
        e[3 * m + 2] = 0;
        uint32_t lo = vlo[2 * m + 1], hi = vhi[2 * m + 1];
        o[3 * m + 1] = lo << 16;
        o[3 * m + 2] = (lo >> 16) | (hi << 16);
        if (m < 7) o[3 * m + 3] = hi >> 16;  // the top column is < 2^42: nothing above word 23
    }
    o[0] = 0;
    uint32_t cc = 0;
    T[0] = e[0];
    T[1] = add_cc(e[1], o[1], cc);
#pragma unroll
    for (int i = 2; i < 24; i++) T[i] = addc_cc(e[i], o[i], cc);
}

template <bool SQR>
KZG_HD void hy_product(uint32_t *T, const fp_t &a, const fp_t &b) {
    double x[8], y[8], x2[8];
    hy_limbs48(x, a);
    if (SQR) {
#pragma unroll
        for (int i = 0; i < 8; i++) { y[i] = x[i]; x2[i] = hy_add(x[i], x[i]); }
    } else {
        hy_limbs48(y, b);
    }
    uint32_t vlo[16], vhi[16];
    uint32_t plo = 0, phi = 0;  // h of the previous column (mantissa bits of its A)
#pragma unroll
    for (int k = 0; k < 15; k++) {
        HyColumn c;
        c.start();
        const int i0 = k > 7 ? k - 7 : 0, i1 = k < 7 ? k : 7;
        if (SQR) {
            // 2 x_i x_j for i < j, x_i^2 on the diagonal; 4 doubled terms (< 2^97 each) + one
            // square keep A below 2^101
#pragma unroll
            for (int i = i0; i <= i1; i++) {
                int j = k - i;
                if (i < j) c.mac(x2[i], y[j]);
                else if (i == j) c.mac(x[i], y[j]);
            }
        } else {
#pragma unroll
            for (int i = i0; i <= i1; i++) c.mac(x[i], y[k - i]);
        }
        // V_k = l_k + h_(k-1): add the raw encodings; the mantissa fields cannot overflow (< 2^52)
        uint32_t cc = 0;
        uint32_t llo = hy_lo(c.L), lhi = hy_hi(c.L);
        vlo[k] = add_cc(llo, plo, cc);
        vhi[k] = addc(lhi, phi, cc) & 0xfffffu;
        plo = hy_lo(c.A);
        phi = hy_hi(c.A) & 0xfffffu;
    }
    vlo[15] = plo;
    vhi[15] = phi;
    hy_assemble(T, vlo, vhi);
}

// Montgomery reduction of a 2N-limb value T < mod * 2^(32N):  r = T / 2^(32N) mod m, canonical.
// Same even/odd 64-bit-column scheme as fe_mul (bigint.cuh) without the operand rows:
// U = (T_lo + M mod) / R is built in N steps of one IMAD (m), N wide multiply-adds (m * mod) and
// a shift; the result is U + T_hi, reduced once.
template <class P, bool REDUCE = true> KZG_HD void fe_redc(Fe<P> &r, const uint32_t *T) {
    constexpr int N = P::N;
    uint32_t mp[N];
#pragma unroll
    for (int i = 0; i < N; i++) mp[i] = P::mod(i);
    uint32_t x[N], y[N];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = T[i]; y[i] = 0; }
    uint32_t cc = 0;
    {
        uint32_t m = mul_lo(x[0], P::n0);
        cmad_row<N>(y, mp + 1, m, cc);
        cmad_row<N>(x, mp, m, cc);
        y[N - 1] = addc(y[N - 1], 0, cc);
    }
#pragma unroll
    for (int i = 1; i < N; i++) {
        uint32_t *E = (i & 1) ? y : x;
        uint32_t *O = (i & 1) ? x : y;
        E[0] = add_cc(E[0], O[1], cc);  // stray limb of the shifted accumulator; its carry enters O's chain
        uint32_t m = mul_lo(E[0], P::n0);
#pragma unroll
        for (int k = 0; k < N - 2; k += 2) {
            O[k] = madc_lo_cc(mp[k + 1], m, O[k + 2], cc);
            O[k + 1] = madc_hi_cc(mp[k + 1], m, O[k + 3], cc);
        }
        O[N - 2] = madc_lo_cc(mp[N - 1], m, 0, cc);
        O[N - 1] = madc_hi(mp[N - 1], m, 0, cc);
        cmad_row<N>(E, mp, m, cc);
        O[N - 1] = addc(O[N - 1], 0, cc);
    }
    uint32_t *E = ((N - 1) & 1) ? y : x;
    uint32_t *O = ((N - 1) & 1) ? x : y;
    Fe<P> t;
    t.l[0] = add_cc(O[0], E[1], cc);
#pragma unroll
    for (int k = 1; k < N - 1; k++) t.l[k] = addc_cc(O[k], E[k + 1], cc);
    t.l[N - 1] = addc(O[N - 1], 0, cc);
    t.l[0] = add_cc(t.l[0], T[N], cc);
#pragma unroll
    for (int k = 1; k < N; k++) t.l[k] = addc_cc(t.l[k], T[N + k], cc);
    uint32_t top = addc(0, 0, cc);
    if (REDUCE) fe_reduce_once(t, top);  // lazy form: T < 4 mod^2 gives (T + M mod) / R < 2 mod < 2^(32N), no carry out
    r = t;
}

KZG_HD void fp_mul_hybrid(fp_t &r, const fp_t &a, const fp_t &b) {
    uint32_t T[24];
    hy_product<false>(T, a, b);
    fe_redc(r, T);
}
KZG_HD void fp_sqr_hybrid(fp_t &r, const fp_t &a) {
    uint32_t T[24];
    hy_product<true>(T, a, a);
    fe_redc(r, T);
}
// a in [0, 2p) -> a^2 / R in [0, 2p)
KZG_HD void fp_sqr_hybrid_lazy(fp_t &r, const fp_t &a) {
    uint32_t T[24];
    hy_product<true>(T, a, a);
    fe_redc<FpParams, false>(r, T);
}

}  // namespace kzg

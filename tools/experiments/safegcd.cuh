// safegcd.cuh -- modular inversion by Bernstein-Yang divsteps, measured as a replacement of the binary extended Euclid
// (fe_inv_binary, csrc/bigint.cuh) and NOT kept in the product: on the GPU its 64-bit signed accumulators are emulated
// (mul.wide + two-instruction adds and shifts), so the instruction count is not the 3x lower it is on a CPU --
// blob_to_kzg_commitment of one blob 0.741 -> 0.726 ms, but the Horner pass of 65,536 blobs (one inversion per thread)
// 14.0 -> 16.0 ms (profiles/criterion_r2ax.jsonl, bench_r2ax.json).  Bit-identical with the binary algorithm and Python
// (tests/test_host_logic.py::test_divstep_inversion_vs_python); 483 GPU tests passed with it.
#pragma once
#include "../../kzg_rust_b200/csrc/fields.cuh"

namespace kzg {

// ------------------------------------------------------------------ inversion by divsteps ("safegcd")
// Bernstein-Yang divsteps in the half-delta form, 30 at a time (the layout of libsecp256k1's modinv32, restated for these
// moduli): the low 30 bits of f and g decide 30 steps and give a 2 x 2 transition matrix t with t (f, g) = 2^30 (f', g');
// t is then applied once to the full-width f, g (exact division by 2^30) and to d, e (division mod m), which keep
// d x = f, e x = g (mod m).  Signed limbs of 30 bits in int32; f = m, g = x at the start, g = 0 and f = +-1 at the end,
// d = +-1/x.  The binary algorithm (bigint.cuh) costs ~76,000 dependent instructions for Fp (120 us on one GPU thread -- a
// visible part of every small call: the conversion to affine after the Horner pass, the sums of phase B); this one
// ~25,000.  Steps: the half-delta bound is floor((45907 bits + 26313) / 19929) divsteps -- 879 for 381 bits, 589 for 255 --
// rounded up to whole batches plus one (extra steps with g = 0 change nothing).
template <class P> struct SafeGcd {
    static constexpr int LEN = (32 * P::N + 29) / 30;          // 13 limbs for Fp (390 bits), 9 for Fr (270 bits)
    static constexpr int BITS = 32 * P::N;                      // bound on the inputs (m < 2^BITS)
    static constexpr int BATCHES = ((45907 * BITS + 26313) / 19929 + 29) / 30 + 1;
    KZG_HD static constexpr int32_t mod30(int i) {              // limb i of the modulus in radix 2^30
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        uint64_t v = (uint64_t)(w < P::N ? P::mod(w) : 0u) | ((uint64_t)(w + 1 < P::N ? P::mod(w + 1) : 0u) << 32);
        return (int32_t)((v >> sh) & 0x3fffffffu);
    }
    KZG_HD static constexpr uint32_t inv30() {                  // m^-1 mod 2^30 (Newton)
        uint32_t m = P::mod(0), x = 1;
        for (int i = 0; i < 5; i++) x *= 2u - m * x;
        return x & 0x3fffffffu;
    }
};
struct DivstepMatrix { int32_t u, v, q, r; };
// 30 divsteps on the low limbs; zeta = -(delta + 1/2)
KZG_HD int32_t divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, DivstepMatrix &t) {
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 1
    for (int i = 0; i < 30; i++) {
        uint32_t c1 = (uint32_t)(zeta >> 31);  // all ones when zeta < 0
        const uint32_t c2 = 0u - (g & 1u);     // all ones when g is odd
        const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;  // +-(f, u, v)
        g += x & c2;
        q += y & c2;
        r += z & c2;
        c1 &= c2;                                // zeta < 0 and g odd: swap
        zeta = (int32_t)(((uint32_t)zeta ^ c1) - 1u);
        f += g & c1;
        u += q & c1;
        v += r & c1;
        g >>= 1;
        u <<= 1;
        v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return zeta;
}
template <class P> KZG_HD void safegcd_update_fg(int32_t *f, int32_t *g, const DivstepMatrix &t) {
    constexpr int LEN = SafeGcd<P>::LEN;
    const int32_t M30 = 0x3fffffff;
    int64_t cf = (int64_t)t.u * f[0] + (int64_t)t.v * g[0];
    int64_t cg = (int64_t)t.q * f[0] + (int64_t)t.r * g[0];
    cf >>= 30;  // the low 30 bits are zero
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < LEN; i++) {
        const int32_t fi = f[i], gi = g[i];
        cf += (int64_t)t.u * fi + (int64_t)t.v * gi;
        cg += (int64_t)t.q * fi + (int64_t)t.r * gi;
        f[i - 1] = (int32_t)cf & M30; cf >>= 30;
        g[i - 1] = (int32_t)cg & M30; cg >>= 30;
    }
    f[LEN - 1] = (int32_t)cf;
    g[LEN - 1] = (int32_t)cg;
}
// (d, e) <- t (d, e) / 2^30 mod m, d and e in (-2m, m)
template <class P> KZG_HD void safegcd_update_de(int32_t *d, int32_t *e, const DivstepMatrix &t) {
    constexpr int LEN = SafeGcd<P>::LEN;
    const int32_t M30 = 0x3fffffff;
    const int32_t sd = d[LEN - 1] >> 31, se = e[LEN - 1] >> 31;  // sign masks
    int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
    int64_t cd = (int64_t)t.u * d[0] + (int64_t)t.v * e[0];
    int64_t ce = (int64_t)t.q * d[0] + (int64_t)t.r * e[0];
    // multiples of m that clear the low 30 bits
    md -= (int32_t)((SafeGcd<P>::inv30() * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)((SafeGcd<P>::inv30() * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += (int64_t)SafeGcd<P>::mod30(0) * md;
    ce += (int64_t)SafeGcd<P>::mod30(0) * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < LEN; i++) {
        const int32_t di = d[i], ei = e[i];
        cd += (int64_t)t.u * di + (int64_t)t.v * ei;
        ce += (int64_t)t.q * di + (int64_t)t.r * ei;
        cd += (int64_t)SafeGcd<P>::mod30(i) * md;
        ce += (int64_t)SafeGcd<P>::mod30(i) * me;
        d[i - 1] = (int32_t)cd & M30; cd >>= 30;
        e[i - 1] = (int32_t)ce & M30; ce >>= 30;
    }
    d[LEN - 1] = (int32_t)cd;
    e[LEN - 1] = (int32_t)ce;
}
// d in (-2m, m), negated when sign < 0  ->  [0, m), limbs in [0, 2^30)
template <class P> KZG_HD void safegcd_normalize(int32_t *r, int32_t sign) {
    constexpr int LEN = SafeGcd<P>::LEN;
    const int32_t M30 = 0x3fffffff;
    int32_t cond_add = r[LEN - 1] >> 31;
    const int32_t cond_neg = sign >> 31;
#pragma unroll
    for (int i = 0; i < LEN; i++) {
        r[i] += SafeGcd<P>::mod30(i) & cond_add;
        r[i] = (r[i] ^ cond_neg) - cond_neg;
    }
#pragma unroll
    for (int i = 0; i < LEN - 1; i++) { r[i + 1] += r[i] >> 30; r[i] &= M30; }
    cond_add = r[LEN - 1] >> 31;
#pragma unroll
    for (int i = 0; i < LEN; i++) r[i] += SafeGcd<P>::mod30(i) & cond_add;
#pragma unroll
    for (int i = 0; i < LEN - 1; i++) { r[i + 1] += r[i] >> 30; r[i] &= M30; }
}
// Montgomery in, Montgomery out (like fe_inv_binary).  a must not be zero mod m.
template <class P> KZG_HD void fe_inv_safegcd(Fe<P> &r, const Fe<P> &a) {
    constexpr int N = P::N, LEN = SafeGcd<P>::LEN;
    int32_t d[LEN], e[LEN], f[LEN], g[LEN];
#pragma unroll
    for (int i = 0; i < LEN; i++) {
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        const uint64_t v = (uint64_t)(w < N ? a.l[w] : 0u) | ((uint64_t)(w + 1 < N ? a.l[w + 1] : 0u) << 32);
        g[i] = (int32_t)((v >> sh) & 0x3fffffffu);
        f[i] = SafeGcd<P>::mod30(i);
        d[i] = 0;
        e[i] = 0;
    }
    e[0] = 1;
    int32_t zeta = -1;
#pragma unroll 1
    for (int it = 0; it < SafeGcd<P>::BATCHES; it++) {
        DivstepMatrix t;
        zeta = divsteps_30(zeta, (uint32_t)f[0], (uint32_t)g[0], t);
        safegcd_update_de<P>(d, e, t);
        safegcd_update_fg<P>(f, g, t);
    }
    safegcd_normalize<P>(d, f[LEN - 1]);
    Fe<P> t, r2;
#pragma unroll
    for (int w = 0; w < N; w++) {  // 30-bit limbs -> 32-bit words
        const int bit = 32 * w, i = bit / 30, sh = bit - 30 * i;
        uint64_t v = (uint64_t)(uint32_t)d[i] >> sh;
        if (i + 1 < LEN) v |= (uint64_t)(uint32_t)d[i + 1] << (30 - sh);
        if (i + 2 < LEN) v |= (uint64_t)(uint32_t)d[i + 2] << (60 - sh);
        t.l[w] = (uint32_t)v;
    }
#pragma unroll
    for (int i = 0; i < N; i++) r2.l[i] = P::r2(i);
    fe_mul(t, t, r2);
    fe_mul(r, t, r2);
}

}  // namespace kzg

// fields.cuh -- the two BLS12-381 fields as instances of the Montgomery template.
//   Fp : 381-bit base field, 12 x 32-bit limbs (coordinates of G1 points)
//   Fr : 255-bit scalar field, 8 x 32-bit limbs (blob elements, reference `fr_t`)
// Constants come from tools/gen_consts.py (consts_gen.cuh).
#pragma once
#include "bigint.cuh"
#include "consts_gen.cuh"

namespace kzg {

struct FpParams {
    static constexpr int N = 12;
    static constexpr uint32_t n0 = FP_N0;
    KZG_HD static constexpr uint32_t mod(int i) { constexpr uint32_t v[N] = {FP_P_LIMBS}; return v[i]; }
    KZG_HD static constexpr uint32_t mod2(int i) {  // 2 p (fits: p < 2^381)
        constexpr uint32_t v[N] = {FP_P_LIMBS};
        return (v[i] << 1) | (i ? v[i - 1] >> 31 : 0u);
    }
    KZG_HD static constexpr uint32_t r1(int i) { constexpr uint32_t v[N] = {FP_R1_LIMBS}; return v[i]; }
    KZG_HD static constexpr uint32_t r2(int i) { constexpr uint32_t v[N] = {FP_R2_LIMBS}; return v[i]; }
};
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t n0 = FR_N0;
    KZG_HD static constexpr uint32_t mod(int i) { constexpr uint32_t v[N] = {FR_R_LIMBS}; return v[i]; }
    KZG_HD static constexpr uint32_t mod2(int i) {  // 2 r (fits: r < 2^255)
        constexpr uint32_t v[N] = {FR_R_LIMBS};
        return (v[i] << 1) | (i ? v[i - 1] >> 31 : 0u);
    }
    KZG_HD static constexpr uint32_t r1(int i) { constexpr uint32_t v[N] = {FR_R1_LIMBS}; return v[i]; }
    KZG_HD static constexpr uint32_t r2(int i) { constexpr uint32_t v[N] = {FR_R2_LIMBS}; return v[i]; }
};
typedef Fe<FpParams> fp_t;
typedef Fe<FrParams> fr_t;

// exponent tables: runtime-indexed, so they live in constant memory on the device
#if defined(__CUDACC__)
#define KZG_CONST_TABLE(name, ...) \
    static __device__ __constant__ uint32_t d_##name[] = {__VA_ARGS__}; \
    static const uint32_t h_##name[] = {__VA_ARGS__};
#else
#define KZG_CONST_TABLE(name, ...) static const uint32_t h_##name[] = {__VA_ARGS__};
#endif
#if defined(__CUDA_ARCH__)
#define KZG_TABLE(name) d_##name
#else
#define KZG_TABLE(name) h_##name
#endif
KZG_CONST_TABLE(fp_p_minus_2, FP_P_MINUS_2_LIMBS)
KZG_CONST_TABLE(fp_p_plus_1_div_4, FP_P_PLUS_1_DIV_4_LIMBS)
KZG_CONST_TABLE(fp_p_minus_1_div_2, FP_P_MINUS_1_DIV_2_LIMBS)
KZG_CONST_TABLE(fr_r_minus_2, FR_R_MINUS_2_LIMBS)

// ---- Fp helpers
KZG_HD void fp_inv(fp_t &r, const fp_t &a) { fe_inv_binary(r, a); }
KZG_HD void fp_inv_fermat(fp_t &r, const fp_t &a) { fe_pow(r, a, KZG_TABLE(fp_p_minus_2), 12); }
// sqrt for p = 3 (mod 4): candidate a^((p+1)/4); false when a is not a square
KZG_HD bool fp_sqrt(fp_t &r, const fp_t &a) {
    fp_t s, s2;
    fe_pow(s, a, KZG_TABLE(fp_p_plus_1_div_4), 12);
    fe_sqr(s2, s);
    r = s;
    return fe_eq(s2, a);
}
// blst's sign convention for compressed points: canonical y > (p-1)/2
KZG_HD bool fp_is_lexicographically_largest(const fp_t &mont) {
    fp_t c;
    fe_from_mont(c, mont);
    constexpr uint32_t half[12] = {FP_P_MINUS_1_DIV_2_LIMBS};
    // c > half  <=>  !(half >= c)
    uint32_t cc = 0;
    (void)sub_cc(half[0], c.l[0], cc);
#pragma unroll
    for (int i = 1; i < 12; i++) (void)subc_cc(half[i], c.l[i], cc);
    return subc(0, 0, cc) != 0;
}
KZG_HD fp_t fp_const_b() { fp_t r; constexpr uint32_t v[12] = {FP_B_MONT_LIMBS};
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = v[i];
    return r; }

// ---- Fr helpers
KZG_HD void fr_inv(fr_t &r, const fr_t &a) { fe_inv_binary(r, a); }
KZG_HD void fr_inv_fermat(fr_t &r, const fr_t &a) { fe_pow(r, a, KZG_TABLE(fr_r_minus_2), 8); }
// 32 big-endian bytes (as 8 big-endian words already byte-swapped to host order, most
// significant first) -> canonical little-endian limbs
KZG_HD bool fr_is_canonical(const fr_t &a) {
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = FrParams::mod(i);
    return !limbs_geq<8>(a.l, m);
}

}  // namespace kzg

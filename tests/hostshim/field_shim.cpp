// Host-side shim exposing the product's field templates (host code path of
// kzg_rust_b200/csrc/bigint.cuh) to ctypes, so tests can compare them with Python ints.
#include "../../kzg_rust_b200/csrc/fields.cuh"
#include "../../tools/experiments/fp_hybrid.cuh"
#include "../../tools/experiments/safegcd.cuh"
using namespace kzg;
extern "C" {
void shim_fp_mul2(const uint32_t *a1, const uint32_t *b1, const uint32_t *a2, const uint32_t *b2, uint32_t *r1, uint32_t *r2) {
    fp_t x1, y1, x2, y2; memcpy(x1.l, a1, 48); memcpy(y1.l, b1, 48); memcpy(x2.l, a2, 48); memcpy(y2.l, b2, 48);
    fe_mul2(x1, x1, y1, y2, x2, y2);  /* outputs alias inputs on purpose */
    memcpy(r1, x1.l, 48); memcpy(r2, y2.l, 48);
}
void shim_fp_mul(const uint32_t *a, const uint32_t *b, uint32_t *r) { fp_t x, y, z; memcpy(x.l, a, 48); memcpy(y.l, b, 48); fe_mul(z, x, y); memcpy(r, z.l, 48); }
void shim_fp_add(const uint32_t *a, const uint32_t *b, uint32_t *r) { fp_t x, y, z; memcpy(x.l, a, 48); memcpy(y.l, b, 48); fe_add(z, x, y); memcpy(r, z.l, 48); }
void shim_fp_sub(const uint32_t *a, const uint32_t *b, uint32_t *r) { fp_t x, y, z; memcpy(x.l, a, 48); memcpy(y.l, b, 48); fe_sub(z, x, y); memcpy(r, z.l, 48); }
void shim_fp_inv_safegcd(const uint32_t *a, uint32_t *r) { fp_t x, z; memcpy(x.l, a, 48); fe_inv_safegcd(z, x); memcpy(r, z.l, 48); }
void shim_fr_inv_safegcd(const uint32_t *a, uint32_t *r) { fr_t x, z; memcpy(x.l, a, 32); fe_inv_safegcd(z, x); memcpy(r, z.l, 32); }
void shim_fp_inv_binary(const uint32_t *a, uint32_t *r) { fp_t x, z; memcpy(x.l, a, 48); fe_inv_binary(z, x); memcpy(r, z.l, 48); }
void shim_fr_inv_binary(const uint32_t *a, uint32_t *r) { fr_t x, z; memcpy(x.l, a, 32); fe_inv_binary(z, x); memcpy(r, z.l, 32); }
void shim_fp_inv(const uint32_t *a, uint32_t *r) { fp_t x, z; memcpy(x.l, a, 48); fp_inv(z, x); memcpy(r, z.l, 48); }
int shim_fp_sqrt(const uint32_t *a, uint32_t *r) { fp_t x, z; memcpy(x.l, a, 48); bool ok = fp_sqrt(z, x); memcpy(r, z.l, 48); return ok; }
int shim_fp_large(const uint32_t *a) { fp_t x; memcpy(x.l, a, 48); return fp_is_lexicographically_largest(x); }
void shim_fr_mul(const uint32_t *a, const uint32_t *b, uint32_t *r) { fr_t x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32); fe_mul(z, x, y); memcpy(r, z.l, 32); }
void shim_fr_add(const uint32_t *a, const uint32_t *b, uint32_t *r) { fr_t x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32); fe_add(z, x, y); memcpy(r, z.l, 32); }
void shim_fr_sub(const uint32_t *a, const uint32_t *b, uint32_t *r) { fr_t x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32); fe_sub(z, x, y); memcpy(r, z.l, 32); }
void shim_fr_inv(const uint32_t *a, uint32_t *r) { fr_t x, z; memcpy(x.l, a, 32); fr_inv(z, x); memcpy(r, z.l, 32); }
void shim_fp_inv_fermat(const uint32_t *a, uint32_t *r) { fp_t x, z; memcpy(x.l, a, 48); fp_inv_fermat(z, x); memcpy(r, z.l, 48); }
void shim_fr_inv_fermat(const uint32_t *a, uint32_t *r) { fr_t x, z; memcpy(x.l, a, 32); fr_inv_fermat(z, x); memcpy(r, z.l, 32); }
int shim_fr_canonical(const uint32_t *a) { fr_t x; memcpy(x.l, a, 32); return fr_is_canonical(x); }
}
extern "C" {
void shim_fp_mul_many(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t count) {
    for (size_t i = 0; i < count; i++) shim_fp_mul(a + 12 * i, b + 12 * i, r + 12 * i);
}
void shim_fr_mul_many(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t count) {
    for (size_t i = 0; i < count; i++) shim_fr_mul(a + 8 * i, b + 8 * i, r + 8 * i);
}
}
// the two-pipe (FP64 product + IMAD reduction) multiplication of fp_hybrid.cuh, host code path
extern "C" {
void shim_fp_mul_hybrid_many(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t count) {
    for (size_t i = 0; i < count; i++) {
        fp_t x, y, z;
        memcpy(x.l, a + 12 * i, 48); memcpy(y.l, b + 12 * i, 48);
        fp_mul_hybrid(z, x, y);
        memcpy(r + 12 * i, z.l, 48);
    }
}
void shim_fp_sqr_hybrid_many(const uint32_t *a, uint32_t *r, size_t count) {
    for (size_t i = 0; i < count; i++) {
        fp_t x, z;
        memcpy(x.l, a + 12 * i, 48);
        fp_sqr_hybrid(z, x);
        memcpy(r + 12 * i, z.l, 48);
    }
}
void shim_fp_product_many(const uint32_t *a, const uint32_t *b, uint32_t *t, size_t count, int sqr) {
    for (size_t i = 0; i < count; i++) {
        fp_t x, y;
        memcpy(x.l, a + 12 * i, 48); memcpy(y.l, b + 12 * i, 48);
        if (sqr) hy_product<true>(t + 24 * i, x, x); else hy_product<false>(t + 24 * i, x, y);
    }
}
void shim_fp_redc_many(const uint32_t *t, uint32_t *r, size_t count) {
    for (size_t i = 0; i < count; i++) { fp_t z; fe_redc(z, t + 24 * i); memcpy(r + 12 * i, z.l, 48); }
}
}
// lazy residues in [0, 2p) (bigint.cuh): used by the MSM levels
extern "C" {
void shim_fp_mul_lazy_many(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t count) {
    for (size_t i = 0; i < count; i++) { fp_t x, y, z; memcpy(x.l, a + 12 * i, 48); memcpy(y.l, b + 12 * i, 48); fe_mul_lazy(z, x, y); memcpy(r + 12 * i, z.l, 48); }
}
void shim_fp_sub_lazy_many(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t count) {
    for (size_t i = 0; i < count; i++) { fp_t x, y, z; memcpy(x.l, a + 12 * i, 48); memcpy(y.l, b + 12 * i, 48); fe_sub_lazy(z, x, y); memcpy(r + 12 * i, z.l, 48); }
}
int shim_fp_is_zero_lazy(const uint32_t *a) { fp_t x; memcpy(x.l, a, 48); return fe_is_zero_lazy(x); }
void shim_fp_canonical(const uint32_t *a, uint32_t *r) { fp_t x; memcpy(x.l, a, 48); fe_canonical(x); memcpy(r, x.l, 48); }
}
extern "C" {
void shim_fp_sub_lazy4_many(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t count) {
    for (size_t i = 0; i < count; i++) { fp_t x, y, z; memcpy(x.l, a + 12 * i, 48); memcpy(y.l, b + 12 * i, 48); fe_sub_lazy4(z, x, y); memcpy(r + 12 * i, z.l, 48); }
}
int shim_fp_is_zero_lazy4(const uint32_t *a) { fp_t x; memcpy(x.l, a, 48); return fe_is_zero_lazy4(x); }
}

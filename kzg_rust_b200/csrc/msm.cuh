// msm.cuh -- the fixed-base G1 multi-scalar multiplication behind
// `poly_to_kzg_commitment` / `g1_lincomb_fast` (reference src/kzg.rs:396-398,
// src/utils.rs:367-410), restructured for a B200:
//
//   * The 4096 Lagrange-basis points never change, so every multiple a signed c-bit digit
//     can ask for is precomputed once per context and kept in HBM:
//         table[(j*n + i)*D + (d-1)] = d * 2^(c*j) * G_i,   d = 1 .. D = 2^(c-1)
//     (c = 15: 17 windows, 102 GiB of the 180 GB).  A commitment is then just the sum of
//     W*n table entries -- no buckets, no bucket reduction, no doublings.
//   * That sum is a binary tree of *affine* additions.  Each level is one launch of
//     batch_add_kernel; a thread takes k independent additions, multiplies their
//     denominators together, inverts once (Montgomery's trick) and unwinds: 5 mul + 1 sqr
//     per addition plus an inversion amortised over k.
//   * All special cases of the group law are handled exactly (see g1.cuh), because the
//     reference's own vectors hit them: the all-zero blob sums 69,632 infinities.
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif

#include "g1.cuh"

namespace kzg {

// 128-bit loads/stores of field elements and points
KZG_HD void ld_fp(fp_t &r, const fp_t *p) {
#if !defined(__CUDA_ARCH__)
    r = *p;
#else
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
#endif
}
KZG_HD void st_fp(fp_t *p, const fp_t &r) {
#if !defined(__CUDA_ARCH__)
    *p = r;
#else
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.l[0], r.l[1], r.l[2], r.l[3]);
    q[1] = make_uint4(r.l[4], r.l[5], r.l[6], r.l[7]);
    q[2] = make_uint4(r.l[8], r.l[9], r.l[10], r.l[11]);
#endif
}

// ------------------------------------------------------------------ where the operands of addition #g live
// A policy maps the flat addition index g to two source points (nullptr = infinity, with
// an optional negation of y) and one destination.

// Level >= 1 of the tree: out[b][t] = in[b][2t] + in[b][2t+1]
struct TreePolicy {
    const g1_affine_t *in;
    g1_affine_t *out;
    uint32_t cnt_in, cnt_out;  // points per blob before / after this level
    KZG_HD const g1_affine_t *src(uint64_t g, int which, bool &neg) const {
        uint64_t b = g / cnt_out;
        uint32_t t = (uint32_t)(g - b * cnt_out);
        uint32_t e = 2 * t + which;
        neg = false;
        return e < cnt_in ? in + b * cnt_in + e : nullptr;
    }
    KZG_HD g1_affine_t *dst(uint64_t g) const { return out + g; }
};

// Level 0: operands are table entries selected by the signed digits of the scalars.
//   digits[(b*W + j)*n + i]  (int16, |d| <= D)
struct GatherPolicy {
    const g1_affine_t *table;
    const int16_t *digits;
    g1_affine_t *out;
    uint32_t per_blob;  // W*n
    uint32_t D;
    KZG_HD const g1_affine_t *src(uint64_t g, int which, bool &neg) const {
        uint32_t half = per_blob >> 1;
        uint64_t b = g / half;
        uint32_t e = 2 * (uint32_t)(g - b * half) + which;
        int d = digits[b * per_blob + e];
        neg = d < 0;
        if (d == 0) return nullptr;
        uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
        return table + ((uint64_t)e * D + (mag - 1));
    }
    KZG_HD g1_affine_t *dst(uint64_t g) const { return out + g; }
};

// Table construction, level L: for every slice s (= j*n + i) and d in (2^L, 2^(L+1)]:
//   table[s][d] = table[s][d >> 1] + table[s][(d + 1) >> 1]
struct TableLevelPolicy {
    g1_affine_t *table;
    uint32_t D;
    uint32_t level;
    KZG_HD uint64_t index(uint64_t g, uint32_t &d) const {
        uint64_t s = g >> level;
        d = (1u << level) + 1u + (uint32_t)(g & ((1ull << level) - 1));
        return s * D;
    }
    KZG_HD const g1_affine_t *src(uint64_t g, int which, bool &neg) const {
        uint32_t d;
        uint64_t base = index(g, d);
        neg = false;
        uint32_t e = which == 0 ? (d >> 1) : ((d + 1) >> 1);
        return table + base + (e - 1);
    }
    KZG_HD g1_affine_t *dst(uint64_t g) const {
        uint32_t d;
        uint64_t base = index(g, d);
        return table + base + (d - 1);
    }
};

template <class Policy>
KZG_HD void load_x(const Policy &pol, uint64_t g, int which, fp_t &x) {
    bool neg;
    const g1_affine_t *p = pol.src(g, which, neg);
    if (p == nullptr) {
#pragma unroll
        for (int i = 0; i < 12; i++) x.l[i] = 0xffffffffu;
    } else {
        ld_fp(x, &p->x);
    }
}
template <class Policy>
KZG_HD void load_y(const Policy &pol, uint64_t g, int which, fp_t &y) {
    bool neg;
    const g1_affine_t *p = pol.src(g, which, neg);
    if (p == nullptr) { fe_set_zero(y); return; }
    ld_fp(y, &p->y);
    if (neg) fe_neg(y, y);
}

// ------------------------------------------------------------------ the hot kernel
// total additions, spread as g = base + j*T + tid (j < k) so neighbouring threads touch
// neighbouring memory.  scratch holds T*k prefix products (48 B each).
#ifndef KZG_ADD_THREADS
#define KZG_ADD_THREADS 128
#endif
#ifndef KZG_ADD_MIN_BLOCKS
#define KZG_ADD_MIN_BLOCKS 3
#endif
// One thread's share; a plain function so a CPU test can walk it thread by thread.
template <class Policy>
KZG_HD void batch_add_thread(const Policy &pol, uint64_t total, fp_t *scratch, int k, uint64_t T, uint64_t tid) {
    for (uint64_t base = 0; base < total; base += T * (uint64_t)k) {
        // pass 1: running product of the denominators
        fp_t acc = fe_one<FpParams>();
        int cnt = 0;
#pragma unroll 1
        for (int j = 0; j < k; j++) {
            uint64_t g = base + (uint64_t)j * T + tid;
            if (g >= total) break;
            fp_t x1, x2, den;
            load_x(pol, g, 0, x1);
            load_x(pol, g, 1, x2);
            add_denominator(den, x1, x2, [&](fp_t &y) { load_y(pol, g, 0, y); }, [&](fp_t &y) { load_y(pol, g, 1, y); });
            st_fp(scratch + (uint64_t)j * T + tid, acc);
            fe_mul(acc, acc, den);
            cnt++;
        }
        if (cnt == 0) continue;
        fp_t inv;
        fp_inv(inv, acc);
        // pass 2: unwind
#pragma unroll 1
        for (int j = cnt - 1; j >= 0; j--) {
            uint64_t g = base + (uint64_t)j * T + tid;
            g1_affine_t p1, p2, r;
            load_x(pol, g, 0, p1.x);
            load_x(pol, g, 1, p2.x);
            load_y(pol, g, 0, p1.y);
            load_y(pol, g, 1, p2.y);
            fp_t den;
            int kind = add_denominator(den, p1.x, p2.x, [&](fp_t &y) { y = p1.y; }, [&](fp_t &y) { y = p2.y; });
            fp_t pre, inv_j;
            ld_fp(pre, scratch + (uint64_t)j * T + tid);
            fe_mul(inv_j, inv, pre);
            fe_mul(inv, inv, den);
            add_finish(r, kind, p1, p2, inv_j);
            g1_affine_t *o = pol.dst(g);
            st_fp(&o->x, r.x);
            st_fp(&o->y, r.y);
        }
    }
}
#if defined(__CUDACC__)
template <class Policy>
__global__ void __launch_bounds__(KZG_ADD_THREADS, KZG_ADD_MIN_BLOCKS)
batch_add_kernel(Policy pol, uint64_t total, fp_t *__restrict__ scratch, int k) {
    batch_add_thread(pol, total, scratch, k, (uint64_t)gridDim.x * blockDim.x,
                     (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
#endif

}  // namespace kzg

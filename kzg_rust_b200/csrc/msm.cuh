// msm.cuh -- the fixed-base G1 multi-scalar multiplication behind
// `poly_to_kzg_commitment` / `g1_lincomb_fast` (reference src/kzg.rs:396-398,
// src/utils.rs:367-410), restructured for a B200:
//
//   * The 4096 Lagrange-basis points never change, so every multiple a signed c-bit digit
//     can ask for is precomputed once per context and kept in HBM:
//         table[i*D + (d-1)] = d * G_i,   d = 1 .. D = 2^(c-1)
//     (c = 18: 51.5 GB, c = 19: 103 GB of the 180 GB).  With s_i = sum_j d_ij 2^(c j),
//         sum_i s_i G_i = sum_j 2^(c j) S_j,   S_j = sum_i d_ij G_i  (4096 table entries)
//     so a commitment is W = ceil(255/c) sums of 4096 table entries -- no buckets and no
//     bucket reduction -- followed by one short Horner pass (W-1 times c doublings + 1 add).
//   * Each S_j is a binary tree of *affine* additions.  Each level is one launch of
//     batch_add_kernel; a thread takes k independent additions, multiplies their
//     denominators together, the warp inverts once (Montgomery's trick across k x 32
//     additions) and every thread unwinds: 5 mul + 1 sqr per addition.
//   * All special cases of the group law are handled exactly (see g1.cuh), because the
//     reference's own vectors hit them: the all-zero blob sums 61,440 infinities.
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

// Division of a 32-bit index by a launch constant (the rows-per-point count R below) without a
// divide: q = floor(g * ceil(2^64 / d) / 2^64) is exact for g, d < 2^32.
struct FastDiv {
    uint64_t magic;
    uint32_t d;
    static FastDiv make(uint32_t d) {
        FastDiv f;
        f.d = d;
        f.magic = d > 1 ? ~0ull / d + 1 : 0;  // ceil(2^64 / d) (d does not divide 2^64 unless it is a power of two: +1 is still exact)
        return f;
    }
    KZG_HD uint32_t div(uint32_t g) const {
        if (d == 1) return g;
#if defined(__CUDA_ARCH__)
        return (uint32_t)__umul64hi((uint64_t)g, magic);
#else
        return (uint32_t)(((unsigned __int128)g * magic) >> 64);
#endif
    }
};

// The window sums run in a POINT-MAJOR layout: a level holds `rows` rows of R = (blobs in the
// chunk) x W points, row q of level L being the partial sums over points [q 2^L, (q+1) 2^L) of
// every (window, blob) pair:  level[q*R + (j*count + b)].  Adjacent threads work on adjacent
// (window, blob) pairs of the SAME point pair, so at any moment the whole GPU reads the table
// rows of a handful of points (2^(c-1) x 96 B = 25 MB each at c = 19) instead of scattering over
// the whole 100 GB table -- DRAM pages and TLB entries are reused -- and every load and store of
// the tree levels is coalesced.
//
// Level >= 1: out[q*R + r] = in[2q*R + r] + in[(2q+1)*R + r]
struct PairPolicy {
    const g1_affine_t *in;
    g1_affine_t *out;
    FastDiv R;
    KZG_HD const g1_affine_t *src(uint64_t g, int which, bool &neg) const {
        neg = false;
        uint32_t q = R.div((uint32_t)g);
        return in + g + (uint64_t)(q + (uint32_t)which) * R.d;
    }
    KZG_HD g1_affine_t *dst(uint64_t g) const { return out + g; }
};

// General form (odd group sizes; used by the small sums of batch verification):
// out[b][t] = in[b][2t] + in[b][2t+1]
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
//   digits[(i*W + j)*count + b]  (int32, |d| <= D), point-major like the levels above:
//   out[p*R + r] = sign * table[2p][|d(2p, r)|] + sign * table[2p+1][|d(2p+1, r)|],  r = j*count + b
struct GatherPolicy {
    const g1_affine_t *table;
    const int32_t *digits;
    g1_affine_t *out;
    FastDiv R;
    uint32_t D;
    KZG_HD const g1_affine_t *src(uint64_t g, int which, bool &neg) const {
        uint32_t p = R.div((uint32_t)g);
        uint32_t i = 2 * p + (uint32_t)which;
        int d = digits[g + (uint64_t)(p + (uint32_t)which) * R.d];  // = i*R + r
        neg = d < 0;
        if (d == 0) return nullptr;
        uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
        return table + ((uint64_t)i * D + (mag - 1));
    }
    KZG_HD g1_affine_t *dst(uint64_t g) const { return out + g; }
};

// Table construction, level L: for every point s (= i) and d in (2^L, 2^(L+1)]:
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
// One field inversion per thread block instead of one per thread.  An inversion is 580
// dependent multiplications; executed by every thread it costs a warp as many issue slots as
// ~100 additions.  Instead the running products of the block's threads are combined: shuffle
// scans give every lane the inclusive prefix and suffix products inside its warp (5 steps
// each), the warp totals go through shared memory, warp 0 inverts the block total while the
// other warps wait at the barrier (the other resident blocks keep the multiplier busy), and
// thread t recovers 1/acc_t = 1/total * (product of the other warps' totals) * prefix_{t-1}
// * suffix_{t+1}.  About 17 extra multiplications per thread buy an inversion shared by
// 128 k additions; this is what keeps the short launches at the bottom of the addition tree
// (few additions per thread) efficient.
#if defined(__CUDA_ARCH__)
KZG_D fp_t shfl_up_fp(const fp_t &v, int d) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], d);
    return r;
}
KZG_D fp_t shfl_down_fp(const fp_t &v, int d) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
    return r;
}
// all threads of the block must call this the same number of times (it has barriers)
KZG_D void shared_inverse(fp_t &inv, const fp_t &acc) {
    __shared__ fp_t sh_tot[32];
    __shared__ fp_t sh_inv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    fp_t pre = acc, suf = acc;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        fp_t t = shfl_up_fp(pre, d), u = shfl_down_fp(suf, d);
        if (lane >= d) fe_mul(pre, pre, t);
        if (lane + d < 32) fe_mul(suf, suf, u);
    }
    if (lane == 31) sh_tot[warp] = pre;
    __syncthreads();
    if (warp == 0) {
        fp_t total = sh_tot[0];
#pragma unroll 1
        for (int w = 1; w < nwarps; w++) fe_mul(total, total, sh_tot[w]);
        fp_t tinv;
        fp_inv(tinv, total);
        if (lane == 0) sh_inv = tinv;
    }
    __syncthreads();
    fp_t r = sh_inv;
#pragma unroll 1
    for (int w = 0; w < nwarps; w++)
        if (w != warp) fe_mul(r, r, sh_tot[w]);
    fp_t pe = shfl_up_fp(pre, 1), se = shfl_down_fp(suf, 1);
    if (lane > 0) fe_mul(r, r, pe);
    if (lane < 31) fe_mul(r, r, se);
    inv = r;
    __syncthreads();  // sh_tot / sh_inv are reused by the next batch
}
#else
KZG_HD void shared_inverse(fp_t &inv, const fp_t &acc) { fp_inv(inv, acc); }
#endif

#ifndef KZG_ADD_THREADS
#define KZG_ADD_THREADS 128
#endif
#ifndef KZG_ADD_MIN_BLOCKS
#define KZG_ADD_MIN_BLOCKS 3
#endif
// One thread's share; a plain function so a CPU test can walk it thread by thread.
// Both passes are software-pipelined: the operands of the next addition are requested before
// the multiplications of the current one, so the (random, for the gather level) HBM
// latency hides behind ~1300 pipe cycles of field multiplication even at 12 warps per SM.
template <class Policy>
KZG_HD void batch_add_thread(const Policy &pol, uint64_t total, fp_t *scratch, int k, uint64_t T, uint64_t tid) {
    for (uint64_t base = 0; base < total; base += T * (uint64_t)k) {
        // pass 1: running product of the denominators
        fp_t acc = fe_one<FpParams>();
        int cnt = 0;
        fp_t nx1, nx2;
        {
            uint64_t g = base + tid;
            if (g < total) { load_x(pol, g, 0, nx1); load_x(pol, g, 1, nx2); }
        }
#pragma unroll 1
        for (int j = 0; j < k; j++) {
            uint64_t g = base + (uint64_t)j * T + tid;
            if (g >= total) break;
            fp_t x1 = nx1, x2 = nx2, den;
            uint64_t gn = g + T;
            if (j + 1 < k && gn < total) { load_x(pol, gn, 0, nx1); load_x(pol, gn, 1, nx2); }
            add_denominator(den, x1, x2, [&](fp_t &y) { load_y(pol, g, 0, y); }, [&](fp_t &y) { load_y(pol, g, 1, y); });
            st_fp(scratch + (uint64_t)j * T + tid, acc);
            fe_mul(acc, acc, den);
            cnt++;
        }
        // every lane of the warp takes part (idle lanes contribute acc = 1)
        fp_t inv;
        shared_inverse(inv, acc);
        // pass 2: unwind
        fp_t npre;
        if (cnt > 0) {
            uint64_t g = base + (uint64_t)(cnt - 1) * T + tid;
            load_x(pol, g, 0, nx1);
            load_x(pol, g, 1, nx2);
            ld_fp(npre, scratch + (uint64_t)(cnt - 1) * T + tid);
        }
#pragma unroll 1
        for (int j = cnt - 1; j >= 0; j--) {
            uint64_t g = base + (uint64_t)j * T + tid;
            g1_affine_t p1, p2, r;
            p1.x = nx1;
            p2.x = nx2;
            fp_t pre = npre;
            load_y(pol, g, 0, p1.y);
            load_y(pol, g, 1, p2.y);
            if (j > 0) {
                load_x(pol, g - T, 0, nx1);
                load_x(pol, g - T, 1, nx2);
                ld_fp(npre, scratch + (uint64_t)(j - 1) * T + tid);
            }
            fp_t den;
            int kind = add_denominator(den, p1.x, p2.x, [&](fp_t &y) { y = p1.y; }, [&](fp_t &y) { y = p2.y; });
            fp_t inv_j;
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
// MINB = resident blocks per SM the register allocation is tuned for (3: 168 registers, no
// spills; 4: 128 registers, a few spilled words)
template <class Policy, int MINB>
__global__ void __launch_bounds__(KZG_ADD_THREADS, MINB)
batch_add_kernel(Policy pol, uint64_t total, fp_t *__restrict__ scratch, int k) {
    batch_add_thread(pol, total, scratch, k, (uint64_t)gridDim.x * blockDim.x,
                     (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
#endif

}  // namespace kzg

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
    static constexpr bool lazy = true;  // coordinates in [0, 2p) between the levels of the MSM
    const g1_affine_t *in;
    g1_affine_t *out;
    FastDiv R;
    KZG_HD int digit(uint64_t, int) const { return 0; }
    KZG_HD const g1_affine_t *src_d(uint64_t g, int which, int, bool &neg) const { return src(g, which, neg); }
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
    static constexpr bool lazy = false;
    const g1_affine_t *in;
    g1_affine_t *out;
    uint32_t cnt_in, cnt_out;  // points per blob before / after this level
    KZG_HD int digit(uint64_t, int) const { return 0; }
    KZG_HD const g1_affine_t *src_d(uint64_t g, int which, int, bool &neg) const { return src(g, which, neg); }
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
    static constexpr bool lazy = true;  // canonical table entries in, [0, 2p) out
    const g1_affine_t *table;
    const int32_t *digits;
    g1_affine_t *out;
    FastDiv R;
    uint32_t D;
    // The digit is read one iteration before the table entry it selects (digit -> address -> entry is a
    // chain of two dependent loads; ncu showed 29 % of the gather level's stall samples on the digit test).
    KZG_HD int digit(uint64_t g, int which) const {
        uint32_t p = R.div((uint32_t)g);
        return digits[g + (uint64_t)(p + (uint32_t)which) * R.d];  // = i*R + r
    }
    KZG_HD const g1_affine_t *src_d(uint64_t g, int which, int d, bool &neg) const {
        neg = d < 0;
        if (d == 0) return nullptr;
        uint32_t i = 2 * R.div((uint32_t)g) + (uint32_t)which;
        uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
        return table + ((uint64_t)i * D + (mag - 1));
    }
    KZG_HD const g1_affine_t *src(uint64_t g, int which, bool &neg) const { return src_d(g, which, digit(g, which), neg); }
    KZG_HD g1_affine_t *dst(uint64_t g) const { return out + g; }
};

// Table construction, level L: for every point s (= i) and d in (2^L, 2^(L+1)]:
//   table[s][d] = table[s][d >> 1] + table[s][(d + 1) >> 1]
struct TableLevelPolicy {
    static constexpr bool lazy = false;  // the table is canonical
    g1_affine_t *table;
    uint32_t D;
    uint32_t level;
    KZG_HD int digit(uint64_t, int) const { return 0; }
    KZG_HD const g1_affine_t *src_d(uint64_t g, int which, int, bool &neg) const { return src(g, which, neg); }
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

template <class Policy>
KZG_HD void load_x_d(const Policy &pol, uint64_t g, int which, int d, fp_t &x) {
    bool neg;
    const g1_affine_t *p = pol.src_d(g, which, d, neg);
    if (p == nullptr) {
#pragma unroll
        for (int i = 0; i < 12; i++) x.l[i] = 0xffffffffu;
    } else {
        ld_fp(x, &p->x);
    }
}
template <class Policy>
KZG_HD void load_y_d(const Policy &pol, uint64_t g, int which, int d, fp_t &y) {
    bool neg;
    const g1_affine_t *p = pol.src_d(g, which, d, neg);
    if (p == nullptr) { fe_set_zero(y); return; }
    ld_fp(y, &p->y);
    if (neg) fe_neg(y, y);
}

// y as stored, with the sign still to be applied: the negation is the first consumer of the load, so it is
// deferred until two products later (ncu: 11 % of the gather level's stall samples sat on it)
template <class Policy>
KZG_HD void load_y_raw_d(const Policy &pol, uint64_t g, int which, int d, fp_t &y, bool &neg) {
    const g1_affine_t *p = pol.src_d(g, which, d, neg);
    if (p == nullptr) { fe_set_zero(y); neg = false; return; }
    ld_fp(y, &p->y);
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
        int dn0 = 0, dn1 = 0;  // digits of the addition whose operands are requested next (gather level)
        {
            uint64_t g = base + tid;
            if (g < total) { load_x(pol, g, 0, nx1); load_x(pol, g, 1, nx2); }
            g += T;
            if (1 < k && g < total) { dn0 = pol.digit(g, 0); dn1 = pol.digit(g, 1); }
        }
#pragma unroll 1
        for (int j = 0; j < k; j++) {
            uint64_t g = base + (uint64_t)j * T + tid;
            if (g >= total) break;
            fp_t x1 = nx1, x2 = nx2, den;
            uint64_t gn = g + T;
            if (j + 1 < k && gn < total) { load_x_d(pol, gn, 0, dn0, nx1); load_x_d(pol, gn, 1, dn1, nx2); }
            gn += T;
            if (j + 2 < k && gn < total) { dn0 = pol.digit(gn, 0); dn1 = pol.digit(gn, 1); }
            add_denominator<Policy::lazy>(den, x1, x2, [&](fp_t &y) { load_y(pol, g, 0, y); }, [&](fp_t &y) { load_y(pol, g, 1, y); });
            st_fp(scratch + (uint64_t)j * T + tid, acc);
            fpx_mul<Policy::lazy>(acc, acc, den);
            cnt++;
        }
        // every lane of the warp takes part (idle lanes contribute acc = 1)
        fp_t inv;
        shared_inverse(inv, acc);
        // pass 2: unwind
        fp_t npre;
        int dc0 = 0, dc1 = 0;  // digits of the current addition, of the one before it
        dn0 = dn1 = 0;
        if (cnt > 0) {
            uint64_t g = base + (uint64_t)(cnt - 1) * T + tid;
            dc0 = pol.digit(g, 0);
            dc1 = pol.digit(g, 1);
            load_x_d(pol, g, 0, dc0, nx1);
            load_x_d(pol, g, 1, dc1, nx2);
            ld_fp(npre, scratch + (uint64_t)(cnt - 1) * T + tid);
            if (cnt > 1) { dn0 = pol.digit(g - T, 0); dn1 = pol.digit(g - T, 1); }
        }
#pragma unroll 1
        for (int j = cnt - 1; j >= 0; j--) {
            uint64_t g = base + (uint64_t)j * T + tid;
            g1_affine_t p1, p2, r;
            p1.x = nx1;
            p2.x = nx2;
            fp_t pre = npre;
            bool ng1, ng2;
            load_y_raw_d(pol, g, 0, dc0, p1.y, ng1);
            load_y_raw_d(pol, g, 1, dc1, p2.y, ng2);
            if (j > 0) {
                load_x_d(pol, g - T, 0, dn0, nx1);
                load_x_d(pol, g - T, 1, dn1, nx2);
                ld_fp(npre, scratch + (uint64_t)(j - 1) * T + tid);
                dc0 = dn0;
                dc1 = dn1;
                if (j > 1) { dn0 = pol.digit(g - 2 * T, 0); dn1 = pol.digit(g - 2 * T, 1); }
            }
            fp_t den;
            int kind = add_denominator<Policy::lazy>(
                den, p1.x, p2.x, [&](fp_t &y) { y = p1.y; if (ng1) fe_neg(y, y); }, [&](fp_t &y) { y = p2.y; if (ng2) fe_neg(y, y); });
            fp_t inv_j;
            fpx_mul<Policy::lazy>(inv_j, inv, pre);
            fpx_mul<Policy::lazy>(inv, inv, den);
            // y of a table entry is canonical and never 0 (odd group order): 2p - y is a lazy residue
            if (ng1) { if (Policy::lazy) fe_neg_lazy(p1.y, p1.y); else fe_neg(p1.y, p1.y); }
            if (ng2) { if (Policy::lazy) fe_neg_lazy(p2.y, p2.y); else fe_neg(p2.y, p2.y); }
            add_finish<Policy::lazy>(r, kind, p1, p2, inv_j);
            g1_affine_t *o = pol.dst(g);
            st_fp(&o->x, r.x);
            st_fp(&o->y, r.y);
        }
    }
}
#if defined(__CUDACC__)
#ifdef KZG_TRACE
// experiment only (tools/trace_blocks.py): per-block start / end time and SM id of the last launch
__device__ unsigned long long *g_kzg_trace = nullptr;
KZG_D unsigned long long kzg_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
KZG_D unsigned kzg_smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
#endif
// ------------------------------------------------------------------ the hot kernel, work-pulling form
// batch_add_kernel gives every thread the same number of additions.  On the B200 that wastes a
// fifth of the multiplier: the warp scheduler of an SM sub-partition serves its resident warps by
// PRIORITY, not round-robin, so of the three blocks of an SM two run at full speed and finish at
// 0.70 of the launch (gather level: 0.85) while the third starves and then runs alone -- one warp per
// scheduler cannot keep the multiplier busy (measured with %globaltimer per warp, tools/trace_blocks.py).
// Here the additions are cut into tiles of 32 (one per lane) and every WARP pulls tiles from a global
// counter: whichever warps the scheduler favours simply take more tiles, and all warps finish together.
//   * A warp works in batches: it pulls up to m tiles (pass 1: running product of the denominators,
//     prefix products to its private scratch), inverts once for the whole warp (shuffle scans + one
//     inversion, no block barrier anywhere), and unwinds the same tiles in reverse (pass 2).
//   * m shrinks with the work that is left (guided self-scheduling), so the launch starts with long
//     batches (few inversions) and ends with short ones (small imbalance at the end).
//   * Tiles are handed out in index order, so at any moment all warps of the GPU work inside a window of
//     a few thousand tiles: the locality of the point-major layout survives the dynamic order.
KZG_D void warp_inverse(fp_t &inv, const fp_t &acc) {
#if defined(__CUDA_ARCH__)
    const int lane = threadIdx.x & 31;
    fp_t pre = acc, suf = acc;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        fp_t t = shfl_up_fp(pre, d), u = shfl_down_fp(suf, d);
        if (lane >= d) fe_mul(pre, pre, t);
        if (lane + d < 32) fe_mul(suf, suf, u);
    }
    fp_t total;
#pragma unroll
    for (int i = 0; i < 12; i++) total.l[i] = __shfl_sync(0xffffffffu, pre.l[i], 31);
    fp_t r;
    fp_inv(r, total);  // the same value in every lane: no divergence
    fp_t pe = shfl_up_fp(pre, 1), se = shfl_down_fp(suf, 1);
    if (lane > 0) fe_mul(r, r, pe);
    if (lane < 31) fe_mul(r, r, se);
    inv = r;
#endif
}

struct DynSchedule {
    unsigned int *counter;  // next tile to hand out (zero before the launch)
    uint32_t *tile_ids;     // [warps][cap]: the tiles of the current batch of each warp
    fp_t *scratch;          // [warps][cap][32] prefix products
    uint32_t ntiles;        // ceil(total / 32)
    uint32_t cap;           // longest batch
    uint32_t m_min;         // shortest batch (1 for small launches)
};

template <class Policy, int MINB>
__global__ void __launch_bounds__(KZG_ADD_THREADS, MINB)
batch_add_dyn_kernel(Policy pol, uint64_t total, DynSchedule ds) {
#if defined(__CUDA_ARCH__)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    fp_t *scratch = ds.scratch + (size_t)warp * ds.cap * 32 + lane;
    uint32_t *ids = ds.tile_ids + (size_t)warp * ds.cap;
    const uint32_t NONE = 0xffffffffu;
#ifdef KZG_TRACE
    unsigned long long t0 = kzg_globaltimer();
#endif
    auto pull = [&]() -> uint32_t {  // lane 0 takes the next tile; every lane gets its index (NONE when exhausted)
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(ds.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        return t < ds.ntiles ? t : NONE;
    };
    for (;;) {
        // batch length from the work that is left
        uint32_t seen = 0;
        if (lane == 0) seen = *(volatile unsigned int *)ds.counter;
        seen = __shfl_sync(0xffffffffu, seen, 0);
        if (seen >= ds.ntiles) break;
        uint32_t m = (ds.ntiles - seen) / (2 * nwarps);
        m = m < ds.m_min ? ds.m_min : (m > ds.cap ? ds.cap : m);
        // pass 1
        fp_t acc = fe_one<FpParams>();
        fp_t nx1, nx2;
        uint32_t cnt = 0;
        uint32_t tile = pull(), tile_n = NONE;
        if (tile == NONE) break;
        {
            uint64_t g = (uint64_t)tile * 32 + lane;
            if (g < total) { load_x(pol, g, 0, nx1); load_x(pol, g, 1, nx2); }
        }
        if (m > 1) tile_n = pull();
#pragma unroll 1
        for (uint32_t j = 0; j < m && tile != NONE; j++) {
            const uint64_t g = (uint64_t)tile * 32 + lane;
            fp_t x1 = nx1, x2 = nx2;
            if (lane == 0) ids[j] = tile;
            // operands of the next tile, and the index of the one after it, are requested before the multiplication
            uint32_t tile_nn = NONE;
            if (tile_n != NONE) {
                uint64_t gn = (uint64_t)tile_n * 32 + lane;
                if (gn < total) { load_x(pol, gn, 0, nx1); load_x(pol, gn, 1, nx2); }
                if (j + 2 < m) tile_nn = pull();
            }
            if (g < total) {
                fp_t den;
                add_denominator<Policy::lazy>(den, x1, x2, [&](fp_t &y) { load_y(pol, g, 0, y); }, [&](fp_t &y) { load_y(pol, g, 1, y); });
                st_fp(scratch + (size_t)j * 32, acc);
                fpx_mul<Policy::lazy>(acc, acc, den);
            }
            cnt++;
            tile = tile_n;
            tile_n = tile_nn;
        }
        __syncwarp();
        fp_t inv;
        warp_inverse(inv, acc);
        // pass 2: unwind the same tiles in reverse
        fp_t npre;
        uint32_t t_cur = ids[cnt - 1];
        {
            uint64_t g = (uint64_t)t_cur * 32 + lane;
            if (g < total) { load_x(pol, g, 0, nx1); load_x(pol, g, 1, nx2); ld_fp(npre, scratch + (size_t)(cnt - 1) * 32); }
        }
#pragma unroll 1
        for (int j = (int)cnt - 1; j >= 0; j--) {
            const uint64_t g = (uint64_t)t_cur * 32 + lane;
            g1_affine_t p1, p2, r;
            p1.x = nx1;
            p2.x = nx2;
            fp_t pre = npre;
            const bool live = g < total;
            if (live) { load_y(pol, g, 0, p1.y); load_y(pol, g, 1, p2.y); }
            if (j > 0) {
                t_cur = ids[j - 1];
                uint64_t gp = (uint64_t)t_cur * 32 + lane;
                if (gp < total) { load_x(pol, gp, 0, nx1); load_x(pol, gp, 1, nx2); ld_fp(npre, scratch + (size_t)(j - 1) * 32); }
            }
            if (live) {
                fp_t den;
                int kind = add_denominator<Policy::lazy>(den, p1.x, p2.x, [&](fp_t &y) { y = p1.y; }, [&](fp_t &y) { y = p2.y; });
                fp_t inv_j;
                fpx_mul<Policy::lazy>(inv_j, inv, pre);
                fpx_mul<Policy::lazy>(inv, inv, den);
                add_finish<Policy::lazy>(r, kind, p1, p2, inv_j);
                g1_affine_t *o = pol.dst(g);
                st_fp(&o->x, r.x);
                st_fp(&o->y, r.y);
            }
        }
        __syncwarp();
    }
#ifdef KZG_TRACE
    if (g_kzg_trace && g_kzg_trace[0] == total && lane == 0) {
        unsigned long long *r = g_kzg_trace + 4ull * (1 + warp);
        r[0] = t0; r[1] = kzg_globaltimer(); r[2] = kzg_smid(); r[3] = total;
    }
#endif
#endif
}

// MINB = resident blocks per SM the register allocation is tuned for (3: 168 registers, no
// spills; 4: 128 registers, a few spilled words)
template <class Policy, int MINB>
__global__ void __launch_bounds__(KZG_ADD_THREADS, MINB)
batch_add_kernel(Policy pol, uint64_t total, fp_t *__restrict__ scratch, int k) {
#ifdef KZG_TRACE
    unsigned long long t0 = kzg_globaltimer();
#endif
    batch_add_thread(pol, total, scratch, k, (uint64_t)gridDim.x * blockDim.x,
                     (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
#ifdef KZG_TRACE
    if (g_kzg_trace && g_kzg_trace[0] == total && (threadIdx.x & 31) == 0) {  // slot 0 = the launch to record
        unsigned long long *r = g_kzg_trace + 4ull * (1 + blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32);
        r[0] = t0; r[1] = kzg_globaltimer(); r[2] = kzg_smid(); r[3] = total;
    }
#endif
}
#endif

}  // namespace kzg

// msm.cuh -- the fixed-base G1 multi-scalar multiplication behind
// `poly_to_kzg_commitment` / `g1_lincomb_fast` (reference src/kzg.rs:396-398,
// src/utils.rs:367-410), restructured for a B200 as a signed comb over GROUPS of setup points:
//
//   * The 4096 Lagrange-basis points never change.  They are cut into groups of g consecutive
//     points (g = the comb width, 23 by default) and every signed sum of a group that keeps its
//     first point positive is precomputed once per context and kept in HBM:
//         T[q][idx] = P_{q,0} + sum_{k=1..g-1} (bit_{k-1}(idx) ? + : -) P_{q,k}     2^(g-1) entries per group
//     (g = 23: 179 groups x 4 Mi entries x 96 B = 72 GB; g = 22: 38 GB; g = 20: 10 GB).
//   * Every scalar is written with digits +-1 only, s = sum_j e_j 2^j (blobpath.cuh), so bit position j of a
//     blob needs ONE table entry per group -- the one selected by the g signs e_{q g + k, j}, negated when the
//     first of them is minus:
//         sum_i s_i G_i = sum_j 2^j S_j,      S_j = sum_q +-T[q][idx(q, j)]          (ceil(4096/g) entries)
//     A commitment is 255 sums of ceil(4096/g) table entries plus one Horner pass (254 doublings and
//     additions): 255 x 178 = 45,390 additions per blob at g = 23, where per-point windows of the same table
//     size need 14 x 4095 = 57,330 and the bucket method 90,112.  No buckets, no doublings in the hot loop.
//   * Each S_j is a binary tree of *affine* additions.  Each level is one launch of
//     batch_add_kernel; a thread takes k independent additions, multiplies their
//     denominators together, the thread block inverts once (Montgomery's trick across k x 128
//     additions) and every thread unwinds: 5 mul + 1 sqr per addition.
//   * All special cases of the group law are handled exactly (see g1.cuh), because the
//     reference's own vectors hit them: in the all-zero blob every scalar is recoded as r, and the
//     constant blob selects the same table entry in every group pattern.
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
// A policy maps the flat addition index g to two source points (a stored point; infinity carries its marker, with
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

// The sums run in a GROUP-MAJOR layout: a level holds `rows` rows of R = (blobs in the chunk) x 255
// points, row q of level L being the partial sums over groups [q 2^L, (q+1) 2^L) of every
// (bit position, blob) pair:  level[q*R + (j*count + b)].  Adjacent threads work on adjacent
// (bit position, blob) pairs of the SAME pair of groups, so at any moment the whole GPU reads the tables
// of a handful of groups (2^(g-1) x 96 B = 403 MB each at g = 23) instead of scattering over the
// whole table -- DRAM pages and TLB entries are reused -- and every load and store of
// the tree levels is coalesced.
//
// Level >= 1: out[q*R + r] = in[2q*R + r] + in[(2q+1)*R + r]   (an odd last row is copied by the host code)
struct PairPolicy {
    static constexpr bool lazy = true;  // coordinates in [0, 2p) between the levels of the MSM
    const g1_affine_t *in;
    g1_affine_t *out;
    FastDiv R;
    KZG_HD uint32_t digit(uint64_t, int) const { return 0; }
    KZG_HD const g1_affine_t *src_d(uint64_t g, int which, uint32_t, bool &neg) const { return src(g, which, neg); }
    KZG_HD const g1_affine_t *src(uint64_t g, int which, bool &neg) const {
        neg = false;
        uint32_t q = R.div((uint32_t)g);
        return in + g + (uint64_t)(q + (uint32_t)which) * R.d;
    }
    KZG_HD g1_affine_t *dst(uint64_t g) const { return out + g; }
};

// Level 0: operands are table entries selected by the comb digits (blobpath.cuh: table index, bit 31 = negate).
//   digits[(q*255 + j)*count + b], group-major like the levels above:
//   out[p*R + r] = +-T[2p][idx(2p, r)] + +-T[2p+1][idx(2p+1, r)],  r = j*count + b
struct GatherPolicy {
    static constexpr bool lazy = true;  // canonical table entries in, [0, 2p) out
    const g1_affine_t *table;
    const uint32_t *digits;
    g1_affine_t *out;
    FastDiv R;
    uint32_t E;  // table entries per group, 2^(g-1)
    // The digit is read one iteration before the table entry it selects (digit -> address -> entry is a
    // chain of two dependent loads).  (Pulling the entry into L2 with prefetch.global.L2 one more iteration ahead
    // was measured and made the level 25 % SLOWER -- profiles/sweep_r2c.log -- and requesting the operands two
    // additions ahead in registers made it 8 % slower -- profiles/sweep_r2f.log: the level is limited by the rate
    // of random 64-byte DRAM accesses, not by their latency, so neither is done.)
    KZG_HD uint32_t digit(uint64_t g, int which) const {
        uint32_t p = R.div((uint32_t)g);
        return digits[g + (uint64_t)(p + (uint32_t)which) * R.d];  // = (2p + which)*R + r
    }
    KZG_HD const g1_affine_t *src_d(uint64_t g, int which, uint32_t d, bool &neg) const {
        neg = (d >> 31) != 0;
        uint32_t q = 2 * R.div((uint32_t)g) + (uint32_t)which;
        return table + ((uint64_t)q * E + (d & 0x7fffffffu));
    }
    KZG_HD const g1_affine_t *src(uint64_t g, int which, bool &neg) const { return src_d(g, which, digit(g, which), neg); }
    KZG_HD g1_affine_t *dst(uint64_t g) const { return out + g; }
};

// Table construction.  With P_m[q][u] = P_{q,0} + sum_{k=1..m-1} (bit_{k-1}(u) ? + : -) P_{q,k} (2^(m-1) entries,
// kept in the first slots of the group's table), step m -> m + 1 is
//   upper = 1:  T[q][u + 2^(m-1)] = T[q][u] + P_{q,m}        (new slots)
//   upper = 0:  T[q][u]           = T[q][u] - P_{q,m}        (in place, after the upper half has read the slot)
// bases[q*g + k] = P_{q,k}; the last group is padded with points at infinity.
struct CombLevelPolicy {
    static constexpr bool lazy = false;  // the table is canonical
    g1_affine_t *table;
    const g1_affine_t *bases;
    uint32_t g, m, upper;
    KZG_HD uint32_t digit(uint64_t, int) const { return 0; }
    KZG_HD const g1_affine_t *src_d(uint64_t a, int which, uint32_t, bool &neg) const { return src(a, which, neg); }
    KZG_HD uint64_t slot(uint64_t a, uint32_t &q) const {
        q = (uint32_t)(a >> (m - 1));
        return ((uint64_t)q << (g - 1)) + (a & ((1ull << (m - 1)) - 1));
    }
    KZG_HD const g1_affine_t *src(uint64_t a, int which, bool &neg) const {
        uint32_t q;
        uint64_t s = slot(a, q);
        neg = which == 1 && !upper;
        return which == 0 ? table + s : bases + ((uint64_t)q * g + m);
    }
    KZG_HD g1_affine_t *dst(uint64_t a) const {
        uint32_t q;
        uint64_t s = slot(a, q);
        return table + s + (upper ? (1ull << (m - 1)) : 0ull);
    }
};

// Operand loads.  No policy has a "missing operand": a point at infinity is a stored point with the marker in x
// (g1.cuh), so the loads are unconditional and the special cases are found by looking at the values.
template <class Policy>
KZG_HD void load_x(const Policy &pol, uint64_t g, int which, fp_t &x) {
    bool neg;
    ld_fp(x, &pol.src(g, which, neg)->x);
}
template <class Policy>
KZG_HD void load_y(const Policy &pol, uint64_t g, int which, fp_t &y) {
    bool neg;
    const g1_affine_t *p = pol.src(g, which, neg);
    ld_fp(y, &p->y);
    if (neg) fe_neg(y, y);
}
template <class Policy>
KZG_HD void load_x_d(const Policy &pol, uint64_t g, int which, uint32_t d, fp_t &x) {
    bool neg;
    ld_fp(x, &pol.src_d(g, which, d, neg)->x);
}
// y as stored, with the sign still to be applied: the negation is the first consumer of the load, so it is
// deferred until two products later (ncu: 11 % of the gather level's stall samples sat on it)
template <class Policy>
KZG_HD void load_y_raw_d(const Policy &pol, uint64_t g, int which, uint32_t d, fp_t &y, bool &neg) {
    ld_fp(y, &pol.src_d(g, which, d, neg)->y);
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
#ifndef KZG_ADD_UNROLL
#define KZG_ADD_UNROLL 1
#endif
#define KZG_PRAGMA_(x) _Pragma(#x)
#define KZG_UNROLL(n) KZG_PRAGMA_(unroll n)
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
        uint32_t dn0 = 0, dn1 = 0;  // digits of the addition whose operands are requested next (gather level)
        {
            uint64_t g = base + tid;
            if (g < total) { load_x(pol, g, 0, nx1); load_x(pol, g, 1, nx2); }
            g += T;
            if (1 < k && g < total) { dn0 = pol.digit(g, 0); dn1 = pol.digit(g, 1); }
        }
KZG_UNROLL(KZG_ADD_UNROLL)
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
        uint32_t dc0 = 0, dc1 = 0;  // digits of the current addition, of the one before it
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
KZG_UNROLL(KZG_ADD_UNROLL)
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

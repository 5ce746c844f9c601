// pippenger.cuh -- the variable-base multi-scalar multiplications of batch verification, phase B:
//
//     A = sum_i [r_i] proof_i          B = sum_i [r_i] C_i + sum_i [r_i z_i] proof_i        (r_i = r^(first + i))
//
// (reference verify_kzg_proof_batch, src/kzg.rs:596-622: proof_lincomb, C_minus_y_lincomb without its generator
// terms and proof_z_lincomb, three g1_lincomb_naive calls = 3n full scalar multiplications, src/utils.rs:329-342).
// Here they are ONE bucket-method pass (Pippenger) over three "lanes" of n points each:
//
//   * every scalar is split by the endomorphism, k = k1 + k2 lambda, psi(P) = [lambda]P = (beta^2 x, y), into two halves
//     below 2^128, so a lane is 2n (point, 128-bit scalar) pairs and the result needs 129 bit positions;
//   * the halves are cut into W = ceil(129 / c) signed windows of c bits (digits in [-2^(c-1), 2^(c-1)], Booth
//     recoding: stateless per window); digit d of window w sends +-point into bucket |d| of (lane, w).
//     c grows with n so that a bucket receives about 16 points: the accumulation is a short chain per thread
//     at every batch size (n = 6: c = 1, one bucket per bit; n = 16,384: c = 12, 3 x 11 x 2048 buckets);
//   * buckets are filled by a counting sort (k_pip_digits<false> counts, k_pip_scan turns the counts into
//     offsets, k_pip_digits<true> writes the entries) and summed one thread per bucket with mixed additions;
//   * the classical running-sum reduction sum_b b B_b is a serial chain of 2^c additions; instead every
//     bit position p = c w + j gets the plain sum of the buckets whose number has bit j set,
//         S_p = sum_{b : bit j of b} B_(w, b),           result = sum_p 2^p S_p,
//     one thread block per (output, p): a strided partial sum per thread, a shared-memory tree, and the
//     p doublings on thread 0; the 2 x 129 weighted rows are added by k_jac_sum (frpath.cuh);
//   * the last window holds only 129 - c (W - 1) bits, so its points crowd into few buckets (one, when c divides
//     128: the carry bit): 2^(c - bits) threads share each of those lists, and the row sums add their parts.
//
// Work for n = 16,384: 3 x 2 x 11 n = 1.1 M mixed additions + 0.13 M Jacobian ones, where the 3n ladders of
// k_verify_terms take 6.3 M doublings and 4.7 M additions; the dependent chain is ~16 additions + 7 tree levels
// + 129 doublings whatever n is, where a ladder is 128 doublings AND 96 additions deep.
//
// The thread functions are plain KZG_HD code: tests/hostshim walks them on the CPU against the ladders.
#pragma once
#include "blobpath.cuh"

namespace kzg {

#define KZG_PIP_LANES 3          // 0: (proof_i, r_i) -> A;  1: (C_i, r_i) -> B;  2: (proof_i, r_i z_i) -> B
#define KZG_PIP_SCALAR_BITS 129  // halves are below 2^128; one more bit so that the top window's carry is zero
#define KZG_PIP_MAX_C 12

struct PipPlan {
    int c;            // window width
    int W;            // windows: ceil(129 / c)
    uint32_t nbw;     // buckets per (lane, window): 2^(c-1), bucket b holds the points with |digit| = b + 1
    uint32_t buckets; // 3 W nbw
    int top_bits;     // bits of the last window: 129 - c (W - 1), in [1, c].  Its digits reach only 2^(top_bits-1), so the
                      // lists of its few buckets are 2^(c-top_bits) times longer: each is shared by that many threads
    int top_shift;    // c - top_bits
    uint32_t rows;    // bit positions: c (W - 1) + top_bits = 129
    uint64_t max_entries;  // 3 lanes x 2 halves x W windows x n
    static PipPlan make(size_t n, int force_c = 0) {
        PipPlan p;
        int c = 1;
        // 2n pairs per lane and window over 2^(c-1) buckets: about 12 .. 24 per bucket
        while (c < KZG_PIP_MAX_C && ((size_t)12 << c) <= 2 * n) c++;
        if (force_c > 0) c = force_c < KZG_PIP_MAX_C ? force_c : KZG_PIP_MAX_C;
        p.c = c;
        p.W = (KZG_PIP_SCALAR_BITS + c - 1) / c;
        p.nbw = 1u << (c - 1);
        p.buckets = KZG_PIP_LANES * (uint32_t)p.W * p.nbw;
        p.top_bits = KZG_PIP_SCALAR_BITS - c * (p.W - 1);
        p.top_shift = c - p.top_bits;
        p.rows = KZG_PIP_SCALAR_BITS;
        p.max_entries = (uint64_t)KZG_PIP_LANES * 2 * (uint64_t)p.W * n;
        return p;
    }
};

// entry of a bucket list: blob index | bit 29: psi | bit 30: negate | bit 31: the point is C_i (else proof_i)
#define KZG_PIP_PSI (1u << 29)
#define KZG_PIP_NEG (1u << 30)
#define KZG_PIP_CPT (1u << 31)
#define KZG_PIP_IDX 0x1fffffffu

// r^e for a Montgomery r
KZG_HD void fr_pow_u64(fr_t &out, const fr_t &r, uint64_t e) {
    fr_t acc = fe_one<FrParams>(), base = r;
#pragma unroll 1
    while (e) {
        if (e & 1) fe_mul(acc, acc, base);
        fe_sqr(base, base);
        e >>= 1;
    }
    out = acc;
}

// Blob i: r_i, s_i = r_i y_i -> sy[i] (Montgomery), and the four scalar halves
//   halves[4i + 0 .. 1] = k1, k2 of r_i          halves[4i + 2 .. 3] = k1, k2 of r_i z_i        (4 limbs each)
KZG_HD void pip_scalars_thread(uint32_t i, const uint8_t *zy, const fr_t &r_canon, uint64_t first, uint32_t *halves, fr_t *sy) {
    fr_t r, ri, y, z, s, k;
    fe_to_mont(r, r_canon);
    fr_pow_u64(ri, r, first + i);
    scalar_from_be32(y, zy + 64ull * i + 32);
    fe_to_mont(y, y);
    fe_mul(s, ri, y);
    sy[i] = s;
    scalar_from_be32(z, zy + 64ull * i);
    fe_to_mont(z, z);
    fe_mul(z, ri, z);
    fe_from_mont(k, ri);
    glv_split(halves + 16ull * i, halves + 16ull * i + 4, k.l);
    fe_from_mont(k, z);
    glv_split(halves + 16ull * i + 8, halves + 16ull * i + 12, k.l);
}

// Signed digit of window w of a 128-bit half h: with u = bits [c w, c w + c) and t = bit c w - 1,
// d = u + t - 2^c [bit c-1 of u]; sum_w d_w 2^(c w) = h because c W >= 129.
KZG_HD int pip_digit(const uint32_t h[4], int w, int c) {
    const int lo = c * w - 1;  // lowest bit read (bit -1 is zero)
    uint32_t v = 0;
#pragma unroll 1
    for (int t = 0; t <= c; t++) {
        const int bit = lo + t;
        if (bit >= 0 && bit < 128) v |= ((h[bit >> 5] >> (bit & 31)) & 1u) << t;
    }
    return (int)(v >> 1) + (int)(v & 1u) - (int)(((v >> c) & 1u) << c);
}

// Thread (i, w): the up to six entries of blob i in window w.  SCATTER = false: count[bucket]++;
// SCATTER = true: entries[cursor[bucket]++] = entry (cursor starts at the bucket's offset).
template <bool SCATTER>
KZG_HD void pip_digits_thread(uint32_t i, int w, const PipPlan &pl, const uint32_t *halves, uint32_t *counters, uint32_t *entries) {
    int d[4];
#pragma unroll
    for (int t = 0; t < 4; t++) d[t] = pip_digit(halves + 16ull * i + 4 * t, w, pl.c);
    // (lane, half used, entry flags)
    const int lane_of[6] = {0, 0, 1, 1, 2, 2};
    const int half_of[6] = {0, 1, 0, 1, 2, 3};
    const uint32_t flags_of[6] = {0, KZG_PIP_PSI, KZG_PIP_CPT, KZG_PIP_CPT | KZG_PIP_PSI, 0, KZG_PIP_PSI};
#pragma unroll
    for (int e = 0; e < 6; e++) {
        const int dg = d[half_of[e]];
        if (dg == 0) continue;
        const uint32_t mag = (uint32_t)(dg < 0 ? -dg : dg);
        const uint32_t bucket = ((uint32_t)lane_of[e] * (uint32_t)pl.W + (uint32_t)w) * pl.nbw + (mag - 1);
#if defined(__CUDA_ARCH__)
        const uint32_t pos = atomicAdd(counters + bucket, 1u);
#else
        const uint32_t pos = counters[bucket]++;
#endif
        if (SCATTER) entries[pos] = i | flags_of[e] | (dg < 0 ? KZG_PIP_NEG : 0u);
    }
}

KZG_HD fp_t fp_beta2() {
    fp_t beta, b2;
    constexpr uint32_t bm[12] = {FP_BETA_MONT_LIMBS};
#pragma unroll
    for (int i = 0; i < 12; i++) beta.l[i] = bm[i];
    fe_sqr(b2, beta);
    return b2;
}

// One slot: the sum of the entries of its bucket (points at infinity contribute nothing), Jacobian.  Slots and
// buckets coincide except in the last window, where 2^top_shift consecutive slots share one bucket's list
// (slot s: bucket s >> top_shift, every 2^top_shift-th entry starting at s mod 2^top_shift).
KZG_HD void pip_bucket_thread(uint32_t slot, const PipPlan &pl, const uint32_t *offsets, const uint32_t *entries, const g1_affine_t *cpts,
                              const g1_affine_t *ppts, g1_jac_t *out) {
    const fp_t beta2 = fp_beta2();
    g1_jac_t acc;
    g1j_set_inf(acc);
    const uint32_t in_window = slot & (pl.nbw - 1), lw = slot >> (pl.c - 1);  // lane * W + w
    const int sh = (int)(lw % (uint32_t)pl.W) == pl.W - 1 ? pl.top_shift : 0;
    const uint32_t bucket = lw * pl.nbw + (in_window >> sh), step = 1u << sh;
    const uint32_t lo = offsets[bucket] + (in_window & (step - 1)), hi = offsets[bucket + 1];
#pragma unroll 1
    for (uint32_t e = lo; e < hi; e += step) {
        const uint32_t code = entries[e];
        g1_affine_t p = (code & KZG_PIP_CPT) ? cpts[code & KZG_PIP_IDX] : ppts[code & KZG_PIP_IDX];
        if (g1a_is_inf(p)) continue;
        if (code & KZG_PIP_PSI) fe_mul(p.x, p.x, beta2);
        if (code & KZG_PIP_NEG) fe_neg(p.y, p.y);
        g1j_add_affine(acc, acc, p.x, p.y);
    }
    out[slot] = acc;
}

// Row (o, p), p = c w + j: which slots feed it.  o = 0: lane 0; o = 1: lanes 1 and 2.  Per lane: the bucket numbers b in
// [1, 2^(bits-1)] with bit j set (bits = the window's width), times the slots per bucket.  Element t < pip_row_len.
KZG_HD uint32_t pip_row_per_lane(const PipPlan &pl, int w, int j) {
    const int bits = w == pl.W - 1 ? pl.top_bits : pl.c, sh = w == pl.W - 1 ? pl.top_shift : 0;
    return (j == bits - 1 ? 1u : 1u << (bits - 2)) << sh;
}
KZG_HD uint32_t pip_row_len(const PipPlan &pl, int o, int w, int j) { return (o ? 2u : 1u) * pip_row_per_lane(pl, w, j); }
KZG_HD uint32_t pip_row_slot(const PipPlan &pl, int o, int w, int j, uint32_t t) {
    const int bits = w == pl.W - 1 ? pl.top_bits : pl.c, sh = w == pl.W - 1 ? pl.top_shift : 0;
    const uint32_t per_lane = pip_row_per_lane(pl, w, j);
    const uint32_t lane = (o ? 1u : 0u) + t / per_lane, v = t % per_lane, u = v >> sh, q = v & ((1u << sh) - 1u);
    // the u-th number with bit j set: insert a 1 at position j of u (j = bits-1: the number 2^(bits-1) itself)
    const uint32_t b = j == bits - 1 ? 1u << (bits - 1) : (((u >> j) << (j + 1)) | (1u << j) | (u & ((1u << j) - 1u)));
    return (lane * (uint32_t)pl.W + (uint32_t)w) * pl.nbw + ((b - 1) << sh) + q;
}
// 2^p x (a row sum)
KZG_HD void pip_weight(g1_jac_t &s, uint32_t p) {
    if (g1j_is_inf(s)) return;
#pragma unroll 1
    for (uint32_t t = 0; t < p; t++) g1j_dbl(s, s);
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(64) k_pip_scalars(const uint8_t *__restrict__ zy, fr_t r_canon, uint64_t first, uint32_t count,
                                                    uint32_t *__restrict__ halves, fr_t *__restrict__ sy) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) pip_scalars_thread(i, zy, r_canon, first, halves, sy);
}
template <bool SCATTER>
__global__ void __launch_bounds__(128) k_pip_digits(PipPlan pl, uint32_t count, const uint32_t *__restrict__ halves,
                                                    uint32_t *__restrict__ counters, uint32_t *__restrict__ entries) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)count * pl.W) return;
    // adjacent threads = adjacent blobs of one window: the atomics of a warp spread over the window's buckets
    const int w = (int)(t / count);
    const uint32_t i = (uint32_t)(t - (uint64_t)w * count);
    pip_digits_thread<SCATTER>(i, w, pl, halves, counters, entries);
}
// counts[0 .. nb) -> offsets[0 .. nb] (exclusive prefix sums) and cursor[b] = offsets[b]; one block of 1024 threads
__global__ void __launch_bounds__(1024) k_pip_scan(const uint32_t *__restrict__ counts, uint32_t nb, uint32_t *__restrict__ offsets,
                                                   uint32_t *__restrict__ cursor) {
    __shared__ uint32_t sh[1024];
    const uint32_t per = (nb + 1023) / 1024, lo = threadIdx.x * per, hi = min(nb, lo + per);
    uint32_t sum = 0;
    for (uint32_t b = lo; b < hi; b++) sum += counts[b];
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        uint32_t v = threadIdx.x >= (uint32_t)d ? sh[threadIdx.x - d] : 0u;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = sh[threadIdx.x] - sum;
    for (uint32_t b = lo; b < hi; b++) {
        offsets[b] = run;
        cursor[b] = run;
        run += counts[b];
    }
    if (threadIdx.x == 1023) offsets[nb] = sh[1023];
}
__global__ void __launch_bounds__(64) k_pip_buckets(PipPlan pl, const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ entries,
                                                    const g1_affine_t *__restrict__ cpts, const g1_affine_t *__restrict__ ppts,
                                                    g1_jac_t *__restrict__ out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < pl.buckets) pip_bucket_thread(b, pl, offsets, entries, cpts, ppts, out);
}
// block (o, p): rows[o * pl.rows + p] = 2^p S_p of output o
#define KZG_PIP_ROW_THREADS 128
__global__ void __launch_bounds__(KZG_PIP_ROW_THREADS) k_pip_rows(PipPlan pl, const g1_jac_t *__restrict__ buckets, g1_jac_t *__restrict__ rows) {
    __shared__ g1_jac_t red[KZG_PIP_ROW_THREADS];
    const int o = blockIdx.x / pl.rows;
    const uint32_t p = blockIdx.x - (uint32_t)o * pl.rows;
    const int w = (int)(p / (uint32_t)pl.c), j = (int)(p - (uint32_t)w * pl.c);
    const uint32_t len = pip_row_len(pl, o, w, j);
    g1_jac_t acc;
    g1j_set_inf(acc);
#pragma unroll 1
    for (uint32_t t = threadIdx.x; t < len; t += KZG_PIP_ROW_THREADS) {
        g1_jac_t q = buckets[pip_row_slot(pl, o, w, j, t)];
        g1j_add(acc, acc, q);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = KZG_PIP_ROW_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s && threadIdx.x + s < len) {
            g1_jac_t a = red[threadIdx.x], b = red[threadIdx.x + s];
            g1j_add(a, a, b);
            red[threadIdx.x] = a;
        }
        __syncthreads();
    }
    // the weight 2^p: up to 128 dependent doublings, by three lanes (g1j_dbl_n_coop3, g1.cuh)
    if (threadIdx.x < 3) {
        g1_jac_t s = red[0];
        g1j_dbl_n_coop3(s, p, 0x7u, 0, (int)threadIdx.x);
        if (threadIdx.x == 0) rows[blockIdx.x] = s;
    }
}
#endif  // __CUDACC__

}  // namespace kzg

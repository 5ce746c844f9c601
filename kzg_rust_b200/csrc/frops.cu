// frops.cu -- launch functions of the scalar-field kernels (frpath.cuh): Fiat-Shamir challenges, barycentric
// evaluation + quotient, and the r-power terms and sums of batch verification.
#include <algorithm>

#include "internal.h"
#include "frpath.cuh"
#include "pippenger.cuh"

using namespace kzg;

int fr_setup_roots_device(int n, fr_t **d_roots, cudaStream_t stream) { return fr_setup_roots(n, d_roots, stream); }

template <int G>
static void launch_challenge_group(cudaStream_t st, const uint8_t *d_blobs, const uint8_t *d_commitments, size_t count, int n, fr_t *d_z,
                                   uint32_t blk_begin = 0, uint32_t blk_end = 0xffffffffu, uint32_t *d_state = nullptr) {
    const size_t warps = (count * G + 31) / 32;
    k_challenge_group<G><<<blocks_for(warps, 4), 128, 0, st>>>(d_blobs, d_commitments, (uint32_t)count, n, d_z, 1u, blk_begin, blk_end,
                                                               reinterpret_cast<uint4 *>(d_state));
}
// The most lanes per blob that still leave every warp of the launch a scheduler of its own (frpath.cuh).  The validation of
// the chunk's points (`beside` blocks of four warps, one thread per point) runs beside the hash and is just as latency-bound:
// as long as both together have at most one block per SM, no scheduler carries two warps.  (Measured on whole calls,
// tools/hash_form_ab.py: a hash block sharing an SM with a validation block runs 1.5 - 2x longer.)
// A chunk of a larger call (call_blobs > count) runs beside the other lane's chunk and gets half of the SMs; the 4,096-blob
// chunks of the proof path end at one thread per blob that way (where instructions per blob count and latency does not), the
// 1,024-blob chunks of host verification at four lanes per blob.
static int challenge_lanes_per_blob(size_t count, size_t call_blobs, size_t beside, int sms) {
    // a chunk of a larger call shares the GPU with the other lane's chunk: half of the SMs each
    const size_t mine = call_blobs > count ? (size_t)sms / 2 : (size_t)sms;
    const size_t budget = mine > beside ? mine - beside : 0;
    int g = 32;
    while (g > 1 && (count * g + 127) / 128 > budget) g >>= 1;
    return g;
}
int fr_challenge_form(size_t count, size_t call_blobs, int sms) {
    const int g = env_int("KZG_B200_CHALLENGE_G", 0);  // forces a form
    return g > 0 ? g : challenge_lanes_per_blob(count, call_blobs, (2 * count + 127) / 128, sms);
}
int fr_launch_challenge(cudaStream_t st, const uint8_t *d_blobs, const uint8_t *d_commitments, size_t count, int n, fr_t *d_z,
                        size_t call_blobs, int sms) {
    if (count == 0) return KZG_B200_OK;
    const int g = fr_challenge_form(count, call_blobs, sms);
    switch (g) {
        case 32: launch_challenge_group<32>(st, d_blobs, d_commitments, count, n, d_z); break;
        case 16: launch_challenge_group<16>(st, d_blobs, d_commitments, count, n, d_z); break;
        case 8: launch_challenge_group<8>(st, d_blobs, d_commitments, count, n, d_z); break;
        case 4: launch_challenge_group<4>(st, d_blobs, d_commitments, count, n, d_z); break;
        case 2: launch_challenge_group<2>(st, d_blobs, d_commitments, count, n, d_z); break;
        default: k_challenge<<<blocks_for(count, 64), 64, 0, st>>>(d_blobs, d_commitments, (uint32_t)count, n, d_z);
    }
    CU(cudaGetLastError());
    return KZG_B200_OK;
}
// Blocks [blk_begin, blk_end) of the message only (grouped forms: g lanes per blob, g >= 2 as fr_challenge_form returned it;
// blk_begin a multiple of 32).  d_state: count x 8 words, the running state between the launches of one hash.
int fr_launch_challenge_range(cudaStream_t st, const uint8_t *d_blobs, const uint8_t *d_commitments, size_t count, int n, fr_t *d_z,
                              int g, uint32_t blk_begin, uint32_t blk_end, uint32_t *d_state) {
    if (count == 0) return KZG_B200_OK;
    if (!d_state || (blk_begin & 31u) || blk_begin >= blk_end) return KZG_B200_BAD_ARGS;
    switch (g) {
        case 32: launch_challenge_group<32>(st, d_blobs, d_commitments, count, n, d_z, blk_begin, blk_end, d_state); break;
        case 16: launch_challenge_group<16>(st, d_blobs, d_commitments, count, n, d_z, blk_begin, blk_end, d_state); break;
        case 8: launch_challenge_group<8>(st, d_blobs, d_commitments, count, n, d_z, blk_begin, blk_end, d_state); break;
        case 4: launch_challenge_group<4>(st, d_blobs, d_commitments, count, n, d_z, blk_begin, blk_end, d_state); break;
        case 2: launch_challenge_group<2>(st, d_blobs, d_commitments, count, n, d_z, blk_begin, blk_end, d_state); break;
        default: return KZG_B200_BAD_ARGS;
    }
    CU(cudaGetLastError());
    return KZG_B200_OK;
}
int fr_launch_load_scalars(cudaStream_t st, const uint8_t *d_in, size_t count, fr_t *d_out, int32_t *d_status) {
    if (count == 0) return KZG_B200_OK;
    k_load_scalars<<<blocks_for(count, 128), 128, 0, st>>>(d_in, (uint32_t)count, d_out, d_status);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}
int fr_launch_eval(cudaStream_t st, int quotient, const uint8_t *d_blobs, const fr_t *d_z, const fr_t *d_roots, int n, fr_t *d_inv,
                   fr_t *d_poly, uint8_t *d_zy, int32_t *d_status, size_t count) {
    if (count == 0) return KZG_B200_OK;
    if (quotient)
        k_eval_quotient<true><<<(unsigned)count, KZG_EVAL_THREADS, 0, st>>>(d_blobs, d_z, d_roots, n, d_inv, d_poly, d_zy, d_status);
    else
        k_eval_quotient<false><<<(unsigned)count, KZG_EVAL_THREADS, 0, st>>>(d_blobs, d_z, d_roots, n, d_inv, d_poly, d_zy, d_status);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}
int fr_launch_verify_terms(cudaStream_t st, const g1_affine_t *d_cpts, const g1_affine_t *d_ppts, const uint8_t *d_zy,
                           const fr_t &r_canon, uint64_t first, size_t count, g1_jac_t *d_terms, fr_t *d_sy) {
    if (count == 0) return KZG_B200_OK;
    k_verify_terms<<<blocks_for(3 * count, 96), 96, 0, st>>>(d_cpts, d_ppts, d_zy, r_canon, first, (uint32_t)count, d_terms, d_sy);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}
// sums[0] = sum V_i, sums[1] = sum (U1_i + U2_i) (affine), the scalar sum, and the 224-byte partial record.
// d_partials: scratch for 2 * KZG_VERIFY_SUM_BLOCKS Jacobian points.
int fr_launch_verify_sums(cudaStream_t st, const g1_jac_t *d_terms, const fr_t *d_sy, size_t count, g1_affine_t *d_sums,
                          fr_t *d_sy_total, uint8_t *d_partial, g1_jac_t *d_partials) {
    const uint32_t n = (uint32_t)count;
    uint32_t per = (uint32_t)std::min<size_t>(KZG_VERIFY_SUM_BLOCKS, (2 * count + 4 * KZG_JSUM_THREADS - 1) / (4 * KZG_JSUM_THREADS));
    if (per > 1) {
        k_jac_sum_slices<<<2 * per, KZG_JSUM_THREADS, 0, st>>>(d_terms, n, 2 * n, per, d_partials);
        k_jac_sum<<<2, KZG_JSUM_THREADS, 0, st>>>(d_partials, per, per, d_sums);
    } else {
        k_jac_sum<<<2, KZG_JSUM_THREADS, 0, st>>>(d_terms, n, 2 * n, d_sums);
    }
    k_fr_sum<<<1, 256, 0, st>>>(d_sy, n, d_sy_total);
    k_write_partial<<<1, 32, 0, st>>>(d_sums, d_sy_total, d_partial);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// The same sums by the bucket method (pippenger.cuh): sy[i] = r_i y_i, then rows[o * pl.rows + p] = 2^p S_p of output o;
// fr_launch_verify_sums_rows adds them.  d_ws: fr_pip_workspace_bytes(count) bytes, 256-byte aligned.
size_t fr_pip_workspace_bytes(size_t count) {
    // the entry lists are longest at c = 1 (129 windows), the bucket arrays largest at c = KZG_PIP_MAX_C
    const PipPlan small = PipPlan::make(count, 1), big = PipPlan::make(count, KZG_PIP_MAX_C);
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    return up(16 * sizeof(uint32_t) * count) + 3 * up((big.buckets + 1) * sizeof(uint32_t)) + up(small.max_entries * sizeof(uint32_t)) +
           up(big.buckets * sizeof(g1_jac_t)) + up(2 * (size_t)KZG_PIP_SCALAR_BITS * sizeof(g1_jac_t));
}
int fr_launch_verify_pippenger(cudaStream_t st, const g1_affine_t *d_cpts, const g1_affine_t *d_ppts, const uint8_t *d_zy,
                               const fr_t &r_canon, uint64_t first, size_t count, int force_c, uint8_t *d_ws, fr_t *d_sy,
                               g1_affine_t *d_sums, fr_t *d_sy_total, uint8_t *d_partial) {
    if (count == 0 || count > KZG_PIP_IDX) return KZG_B200_BAD_ARGS;
    const PipPlan pl = PipPlan::make(count, force_c), big = PipPlan::make(count, KZG_PIP_MAX_C), small = PipPlan::make(count, 1);
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    uint32_t *halves = reinterpret_cast<uint32_t *>(d_ws);
    uint32_t *counts = reinterpret_cast<uint32_t *>(d_ws + up(16 * sizeof(uint32_t) * count));
    uint32_t *offsets = counts + up((big.buckets + 1) * sizeof(uint32_t)) / sizeof(uint32_t);
    uint32_t *cursor = offsets + up((big.buckets + 1) * sizeof(uint32_t)) / sizeof(uint32_t);
    uint32_t *entries = cursor + up((big.buckets + 1) * sizeof(uint32_t)) / sizeof(uint32_t);
    g1_jac_t *buckets = reinterpret_cast<g1_jac_t *>(reinterpret_cast<uint8_t *>(entries) + up(small.max_entries * sizeof(uint32_t)));
    g1_jac_t *rows = reinterpret_cast<g1_jac_t *>(reinterpret_cast<uint8_t *>(buckets) + up(big.buckets * sizeof(g1_jac_t)));
    const uint32_t n = (uint32_t)count;
    CU(cudaMemsetAsync(counts, 0, pl.buckets * sizeof(uint32_t), st));
    k_pip_scalars<<<blocks_for(count, 64), 64, 0, st>>>(d_zy, r_canon, first, n, halves, d_sy);
    const uint64_t dt = (uint64_t)count * pl.W;
    k_pip_digits<false><<<blocks_for(dt, 128), 128, 0, st>>>(pl, n, halves, counts, nullptr);
    k_pip_scan<<<1, 1024, 0, st>>>(counts, pl.buckets, offsets, cursor);
    k_pip_digits<true><<<blocks_for(dt, 128), 128, 0, st>>>(pl, n, halves, cursor, entries);
    k_pip_buckets<<<blocks_for(pl.buckets, 64), 64, 0, st>>>(pl, offsets, entries, d_cpts, d_ppts, buckets);
    k_pip_rows<<<2 * pl.rows, KZG_PIP_ROW_THREADS, 0, st>>>(pl, buckets, rows);
    k_jac_sum<<<2, KZG_JSUM_THREADS, 0, st>>>(rows, pl.rows, pl.rows, d_sums);
    k_fr_sum<<<1, 256, 0, st>>>(d_sy, n, d_sy_total);
    k_write_partial<<<1, 32, 0, st>>>(d_sums, d_sy_total, d_partial);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// frops.cu -- launch functions of the scalar-field kernels (frpath.cuh): Fiat-Shamir challenges, barycentric
// evaluation + quotient, and the r-power terms and sums of batch verification.
#include <algorithm>

#include "internal.h"
#include "frpath.cuh"

using namespace kzg;

int fr_setup_roots_device(int n, fr_t **d_roots, cudaStream_t stream) { return fr_setup_roots(n, d_roots, stream); }

int fr_launch_challenge(cudaStream_t st, const uint8_t *d_blobs, const uint8_t *d_commitments, size_t count, int n, fr_t *d_z) {
    if (count == 0) return KZG_B200_OK;
    if (count <= KZG_CHALLENGE_WARP_MAX)
        k_challenge_warp<<<blocks_for(count, 4), 128, 0, st>>>(d_blobs, d_commitments, (uint32_t)count, n, d_z);
    else
        k_challenge<<<blocks_for(count, 64), 64, 0, st>>>(d_blobs, d_commitments, (uint32_t)count, n, d_z);
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

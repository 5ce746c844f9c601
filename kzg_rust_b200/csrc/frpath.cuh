// frpath.cuh -- scalar-field (Fr) side of the blob path: roots of unity, the Fiat-Shamir
// challenge, barycentric evaluation and the quotient polynomial.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "blobpath.cuh"

namespace kzg {

// reference compute_roots_of_unity (src/kzg.rs:764-799) + expand_root_of_unity (:734-761):
// w = 7^((r-1)/n); [w^0 .. w^(n-1)] in bit-reversed order, Montgomery form, on the device.
inline int fr_setup_roots(int n, fr_t **d_roots, cudaStream_t stream) {
    fr_t w;
    {
        const uint32_t w4096[8] = {FR_ROOT_OF_UNITY_4096_LIMBS}, w4[8] = {FR_ROOT_OF_UNITY_4_LIMBS};
        for (int i = 0; i < 8; i++) w.l[i] = n == 4096 ? w4096[i] : w4[i];
    }
    fe_to_mont(w, w);
    std::vector<fr_t> nat(n), rev(n);
    nat[0] = fe_one<FrParams>();
    for (int i = 1; i < n; i++) fe_mul(nat[i], nat[i - 1], w);
    fr_t chk;
    fe_mul(chk, nat[n - 1], w);
    if (!fe_eq(chk, nat[0])) return KZG_INTERNAL;  // w^n must be 1
    for (int i = 0; i < n; i++) {
        uint32_t r = 0, v = (uint32_t)i;
        for (int o = n; o > 1; o >>= 1) { r = (r << 1) | (v & 1); v >>= 1; }
        rev[i] = nat[r];
    }
    if (cudaMalloc(d_roots, (size_t)n * sizeof(fr_t)) != cudaSuccess) return KZG_CUDA;
    if (cudaMemcpyAsync(*d_roots, rev.data(), (size_t)n * sizeof(fr_t), cudaMemcpyHostToDevice, stream) != cudaSuccess) return KZG_CUDA;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return KZG_CUDA;
    return KZG_OK;
}

}  // namespace kzg

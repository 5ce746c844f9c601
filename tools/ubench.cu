// ubench.cu -- issue-rate micro-benchmarks for the integer-multiply roofline of the blob path
// (not part of the product).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
// Every kernel keeps its multiplicands data-dependent so ptxas cannot strength-reduce them.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../kzg_rust_b200/csrc/fields.cuh"

using namespace kzg;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// ---- plain IMAD (32-bit), 8 independent chains
__global__ void k_imad(uint32_t *out, int iters) {
    uint32_t x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 2654435761u + i;
    uint32_t b = blockIdx.x * 40503u + 3;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = x[i] * x[i] + b;
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- IMAD.WIDE (32x32+64), 8 independent chains, multiplicand = low word of the accumulator
__global__ void k_imad_wide(uint64_t *out, int iters) {
    uint64_t x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 2654435761u + i;
    uint32_t b = blockIdx.x * 40503u + 3;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t hi = (uint32_t)(x[(i + 1) & 7] >> 32);
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(hi), "r"(b));
            }
    }
    uint64_t s = 0;
    for (int i = 0; i < 8; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- IMAD.WIDE.X: carry chains of length 8 (mad.lo.cc / madc.hi.cc pairs), 2 independent chains
__global__ void k_imad_wide_x(uint32_t *out, int iters) {
    uint32_t a[16], c[16];
    for (int i = 0; i < 16; i++) { a[i] = threadIdx.x * 2654435761u + i; c[i] = i; }
    uint32_t b = blockIdx.x * 40503u + 3;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t m = c[0] ^ b;
            // two chains of 4 wide MACs each over the same accumulator halves
            asm volatile(
                "mad.lo.cc.u32 %0, %16, %24, %0;\n\tmadc.hi.cc.u32 %1, %16, %24, %1;\n\t"
                "madc.lo.cc.u32 %2, %17, %24, %2;\n\tmadc.hi.cc.u32 %3, %17, %24, %3;\n\t"
                "madc.lo.cc.u32 %4, %18, %24, %4;\n\tmadc.hi.cc.u32 %5, %18, %24, %5;\n\t"
                "madc.lo.cc.u32 %6, %19, %24, %6;\n\tmadc.hi.u32 %7, %19, %24, %7;\n\t"
                "mad.lo.cc.u32 %8, %20, %24, %8;\n\tmadc.hi.cc.u32 %9, %20, %24, %9;\n\t"
                "madc.lo.cc.u32 %10, %21, %24, %10;\n\tmadc.hi.cc.u32 %11, %21, %24, %11;\n\t"
                "madc.lo.cc.u32 %12, %22, %24, %12;\n\tmadc.hi.cc.u32 %13, %22, %24, %13;\n\t"
                "madc.lo.cc.u32 %14, %23, %24, %14;\n\tmadc.hi.u32 %15, %23, %24, %15;"
                : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]),
                  "+r"(c[8]), "+r"(c[9]), "+r"(c[10]), "+r"(c[11]), "+r"(c[12]), "+r"(c[13]), "+r"(c[14]), "+r"(c[15])
                : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(a[12]), "r"(a[14]), "r"(m));
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 16; i++) s ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- DFMA, 8 independent chains
__global__ void k_dfma(double *out, int iters) {
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = 1.0 + threadIdx.x * 1e-9 + i;
    double b = 1.0 + blockIdx.x * 1e-12;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = __fma_rn(x[i], b, x[i]);
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- IMAD + DFMA interleaved (do the two pipes overlap?)
__global__ void k_imad_dfma(double *out, int iters) {
    double x[8];
    uint32_t y[8];
    for (int i = 0; i < 8; i++) { x[i] = 1.0 + threadIdx.x * 1e-9 + i; y[i] = threadIdx.x * 2654435761u + i; }
    double b = 1.0 + blockIdx.x * 1e-12;
    uint32_t bi = blockIdx.x * 40503u + 3;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) { x[i] = __fma_rn(x[i], b, x[i]); y[i] = y[i] * y[i] + bi; }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- dependent Fp / Fr multiplications (the product's fe_mul), `CH` independent chains per thread
template <class F, int CH>
__global__ void k_femul(F *out, int iters) {
    F x[CH], y[CH];
    for (int c = 0; c < CH; c++) {
        x[c] = fe_one<typename std::conditional<F::N == 12, FpParams, FrParams>::type>();
        y[c] = x[c];
        x[c].l[0] += threadIdx.x + c;
        y[c].l[1] ^= blockIdx.x + 7 * c;
    }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) fe_mul(x[c], x[c], y[c]);
#pragma unroll
        for (int c = 0; c < CH; c++) fe_mul(y[c], y[c], x[c]);
    }
    F r = x[0];
    for (int c = 0; c < CH; c++) { fe_add(r, r, x[c]); fe_add(r, r, y[c]); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class F, int CH>
__global__ void k_fesqr(F *out, int iters) {
    F x[CH];
    for (int c = 0; c < CH; c++) {
        x[c] = fe_one<typename std::conditional<F::N == 12, FpParams, FrParams>::type>();
        x[c].l[0] += threadIdx.x + c;
        x[c].l[1] ^= blockIdx.x + 7 * c;
    }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) fe_sqr(x[c], x[c]);
#pragma unroll
        for (int c = 0; c < CH; c++) fe_sqr(x[c], x[c]);
    }
    F r = x[0];
    for (int c = 0; c < CH; c++) fe_add(r, r, x[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K>
static double time_ms(K launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch(true);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch(false);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    void *buf;
    CK(cudaMalloc(&buf, (size_t)sms * 32 * 1024 * 64));
    printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, sms);
    const int iters = 2048;
    {
        int blocks = sms * 8, tpb = 256;
        double ms = time_ms([&](bool warm) { k_imad<<<blocks, tpb>>>((uint32_t *)buf, warm ? 16 : iters); });
        printf("{\"probe\": \"imad\", \"per_s\": %.4e}\n", (double)blocks * tpb * iters * 64.0 / (ms * 1e-3));
        ms = time_ms([&](bool warm) { k_imad_wide<<<blocks, tpb>>>((uint64_t *)buf, warm ? 16 : iters); });
        printf("{\"probe\": \"imad_wide\", \"per_s\": %.4e}\n", (double)blocks * tpb * iters * 64.0 / (ms * 1e-3));
        ms = time_ms([&](bool warm) { k_imad_wide_x<<<blocks, tpb>>>((uint32_t *)buf, warm ? 16 : iters); });
        printf("{\"probe\": \"imad_wide_x (wide MACs in carry chains)\", \"per_s\": %.4e}\n", (double)blocks * tpb * iters * 32.0 / (ms * 1e-3));
        ms = time_ms([&](bool warm) { k_dfma<<<blocks, tpb>>>((double *)buf, warm ? 16 : iters); });
        printf("{\"probe\": \"dfma\", \"per_s\": %.4e}\n", (double)blocks * tpb * iters * 64.0 / (ms * 1e-3));
        ms = time_ms([&](bool warm) { k_imad_dfma<<<blocks, tpb>>>((double *)buf, warm ? 16 : iters); });
        printf("{\"probe\": \"imad+dfma pairs\", \"pairs_per_s\": %.4e}\n", (double)blocks * tpb * iters * 64.0 / (ms * 1e-3));
    }
    // fe_mul at several occupancies (blocks of 128 threads per SM)
    const int occ[] = {2, 3, 4, 6, 8, 12, 16};
    for (int o : occ) {
        int blocks = sms * o, tpb = 128, it = 512;
        double ms = time_ms([&](bool warm) { k_femul<fp_t, 1><<<blocks, tpb>>>((fp_t *)buf, warm ? 8 : it); });
        printf("{\"probe\": \"fp_mul x1\", \"warps_per_sm\": %d, \"mul_per_s\": %.4e}\n", o * 4, (double)blocks * tpb * it * 2.0 / (ms * 1e-3));
        ms = time_ms([&](bool warm) { k_femul<fp_t, 2><<<blocks, tpb>>>((fp_t *)buf, warm ? 8 : it); });
        printf("{\"probe\": \"fp_mul x2\", \"warps_per_sm\": %d, \"mul_per_s\": %.4e}\n", o * 4, (double)blocks * tpb * it * 4.0 / (ms * 1e-3));
        ms = time_ms([&](bool warm) { k_fesqr<fp_t, 1><<<blocks, tpb>>>((fp_t *)buf, warm ? 8 : it); });
        printf("{\"probe\": \"fp_sqr x1\", \"warps_per_sm\": %d, \"mul_per_s\": %.4e}\n", o * 4, (double)blocks * tpb * it * 2.0 / (ms * 1e-3));
        ms = time_ms([&](bool warm) { k_femul<fr_t, 1><<<blocks, tpb>>>((fr_t *)buf, warm ? 8 : it); });
        printf("{\"probe\": \"fr_mul x1\", \"warps_per_sm\": %d, \"mul_per_s\": %.4e}\n", o * 4, (double)blocks * tpb * it * 2.0 / (ms * 1e-3));
    }
    return 0;
}

// ubench.cu -- issue-rate micro-benchmarks for the integer-multiply roofline of the blob path
// (not part of the product).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
// Every kernel keeps its multiplicands data-dependent so ptxas cannot strength-reduce them.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../kzg_rust_b200/csrc/fields.cuh"
#include "experiments/fp_hybrid.cuh"

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

// ---- do wide integer multiply-adds (IMAD.WIDE.U32) and FP64 instructions overlap?
// MODE 0: wide only, 1: DFMA.RZ only, 2: DADD only, 3: wide + DFMA.RZ, 4: wide + DADD, 5: wide + DFMA.RZ + DADD
template <int MODE>
__global__ void k_wide_fp64(double *out, int iters) {
    uint64_t w[8];
    double x[8], z[8];
    for (int i = 0; i < 8; i++) { w[i] = threadIdx.x * 2654435761u + i; x[i] = 1.0 + threadIdx.x * 1e-9 + i; z[i] = 3.0 + i; }
    uint32_t b = blockIdx.x * 40503u + 3;
    double fb = 1.0 + blockIdx.x * 1e-12;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0 || MODE >= 3) {
                    uint32_t hi = (uint32_t)(w[(i + 1) & 7] >> 32);
                    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(hi), "r"(b));
                }
                if (MODE == 1 || MODE == 3 || MODE == 5) x[i] = __fma_rz(x[i], fb, x[i]);
                if (MODE == 2 || MODE == 4 || MODE == 5) z[i] = __dadd_rn(z[i], fb);
            }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i] + z[i] + (double)w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- do ALU instructions issue beside wide multiply-adds for free?
// 16 IMAD.WIDE.U32.X (two carry chains, as in fe_mul) + NALU independent 32-bit adds (IADD3) per group;
// XCHAIN: the adds form one add.cc/addc.cc carry chain (IADD3.X) instead
template <int NALU, bool XCHAIN>
__global__ void k_wide_alu(uint32_t *out, int iters) {
    uint32_t a[8], c[16], e[32];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + i;
    for (int i = 0; i < 16; i++) c[i] = i;
    for (int i = 0; i < 32; i++) e[i] = threadIdx.x + i * 977u;
    uint32_t b = blockIdx.x * 40503u + 3;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t m = c[0] ^ b;
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
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(m));
            if (XCHAIN) {
                if (NALU > 0) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(e[0]) : "r"(b));
#pragma unroll
                for (int i = 1; i < NALU; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(e[i]) : "r"(b));
            } else {
#pragma unroll
                for (int i = 0; i < NALU; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(e[i]) : "r"(b));
            }
        }
    }
    uint32_t x = 0;
    for (int i = 0; i < 16; i++) x ^= c[i];
    for (int i = 0; i < 32; i++) x ^= e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// ---- 32-bit halves: mad.hi (IMAD.HI.U32) and mad.lo (IMAD) with and without FP64 work beside them
// MODE 0: mad.hi only, 1: mad.hi + DFMA, 2: mad.lo + mad.hi (one 32x32 product as two instructions),
// 3: mad.lo + mad.hi + DFMA + DADD, 4: mad.lo.cc / madc.hi.cc rows written lo-chain then hi-chain (unfusable) + DFMA
template <int MODE>
__global__ void k_halves_fp64(double *out, int iters) {
    uint32_t w[8], v[8];
    double x[8], z[8];
    for (int i = 0; i < 8; i++) { w[i] = threadIdx.x * 2654435761u + i; v[i] = w[i] ^ 0x9e3779b9u; x[i] = 1.0 + threadIdx.x * 1e-9 + i; z[i] = 3.0 + i; }
    uint32_t b = blockIdx.x * 40503u + 3;
    double fb = 1.0 + blockIdx.x * 1e-12;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 4) {
                uint32_t m = w[0] ^ b;
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %16, %0;\n\tmadc.lo.cc.u32 %1, %9, %16, %1;\n\tmadc.lo.cc.u32 %2, %10, %16, %2;\n\tmadc.lo.cc.u32 %3, %11, %16, %3;\n\t"
                    "madc.lo.cc.u32 %4, %12, %16, %4;\n\tmadc.lo.cc.u32 %5, %13, %16, %5;\n\tmadc.lo.cc.u32 %6, %14, %16, %6;\n\tmadc.lo.u32 %7, %15, %16, %7;\n\t"
                    "mad.hi.cc.u32 %1, %8, %16, %1;\n\tmadc.hi.cc.u32 %2, %9, %16, %2;\n\tmadc.hi.cc.u32 %3, %10, %16, %3;\n\tmadc.hi.cc.u32 %4, %11, %16, %4;\n\t"
                    "madc.hi.cc.u32 %5, %12, %16, %5;\n\tmadc.hi.cc.u32 %6, %13, %16, %6;\n\tmadc.hi.u32 %7, %14, %16, %7;"
                    : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7])
                    : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(m));
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = __fma_rz(x[i], fb, x[i]);
#pragma unroll
                for (int i = 0; i < 7; i++) z[i] = __dadd_rn(z[i], fb);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (MODE >= 2) w[i] = w[i] * b + v[(i + 1) & 7];
                    v[i] = __umulhi(v[i], b) + w[(i + 3) & 7];
                    if (MODE == 1 || MODE == 3) x[i] = __fma_rz(x[i], fb, x[i]);
                    if (MODE == 3) z[i] = __dadd_rn(z[i], fb);
                }
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i] + z[i] + (double)w[i] + (double)v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- fe_mul as a real call (code-size experiment: the addition kernel inlines 13 copies of it)
__device__ __noinline__ fp_t fp_mul_call(fp_t a, fp_t b) { fp_t r; fe_mul(r, a, b); return r; }
// ---- two-pipe multiplication (fp_hybrid.cuh): dependent chains, and a bit-for-bit check against fe_mul
template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) k_fpmul_variant(fp_t *out, int iters) {
    fp_t x = fe_one<FpParams>(), y = fp_const_b();
    x.l[0] += threadIdx.x;
    y.l[1] ^= blockIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { fe_mul(x, x, y); fe_mul(y, y, x); }
        if (MODE == 1) { fp_mul_hybrid(x, x, y); fp_mul_hybrid(y, y, x); }
        if (MODE == 2) { fp_sqr_hybrid(x, x); fp_sqr_hybrid(y, y); }
        if (MODE == 3) { fp_mul_hybrid(x, x, y); fe_mul(y, y, x); }
        if (MODE == 4) { x = fp_mul_call(x, y); y = fp_mul_call(y, x); }
    }
    fe_add(x, x, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_hybrid_check(unsigned long long *bad, int iters) {
    fp_t x = fe_one<FpParams>(), y = fp_const_b();
    x.l[0] += threadIdx.x * 977u + blockIdx.x;
    y.l[1] ^= blockIdx.x * 2654435761u + threadIdx.x;
    unsigned long long nbad = 0;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        fp_t a, b, c, d;
        fe_mul(a, x, y);
        fp_mul_hybrid(b, x, y);
        fe_mul(c, y, y);
        fp_sqr_hybrid(d, y);
        if (!fe_eq(a, b)) nbad++;
        if (!fe_eq(c, d)) nbad++;
        x = a;
        fe_add(y, c, x);
    }
    if (nbad) atomicAdd(bad, nbad);
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
    {
        int blocks = sms * 8, tpb = 256;
        const char *names[] = {"imad.wide only", "dfma.rz only", "dadd only", "imad.wide + dfma.rz", "imad.wide + dadd", "imad.wide + dfma.rz + dadd"};
        double ms;
#define RUNW(MODE) ms = time_ms([&](bool warm) { k_wide_fp64<MODE><<<blocks, tpb>>>((double *)buf, warm ? 16 : iters); }); \
        printf("{\"probe\": \"%s\", \"groups_per_s\": %.4e}\n", names[MODE], (double)blocks * tpb * iters * 64.0 / (ms * 1e-3));
        RUNW(0) RUNW(1) RUNW(2) RUNW(3) RUNW(4) RUNW(5)
#define RUNA(N, X, label) ms = time_ms([&](bool warm) { k_wide_alu<N, X><<<blocks, tpb>>>((uint32_t *)buf, warm ? 16 : iters); }); \
        printf("{\"probe\": \"%s\", \"wide_per_s\": %.4e}\n", label, (double)blocks * tpb * iters * 32.0 / (ms * 1e-3));
        RUNA(0, false, "16 wide.X + 0 alu") RUNA(4, false, "16 wide.X + 4 IADD3") RUNA(8, false, "16 wide.X + 8 IADD3") RUNA(16, false, "16 wide.X + 16 IADD3")
        RUNA(32, false, "16 wide.X + 32 IADD3") RUNA(8, true, "16 wide.X + 8 IADD3.X chain") RUNA(16, true, "16 wide.X + 16 IADD3.X chain")
        const char *hn[] = {"mad.hi only", "mad.hi + dfma.rz", "mad.lo + mad.hi", "mad.lo + mad.hi + dfma.rz + dadd", "8 lo.cc + 7 hi.cc chains + 8 dfma.rz + 7 dadd"};
#define RUNH(MODE) ms = time_ms([&](bool warm) { k_halves_fp64<MODE><<<blocks, tpb>>>((double *)buf, warm ? 16 : iters); }); \
        printf("{\"probe\": \"%s\", \"groups_per_s\": %.4e}\n", hn[MODE], (double)blocks * tpb * iters * (MODE == 4 ? 8.0 : 64.0) / (ms * 1e-3));
        RUNH(0) RUNH(1) RUNH(2) RUNH(3) RUNH(4)
    }
    {
        unsigned long long *bad;
        CK(cudaMalloc(&bad, 8));
        CK(cudaMemset(bad, 0, 8));
        k_hybrid_check<<<sms * 4, 128>>>(bad, 256);
        CK(cudaDeviceSynchronize());
        unsigned long long h = 0;
        CK(cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost));
        printf("{\"probe\": \"two-pipe mul/sqr vs fe_mul\", \"cases\": %llu, \"mismatches\": %llu}\n", (unsigned long long)sms * 4 * 128 * 256 * 2, h);
        const int it = 512, tpb = 128;
        for (int o : {2, 3, 4, 6}) {
            int blocks = sms * o;
            double ms;
#define RUNV(MODE, MINB, label) \
            ms = time_ms([&](bool warm) { k_fpmul_variant<MODE, MINB><<<blocks, tpb>>>((fp_t *)buf, warm ? 8 : it); }); \
            printf("{\"probe\": \"%s\", \"launch_bounds_min_blocks\": %d, \"warps_per_sm\": %d, \"mul_per_s\": %.4e}\n", label, MINB, o * 4, (double)blocks * tpb * it * 2.0 / (ms * 1e-3));
            RUNV(0, 3, "fe_mul (IMAD only)")
            RUNV(4, 3, "fe_mul through a __noinline__ call")
            RUNV(1, 3, "fp_mul_hybrid (FP64 product + IMAD redc)")
            RUNV(2, 3, "fp_sqr_hybrid")
            RUNV(3, 3, "alternating hybrid / IMAD-only")
            if (o >= 4) {
                RUNV(1, 4, "fp_mul_hybrid (FP64 product + IMAD redc)")
                RUNV(3, 4, "alternating hybrid / IMAD-only")
            }
            if (o >= 6) { RUNV(1, 6, "fp_mul_hybrid (FP64 product + IMAD redc)") }
        }
    }
    // fe_mul at several occupancies (blocks of 128 threads per SM)
    const int occ[] = {2, 3, 4};
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

// g1ops.cu -- the G1 kernels around the MSM: point decoding and subgroup checks (reference `validate_kzg_g1`,
// src/utils.rs:282-315), the Horner pass + 48-byte compression (`bytes_from_g1`, src/utils.rs:221-227),
// the roofline micro-benchmarks and the device-side test aids.
#include "internal.h"

using namespace kzg;

// status slot of point i is i % status_mod (commitments and proofs of one chunk are decoded by one
// launch and share the per-blob status)
__global__ void k_decode_g1(const uint8_t *in, g1_affine_t *out, int32_t *status, uint32_t count, int check_subgroup,
                            uint32_t status_mod) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    g1_affine_t p;
    int rc = g1_decode_thread(p, in + 48ull * i, check_subgroup != 0);
    out[i] = p;
    if (rc != KZG_OK && status) atomicMax(status + (i % status_mod), rc);
}
int g1_launch_decode(cudaStream_t st, const uint8_t *d_in, g1_affine_t *d_out, int32_t *d_status, size_t count, int check_subgroup,
                     size_t status_mod) {
    if (count == 0) return KZG_B200_OK;
    k_decode_g1<<<blocks_for(count, 128), 128, 0, st>>>(d_in, d_out, d_status, (uint32_t)count, check_subgroup, (uint32_t)status_mod);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

__global__ void k_decode_g1_pair(const uint8_t *in_a, const uint8_t *in_b, g1_affine_t *out_a, g1_affine_t *out_b, int32_t *status,
                                 uint32_t count, int check_subgroup) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * count) return;
    const bool second = t >= count;
    const uint32_t i = second ? t - count : t;
    g1_affine_t p;
    int rc = g1_decode_thread(p, (second ? in_b : in_a) + 48ull * i, check_subgroup != 0);
    (second ? out_b : out_a)[i] = p;
    if (rc != KZG_OK && status) atomicMax(status + i, rc);
}
int g1_launch_decode2(cudaStream_t st, const uint8_t *d_commitments, const uint8_t *d_proofs, g1_affine_t *d_cpts, g1_affine_t *d_ppts,
                      int32_t *d_status, size_t count, int check_subgroup) {
    if (count == 0) return KZG_B200_OK;
    k_decode_g1_pair<<<blocks_for(2 * count, 128), 128, 0, st>>>(d_commitments, d_proofs, d_cpts, d_ppts, d_status, (uint32_t)count, check_subgroup);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// the subgroup check alone, on points that k_decode_g1_pair decompressed without it (phase B of a verification on new
// data runs it beside the bucket method instead of in front of it)
// Three lanes per point (g1a_in_subgroup_coop3): ten points per warp, lanes 30 and 31 idle.  This launch is on the critical
// path of verify_kzg_proof (two points), where the single-thread test was 1.2 ms.
__global__ void __launch_bounds__(128) k_subgroup_pair(const g1_affine_t *pts_a, const g1_affine_t *pts_b, int32_t *status, uint32_t count) {
    const uint32_t lane = threadIdx.x & 31, warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (lane >= 30 || warp * 10 >= 2 * count) return;
    const uint32_t role = lane % 3, leader = lane - role;
    // a warp's last groups may be past the end: they test the last point again and write nothing
    const uint32_t t = min(warp * 10 + lane / 3, 2 * count - 1);
    const bool live = warp * 10 + lane / 3 < 2 * count;
    const uint32_t i = t >= count ? t - count : t;
    g1_affine_t p = (t >= count ? pts_b : pts_a)[i];
    const bool inf = g1a_is_inf(p);
    if (inf) { fe_set_zero(p.x); fe_set_zero(p.y); }  // the marker is not a field element; the result is ignored
    const bool ok = g1a_in_subgroup_coop3(p, 0x3fffffffu, (int)leader, (int)role);
    if (role == 0 && live && !inf && !ok) atomicMax(status + i, (int)KZG_BADARGS);
}
int g1_launch_subgroup2(cudaStream_t st, const g1_affine_t *d_cpts, const g1_affine_t *d_ppts, int32_t *d_status, size_t count) {
    if (count == 0) return KZG_B200_OK;
    k_subgroup_pair<<<blocks_for((2 * count + 9) / 10, 4), 128, 0, st>>>(d_cpts, d_ppts, d_status, (uint32_t)count);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// sums of blob i (W affine points at sums[j*stride + i]) -> Horner -> 48-byte compressed point
__global__ void __launch_bounds__(64) k_horner_compress(const g1_affine_t *sums, size_t stride, int W, const int32_t *status,
                                                        uint8_t *out, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint8_t buf[48];
    if (status && status[i] != KZG_OK) {
        for (int k = 0; k < 48; k++) buf[k] = 0;
    } else {
        g1_affine_t p;
        horner_thread(p, sums + i, stride, 1, W);
        g1a_compress(buf, p);
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(out + 48ull * i);
#pragma unroll
    for (int k = 0; k < 12; k++)
        o[k] = (uint32_t)buf[4 * k] | ((uint32_t)buf[4 * k + 1] << 8) | ((uint32_t)buf[4 * k + 2] << 16) | ((uint32_t)buf[4 * k + 3] << 24);
}
// The same with one WARP per blob, for small batches where the pass is pure latency (254 dependent doublings and
// additions per blob, 4.5 ms): lane k folds the sums 8k .. 8k + 7 (W = 255; 2k, 2k + 1 for the 64 sums of the latency
// comb), then five rounds join the 32 partial results, T_k + 2^(8 s) T_(k + s) -- the doublings stay on the critical path (255 of them), the additions leave it (12
// instead of 254), and from the second round on three lanes share every doubling.
#define KZG_HORNER_WARP_MAX 1024
// JAC: the sums are Jacobian points (the small-batch form of the MSM, msm_run_small), canonical.
template <bool JAC>
__global__ void __launch_bounds__(128) k_horner_compress_warp(const void *sums_v, size_t stride, int W, const int32_t *status,
                                                              uint8_t *out, uint32_t count) {
    __shared__ g1_jac_t part[4][32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * 4 + warp;
    if (i >= count) return;  // the whole warp
    uint32_t *o = reinterpret_cast<uint32_t *>(out + 48ull * i);
    if (status && status[i] != KZG_OK) {
        if (lane < 12) o[lane] = 0;
        return;
    }
    const int rpl = (W + 31) / 32;  // sums per lane: 8 for the 255 sums of the comb, 2 for the 64 of the latency comb
    const int lo = rpl * (int)lane, hi = min(lo + rpl, W);
    g1_jac_t acc;
    g1j_set_inf(acc);
#pragma unroll 1
    for (int j = hi - 1; j >= lo; j--) {
        g1j_dbl(acc, acc);
        if (JAC) {
            const g1_jac_t sj = static_cast<const g1_jac_t *>(sums_v)[(size_t)j * stride + i];
            g1j_add(acc, acc, sj);
        } else {
            g1_affine_t sj = horner_load(static_cast<const g1_affine_t *>(sums_v) + (size_t)j * stride + i);
            if (!g1a_is_inf(sj)) g1j_add_affine(acc, acc, sj.x, sj.y);
        }
    }
    part[warp][lane] = acc;
    __syncwarp();
#pragma unroll 1
    for (uint32_t s = 1; s < 32; s <<= 1) {
        const uint32_t role = lane & (2 * s - 1);
        if (s == 1) {
            if (role == 0) {
                g1_jac_t t = part[warp][lane + s];
#pragma unroll 1
                for (uint32_t d = 0; d < (uint32_t)rpl * s; d++) g1j_dbl(t, t);
                g1_jac_t a = part[warp][lane];
                g1j_add(t, t, a);
                part[warp][lane] = t;
            }
        } else if (role < 3) {
            // from the second round on the joining lanes are at least four apart: the two lanes after each of them share
            // its 8 s doublings (g1j_dbl_n_coop3, g1.cuh) -- 240 of the 248 doublings of the critical path
            const uint32_t mask = s == 2 ? 0x77777777u : s == 4 ? 0x07070707u : s == 8 ? 0x00070007u : 0x00000007u;
            g1_jac_t t = part[warp][(lane - role) + s];
            g1j_dbl_n_coop3(t, (uint32_t)rpl * s, mask, (int)(lane - role), (int)role);
            if (role == 0) {
                g1_jac_t a = part[warp][lane];
                g1j_add(t, t, a);
                part[warp][lane] = t;
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        g1_affine_t p;
        g1j_to_affine(p, part[warp][0]);
        uint8_t buf[48];
        g1a_compress(buf, p);
#pragma unroll
        for (int k = 0; k < 12; k++)
            o[k] = (uint32_t)buf[4 * k] | ((uint32_t)buf[4 * k + 1] << 8) | ((uint32_t)buf[4 * k + 2] << 16) | ((uint32_t)buf[4 * k + 3] << 24);
    }
}
int g1_launch_horner_compress(cudaStream_t st, const g1_affine_t *d_sums, size_t stride, int W, const int32_t *d_status,
                              uint8_t *d_out, size_t count) {
    if (count == 0) return KZG_B200_OK;
    if (count <= KZG_HORNER_WARP_MAX && W <= 256)
        k_horner_compress_warp<false><<<blocks_for(count, 4), 128, 0, st>>>(d_sums, stride, W, d_status, d_out, (uint32_t)count);
    else
        k_horner_compress<<<blocks_for(count, 64), 64, 0, st>>>(d_sums, stride, W, d_status, d_out, (uint32_t)count);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

int g1_launch_horner_compress_jac(cudaStream_t st, const g1_jac_t *d_sums, size_t stride, int W, const int32_t *d_status,
                                  uint8_t *d_out, size_t count) {
    if (count == 0) return KZG_B200_OK;
    if (count > KZG_HORNER_WARP_MAX || W > 256) return KZG_B200_BAD_ARGS;
    k_horner_compress_warp<true><<<blocks_for(count, 4), 128, 0, st>>>(d_sums, stride, W, d_status, d_out, (uint32_t)count);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ whole-batch identity check (test aid)
// The bundled setup is the public testing setup with tau = 1337, so a commitment can be
// checked without a second MSM: C == [p(tau)] G1.  y_be = p_i(tau) as 32 big-endian bytes at stride 64
// (the z || y records of the evaluation kernel).  [y]G1 runs as a plain double-and-add ladder on the
// generator -- nothing of the MSM (table, recoding, addition kernel) is on this side of the comparison.
__global__ void __launch_bounds__(64) k_tau_identity(const uint8_t *__restrict__ commitments, const uint8_t *__restrict__ zy,
                                                     uint32_t count, int32_t *__restrict__ ok) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fr_t y;
    scalar_from_be32(y, zy + 64ull * i + 32);
    g1_affine_t gen;
    {
        constexpr uint32_t gx[12] = {G1_GEN_X_MONT_LIMBS}, gy[12] = {G1_GEN_Y_MONT_LIMBS};
#pragma unroll
        for (int k = 0; k < 12; k++) { gen.x.l[k] = gx[k]; gen.y.l[k] = gy[k]; }
    }
    g1_jac_t acc;
    g1j_mul(acc, gen, y.l, 255);
    g1_affine_t a;
    g1j_to_affine(a, acc);
    uint8_t buf[48];
    g1a_compress(buf, a);
    int same = 1;
    for (int k = 0; k < 48; k++) same &= buf[k] == commitments[48ull * i + k];
    ok[i] = same;
}
int g1_launch_tau_identity(cudaStream_t st, const uint8_t *d_commitments, const uint8_t *d_zy, size_t count, int32_t *d_ok) {
    if (count == 0) return KZG_B200_OK;
    k_tau_identity<<<blocks_for(count, 64), 64, 0, st>>>(d_commitments, d_zy, (uint32_t)count, d_ok);
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ micro-benchmarks for the roofline denominators
__global__ void k_peak_imad(uint32_t *out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 3;
    uint32_t x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            asm volatile("mad.lo.u32 %0, %0, %8, %9;\n\tmad.lo.u32 %1, %1, %8, %9;\n\tmad.lo.u32 %2, %2, %8, %9;\n\t"
                         "mad.lo.u32 %3, %3, %8, %9;\n\tmad.lo.u32 %4, %4, %8, %9;\n\tmad.lo.u32 %5, %5, %8, %9;\n\t"
                         "mad.lo.u32 %6, %6, %8, %9;\n\tmad.lo.u32 %7, %7, %8, %9;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
                         : "r"(b), "r"(a));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}
// 32x32+64 multiply-adds in carry chains (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32.X), the
// form the field multiplication uses; the multiplier depends on the running value so ptxas
// cannot strength-reduce it.  32 wide MACs per inner step.
__global__ void k_peak_imad_wide(uint64_t *out, int iters) {
    uint32_t a[8], c[16];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + i;
    for (int i = 0; i < 16; i++) c[i] = i;
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
        }
    }
    uint32_t x = 0;
    for (int i = 0; i < 16; i++) x ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void __launch_bounds__(128, 4) k_peak_fpmul(fp_t *out, int iters) {
    fp_t x = fe_one<FpParams>(), y = fp_const_b();
    x.l[0] += threadIdx.x;
    y.l[1] ^= blockIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        fe_mul(x, x, y);
        fe_mul(y, y, x);
    }
    fe_add(x, x, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int g1_measure_peaks(kzg_b200_ctx *ctx, double *imad_per_s, double *imad_wide_per_s, double *fp_mul_per_s) {
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    struct EventGuard { cudaEvent_t a, b; ~EventGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{e0, e1};
    float ms = 0;
    {
        const int blocks = ctx->sms * 8, tpb = 256, iters = 4096;
        DeviceBuf d;
        CU(d.alloc((size_t)blocks * tpb * 4));
        k_peak_imad<<<blocks, tpb, 0, ctx->stream>>>(d.as<uint32_t>(), 64);
        CU(cudaEventRecord(e0, ctx->stream));
        k_peak_imad<<<blocks, tpb, 0, ctx->stream>>>(d.as<uint32_t>(), iters);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (imad_per_s) *imad_per_s = (double)blocks * tpb * iters * 64.0 / (ms * 1e-3);
    }
    {
        const int blocks = ctx->sms * 8, tpb = 256, iters = 4096;
        DeviceBuf d;
        CU(d.alloc((size_t)blocks * tpb * 8));
        k_peak_imad_wide<<<blocks, tpb, 0, ctx->stream>>>(d.as<uint64_t>(), 64);
        CU(cudaEventRecord(e0, ctx->stream));
        k_peak_imad_wide<<<blocks, tpb, 0, ctx->stream>>>(d.as<uint64_t>(), iters);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (imad_wide_per_s) *imad_wide_per_s = (double)blocks * tpb * iters * 32.0 / (ms * 1e-3);
    }
    {
        const int blocks = ctx->sms * 4, tpb = 128, iters = 2048;
        DeviceBuf d;
        CU(d.alloc((size_t)blocks * tpb * sizeof(fp_t)));
        k_peak_fpmul<<<blocks, tpb, 0, ctx->stream>>>(d.as<fp_t>(), 16);
        CU(cudaEventRecord(e0, ctx->stream));
        k_peak_fpmul<<<blocks, tpb, 0, ctx->stream>>>(d.as<fp_t>(), iters);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (fp_mul_per_s) *fp_mul_per_s = (double)blocks * tpb * iters * 2.0 / (ms * 1e-3);
    }
    ctx->launches += 6;
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ device field operations on arrays (unit tests)
// op 0: Fp mul, 1: Fp inverse (b ignored), 2: Fr mul (8 words per element, all others 12), 3: Fp add, 4: Fp sub,
// 7: lazy Fp mul, 8: lazy Fp sub -- operands and results in [0, 2p).
__global__ void k_debug_field_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out, uint64_t count) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (op == 2) {
        fr_t x, y, z;
        for (int k = 0; k < 8; k++) { x.l[k] = a[8 * i + k]; y.l[k] = b[8 * i + k]; }
        fe_mul(z, x, y);
        for (int k = 0; k < 8; k++) out[8 * i + k] = z.l[k];
        return;
    }
    fp_t x, y, z;
    for (int k = 0; k < 12; k++) { x.l[k] = a[12 * i + k]; y.l[k] = b[12 * i + k]; }
    if (op == 0) fe_mul(z, x, y);
    else if (op == 1) fp_inv(z, x);
    else if (op == 3) fe_add(z, x, y);
    else if (op == 7) fe_mul_lazy(z, x, y);
    else if (op == 8) fe_sub_lazy(z, x, y);
    else fe_sub(z, x, y);
    for (int k = 0; k < 12; k++) out[12 * i + k] = z.l[k];
}
int g1_debug_field_op(kzg_b200_ctx *ctx, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, uint64_t count) {
    if (!(op >= 0 && op <= 4) && op != 7 && op != 8) return KZG_B200_BAD_ARGS;
    const size_t w = op == 2 ? 8 : 12;
    DeviceBuf d;
    CU(d.alloc(3 * count * w * 4));
    uint32_t *p = d.as<uint32_t>();
    CU(cudaMemcpy(p, a, count * w * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p + count * w, b, count * w * 4, cudaMemcpyHostToDevice));
    k_debug_field_op<<<blocks_for(count, 128), 128, 0, ctx->stream>>>(op, p, p + count * w, p + 2 * count * w, count);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, p + 2 * count * w, count * w * 4, cudaMemcpyDeviceToHost));
    return KZG_B200_OK;
}

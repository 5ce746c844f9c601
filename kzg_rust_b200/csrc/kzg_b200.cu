// kzg_b200.cu -- kernels, context and C ABI of the B200 blob path (see include/kzg_b200.h).
//
// Host code here only sequences kernels and moves bytes; every arithmetic step of the
// path runs on the GPU.  There is deliberately no CPU fallback: if CUDA is unusable the
// entry points return KZG_B200_CUDA_ERROR.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/kzg_b200.h"
#include "blobpath.cuh"
#include "frpath.cuh"
#include "host_pairing.h"
#include "msm.cuh"
#include "fp_hybrid.cuh"

using namespace kzg;

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            if (getenv("KZG_B200_DEBUG")) fprintf(stderr, "[kzg_b200] %s -> %s (%s:%d)\n", #expr, \
                                                  cudaGetErrorString(e_), __FILE__, __LINE__);    \
            return KZG_B200_CUDA_ERROR;                                                            \
        }                                                                                          \
    } while (0)
#define RC(expr)                      \
    do {                              \
        int rc_ = (expr);             \
        if (rc_ != KZG_B200_OK) return rc_; \
    } while (0)

// ------------------------------------------------------------------ context
#define KZG_SLOTS 3
#define KZG_DYN_COUNTERS 32
struct kzg_b200_ctx {
    int device = 0;
    int n = 0;        // FIELD_ELEMENTS_PER_BLOB
    int c = 0;        // window bits
    int W = 0;        // windows
    uint32_t D = 0;   // table entries per point: 2^(c-1)
    int sms = 0;
    int max_k = 4096; // additions per thread per batch (one shared inversion per block and batch)
    int add_blocks = 3;  // resident blocks per SM of the addition kernel (KZG_B200_ADD_BLOCKS)
    int grid_mult = 1;   // grid = grid_mult x grid_blocks x SMs (KZG_B200_GRID_MULT)
    int tree_blocks = 0; // blocks per SM of the tree levels of the MSM (KZG_B200_TREE_BLOCKS; 0 = grid_blocks). 2 is 5% faster per launch but loses the overlap of the two lanes
    bool dynamic = false; // KZG_B200_DYNAMIC=1: MSM levels pull 32-addition tiles from a counter instead of equal static shares
    int dyn_min_batch = 32;  // shortest batch (tiles per inversion) at the end of a work-pulling launch (KZG_B200_DYN_MIN_BATCH)
    int grid_blocks = 3; // blocks per SM a launch asks for (KZG_B200_GRID_BLOCKS); below add_blocks leaves room for the other lane
    g1_affine_t *d_table = nullptr;
    fr_t *d_roots = nullptr;       // roots of unity, Montgomery form, bit-reversed (src/kzg.rs:764-799)
    uint8_t g2_tau[96];            // [tau]G2 = g2_values[1]
    // Work is cut into chunks of `chunk` blobs.  Two chunks are in flight at a time, each on its
    // own lane (stream + workspace), so the latency-bound end of one chunk (the small levels of the
    // addition tree, the Horner pass) runs under the big levels of the next.  Lane 0 launches on
    // `stream`, the stream callers synchronise with; `cur` is the lane of the chunk being enqueued
    // (calls on a context are serialised by `mu`).
    struct Lane {
        cudaStream_t stream = nullptr;
        int32_t *d_digits = nullptr;      // [b][j][i] as the per-blob kernels write them
        int32_t *d_digits_t = nullptr;    // [i][j][b] point-major copy the gather level reads
        unsigned int *d_counters = nullptr;  // tile counters of the work-pulling launches, one per level
        g1_affine_t *d_buf_a = nullptr, *d_buf_b = nullptr;
        fp_t *d_scratch = nullptr;
        size_t scratch_elems = 0;
        fr_t *d_poly = nullptr;           // chunk x n Montgomery evaluations (proof / verify paths)
        fr_t *d_inv = nullptr;            // chunk x n prefix products, then 1/(z - w_i)
        fr_t *d_z = nullptr;              // chunk challenges / evaluation points (canonical)
        uint8_t *d_zy = nullptr;          // chunk x 64 B: z || y big-endian
        g1_affine_t *d_pts = nullptr;     // chunk x 2 decoded commitments / proofs
        cudaEvent_t ev_done = nullptr;
    };
    Lane lanes[2];
    int nlanes = 2;                   // KZG_B200_LANES
    Lane *cur = nullptr;
    cudaEvent_t ev_start = nullptr;
    size_t chunk = 0;
    // host-call staging, KZG_SLOTS slots: chunks i+1, i+2 are uploaded on copy_stream while chunk i computes
    uint8_t *d_stage_in = nullptr;    // slots x chunk blobs
    uint8_t *d_stage_aux = nullptr;   // slots x chunk x 96 B (commitments / proofs / z)
    uint8_t *d_stage_out = nullptr;   // slots x chunk x 96 B
    int32_t *d_status = nullptr;      // slots x chunk
    cudaStream_t copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;   // small verification calls: point validation beside the challenge hash
    cudaEvent_t ev_side_fork = nullptr, ev_side_join = nullptr;
    cudaEvent_t ev_h2d[KZG_SLOTS] = {nullptr, nullptr, nullptr}, ev_free[KZG_SLOTS] = {nullptr, nullptr, nullptr};
    host_g2_prepared *tau_prepared = nullptr;  // Miller-loop lines of [tau]G2
    g1_affine_t *d_sums_all = nullptr; // window sums of a whole device-resident call, [window][blob] (grow-only)
    size_t sums_all_elems = 0;
    fr_t *d_z_all = nullptr;          // challenges of a whole device-resident proof call (grow-only)
    size_t z_all_elems = 0;
    uint8_t *d_vb = nullptr;          // phase-B buffer of batch verification (grow-only)
    size_t vb_bytes = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    // optional per-stage device timing (CUDA events on `stream`), see kzg_b200_profile_*
    bool profile = false;
    struct StageRec { int stage; cudaEvent_t a, b; };
    std::vector<StageRec> pending;
    double stage_ms[KZG_B200_NUM_STAGES] = {0};
    uint64_t stage_launches[KZG_B200_NUM_STAGES] = {0};
    std::mutex mu;
};

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// ------------------------------------------------------------------ stage timing
static void stage_begin(kzg_b200_ctx *ctx, int stage) {
    if (!ctx->profile) return;
    kzg_b200_ctx::StageRec r;
    r.stage = stage;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, ctx->cur->stream);
    ctx->pending.push_back(r);
}
static void stage_end(kzg_b200_ctx *ctx, uint64_t launches) {
    if (!ctx->profile || ctx->pending.empty()) return;
    cudaEventRecord(ctx->pending.back().b, ctx->cur->stream);
    ctx->stage_launches[ctx->pending.back().stage] += launches;
}
static void stage_collect(kzg_b200_ctx *ctx) {  // stream must be idle
    for (auto &r : ctx->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) ctx->stage_ms[r.stage] += ms;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->pending.clear();
}

// ------------------------------------------------------------------ kernels
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
// table[i*D] = decoded[bitrev(i)]   (reference bit_reversal_permutation, src/kzg.rs:717-731)
__global__ void k_place_bases(const g1_affine_t *decoded, g1_affine_t *table, uint32_t n, uint32_t D) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r = 0, v = i;
    for (uint32_t o = n; o > 1; o >>= 1) { r = (r << 1) | (v & 1); v >>= 1; }
    table[(uint64_t)i * D] = decoded[r];
}
__global__ void k_blob_digits(const uint8_t *blobs, uint64_t total, int n, int c, int W, int32_t *digits,
                              int32_t *status) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    blob_digits_thread(blobs, e, n, c, W, digits, status);
}
__global__ void k_fr_digits(const fr_t *evals, uint64_t total, int n, int c, int W, int32_t *digits) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    fr_digits_thread(evals, e, n, c, W, digits);
}
// digits[(b*W + j)*n + i] (as the per-blob producers write them, coalesced in i)
//   -> digits_t[(i*W + j)*count + b] (point-major, what the gather level reads, coalesced in b)
// grid (n/32, ceil(count/32), W), block (32, 8)
__global__ void k_transpose_digits(const int32_t *__restrict__ in, int32_t *__restrict__ out, uint32_t n, uint32_t W,
                                   uint32_t count) {
    __shared__ int32_t tile[32][33];
    const uint32_t j = blockIdx.z, i0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (uint32_t r = threadIdx.y; r < 32; r += 8) {
        uint32_t b = b0 + r, i = i0 + threadIdx.x;
        if (b < count && i < n) tile[r][threadIdx.x] = in[((uint64_t)b * W + j) * n + i];
    }
    __syncthreads();
    for (uint32_t r = threadIdx.y; r < 32; r += 8) {
        uint32_t i = i0 + r, b = b0 + threadIdx.x;
        if (b < count && i < n) out[((uint64_t)i * W + j) * count + b] = tile[threadIdx.x][r];
    }
}
// window sums of blob i (W affine points at sums[j*count + i]) -> Horner -> 48-byte compressed point
__global__ void __launch_bounds__(64) k_horner_compress(const g1_affine_t *sums, int c, int W, const int32_t *status, uint8_t *out,
                                                        uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint8_t buf[48];
    if (status && status[i] != KZG_OK) {
        for (int k = 0; k < 48; k++) buf[k] = 0;
    } else {
        g1_affine_t p;
        horner_thread(p, sums + i, count, c, W);
        g1a_compress(buf, p);
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(out + 48ull * i);
#pragma unroll
    for (int k = 0; k < 12; k++)
        o[k] = (uint32_t)buf[4 * k] | ((uint32_t)buf[4 * k + 1] << 8) | ((uint32_t)buf[4 * k + 2] << 16) | ((uint32_t)buf[4 * k + 3] << 24);
}
__global__ void k_fill_i32(int32_t *p, int32_t v, uint64_t count) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

// ---- micro-benchmarks for the roofline denominators
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

// ------------------------------------------------------------------ launch helpers
static inline unsigned blocks_for(uint64_t total, unsigned tpb) { return (unsigned)((total + tpb - 1) / tpb); }

static int ensure_scratch(kzg_b200_ctx *ctx, size_t elems) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    if (elems <= ln->scratch_elems) return KZG_B200_OK;
    if (ln->d_scratch) CU(cudaFree(ln->d_scratch));
    ln->d_scratch = nullptr;
    ln->scratch_elems = 0;
    CU(cudaMalloc(&ln->d_scratch, elems * sizeof(fp_t)));
    ln->scratch_elems = elems;
    return KZG_B200_OK;
}

// blocks_per_sm = 0: the context's default (grid_blocks)
template <class Policy>
static int launch_batch_add(kzg_b200_ctx *ctx, const Policy &pol, uint64_t total, int blocks_per_sm = 0) {
    if (total == 0) return KZG_B200_OK;
    const unsigned tpb = KZG_ADD_THREADS;
    const uint64_t t_max = (uint64_t)ctx->sms * (blocks_per_sm > 0 ? blocks_per_sm : ctx->grid_blocks) * tpb * ctx->grid_mult;
    uint64_t T;
    int k;
    if (total <= t_max) {
        T = (total + tpb - 1) / tpb * tpb;
        k = 1;
    } else {
        T = t_max;
        uint64_t need = (total + T - 1) / T;
        k = (int)std::min<uint64_t>(need, (uint64_t)ctx->max_k);
    }
    RC(ensure_scratch(ctx, (size_t)(T * k)));
    if (ctx->add_blocks >= 4)
        batch_add_kernel<Policy, 4><<<(unsigned)(T / tpb), tpb, 0, ctx->cur->stream>>>(pol, total, ctx->cur->d_scratch, k);
    else
        batch_add_kernel<Policy, 3><<<(unsigned)(T / tpb), tpb, 0, ctx->cur->stream>>>(pol, total, ctx->cur->d_scratch, k);
    ctx->launches++;
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// Work-pulling launch (batch_add_dyn_kernel): warps take 32-addition tiles from counter #slot of the lane.
template <class Policy>
static int launch_batch_add_dyn(kzg_b200_ctx *ctx, const Policy &pol, uint64_t total, int slot) {
    if (total == 0) return KZG_B200_OK;
    kzg_b200_ctx::Lane *ln = ctx->cur;
    const unsigned tpb = KZG_ADD_THREADS;
    const uint64_t ntiles = (total + 31) / 32;
    uint64_t blocks = (uint64_t)ctx->sms * ctx->grid_blocks;
    blocks = std::min<uint64_t>(blocks, (ntiles + tpb / 32 - 1) / (tpb / 32));
    const uint64_t nwarps = blocks * (tpb / 32);
    DynSchedule ds;
    ds.ntiles = (uint32_t)ntiles;
    ds.m_min = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(ntiles / (nwarps * 8), 1), (uint64_t)ctx->dyn_min_batch);
    ds.cap = (uint32_t)std::max<uint64_t>(ntiles / (2 * nwarps) + 1, ds.m_min);
    const size_t prefix_elems = (size_t)nwarps * ds.cap * 32;
    const size_t id_elems = ((size_t)nwarps * ds.cap * sizeof(uint32_t) + sizeof(fp_t) - 1) / sizeof(fp_t);
    RC(ensure_scratch(ctx, prefix_elems + id_elems));
    ds.scratch = ln->d_scratch;
    ds.tile_ids = reinterpret_cast<uint32_t *>(ln->d_scratch + prefix_elems);
    ds.counter = ln->d_counters + slot;
    if (ctx->add_blocks >= 4)
        batch_add_dyn_kernel<Policy, 4><<<(unsigned)blocks, tpb, 0, ln->stream>>>(pol, total, ds);
    else
        batch_add_dyn_kernel<Policy, 3><<<(unsigned)blocks, tpb, 0, ln->stream>>>(pol, total, ds);
    ctx->launches++;
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// the W window sums (4096 table entries each, selected by d_digits) of `count` blobs, window-major:
// (*out)[j*count + b] = S_j of blob b
static int run_msm(kzg_b200_ctx *ctx, size_t count, const g1_affine_t **out) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    const uint32_t n = (uint32_t)ctx->n, W = (uint32_t)ctx->W;
    const uint64_t R = (uint64_t)count * W;  // (window, blob) pairs = points per row of a level
    if (R * (n / 2) >= (1ull << 32)) return KZG_B200_BAD_ARGS;  // chunk sizes keep every level below 2^32 additions
    stage_begin(ctx, KZG_B200_STAGE_DIGITS);
    {
        dim3 grid((n + 31) / 32, (unsigned)((count + 31) / 32), W), block(32, 8);
        k_transpose_digits<<<grid, block, 0, ln->stream>>>(ln->d_digits, ln->d_digits_t, n, W, (uint32_t)count);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    stage_end(ctx, 1);
    const FastDiv fd = FastDiv::make((uint32_t)R);
    uint32_t rows = n / 2;
    int slot = 0;
    if (ctx->dynamic) CU(cudaMemsetAsync(ln->d_counters, 0, KZG_DYN_COUNTERS * sizeof(unsigned int), ln->stream));
    GatherPolicy gp{ctx->d_table, ln->d_digits_t, ln->d_buf_a, fd, ctx->D};
    stage_begin(ctx, KZG_B200_STAGE_MSM_GATHER);
    if (ctx->dynamic) RC(launch_batch_add_dyn(ctx, gp, R * rows, slot++));
    else RC(launch_batch_add(ctx, gp, R * rows));
    stage_end(ctx, 1);
    g1_affine_t *in = ln->d_buf_a, *o = ln->d_buf_b;
    stage_begin(ctx, KZG_B200_STAGE_MSM_TREE);
    uint64_t levels = 0;
    while (rows > 1) {
        rows /= 2;
        PairPolicy tp{in, o, fd};
        if (ctx->dynamic) RC(launch_batch_add_dyn(ctx, tp, R * rows, slot++));
        else RC(launch_batch_add(ctx, tp, R * rows, ctx->tree_blocks));
        std::swap(in, o);
        levels++;
    }
    stage_end(ctx, levels);
    *out = in;
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ workspace
static size_t per_blob_workspace(const kzg_b200_ctx *ctx) {
    size_t wn = (size_t)ctx->W * ctx->n;
    size_t lane = 2 * wn * 4 /*digits, both layouts*/ + wn / 2 * sizeof(g1_affine_t) + (wn / 4 + 1) * sizeof(g1_affine_t) +
                  2 * (size_t)ctx->n * sizeof(fr_t) /*poly, inv*/ + 64 + sizeof(fr_t) + 2 * sizeof(g1_affine_t) +
                  (wn / 2) * sizeof(fp_t) /*scratch of the gather level*/;
    return ctx->nlanes * lane + KZG_SLOTS * ((size_t)ctx->n * 32 + 96 * 2 + 4);
}
static void free_workspace(kzg_b200_ctx *ctx) {
    for (auto &ln : ctx->lanes) {
        cudaFree(ln.d_digits); cudaFree(ln.d_digits_t); cudaFree(ln.d_counters); cudaFree(ln.d_buf_a); cudaFree(ln.d_buf_b); cudaFree(ln.d_poly); cudaFree(ln.d_inv);
        cudaFree(ln.d_z); cudaFree(ln.d_zy); cudaFree(ln.d_pts); cudaFree(ln.d_scratch);
        ln.d_digits = ln.d_digits_t = nullptr; ln.d_counters = nullptr; ln.d_buf_a = ln.d_buf_b = nullptr; ln.d_poly = ln.d_inv = ln.d_z = nullptr;
        ln.d_zy = nullptr; ln.d_pts = nullptr; ln.d_scratch = nullptr; ln.scratch_elems = 0;
    }
    cudaFree(ctx->d_stage_in); cudaFree(ctx->d_stage_aux); cudaFree(ctx->d_stage_out); cudaFree(ctx->d_status);
    ctx->d_stage_in = ctx->d_stage_aux = ctx->d_stage_out = nullptr;
    ctx->d_status = nullptr;
    ctx->chunk = 0;
}
static int alloc_workspace(kzg_b200_ctx *ctx, size_t chunk) {
    free_workspace(ctx);
    size_t wn = (size_t)ctx->W * ctx->n;
    for (int l = 0; l < ctx->nlanes; l++) {
        kzg_b200_ctx::Lane &ln = ctx->lanes[l];
        CU(cudaMalloc(&ln.d_digits, chunk * wn * sizeof(int32_t)));
        CU(cudaMalloc(&ln.d_digits_t, chunk * wn * sizeof(int32_t)));
        CU(cudaMalloc(&ln.d_counters, KZG_DYN_COUNTERS * sizeof(unsigned int)));
        CU(cudaMalloc(&ln.d_buf_a, chunk * (wn / 2) * sizeof(g1_affine_t)));
        CU(cudaMalloc(&ln.d_buf_b, chunk * (wn / 4 + 1) * sizeof(g1_affine_t)));
        CU(cudaMalloc(&ln.d_poly, chunk * (size_t)ctx->n * sizeof(fr_t)));
        CU(cudaMalloc(&ln.d_inv, chunk * (size_t)ctx->n * sizeof(fr_t)));
        CU(cudaMalloc(&ln.d_z, chunk * sizeof(fr_t)));
        CU(cudaMalloc(&ln.d_zy, chunk * 64));
        CU(cudaMalloc(&ln.d_pts, chunk * 2 * sizeof(g1_affine_t)));
    }
    CU(cudaMalloc(&ctx->d_stage_in, KZG_SLOTS * chunk * (size_t)ctx->n * 32));
    CU(cudaMalloc(&ctx->d_stage_aux, KZG_SLOTS * chunk * 96));
    CU(cudaMalloc(&ctx->d_stage_out, KZG_SLOTS * chunk * 96));
    CU(cudaMalloc(&ctx->d_status, KZG_SLOTS * chunk * sizeof(int32_t)));
    ctx->chunk = chunk;
    return KZG_B200_OK;
}

// Chunks of one call alternate between the lanes.  lanes_begin makes the extra lane wait for what is
// already queued on the caller-visible stream, lanes_end makes that stream wait for the extra lane.
static int lanes_in_use(const kzg_b200_ctx *ctx) { return ctx->profile ? 1 : ctx->nlanes; }  // profiling wants unoverlapped stages
static int lanes_begin(kzg_b200_ctx *ctx) {
    ctx->cur = &ctx->lanes[0];
    if (lanes_in_use(ctx) > 1) {
        CU(cudaEventRecord(ctx->ev_start, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->lanes[1].stream, ctx->ev_start, 0));
    }
    return KZG_B200_OK;
}
static void lane_select(kzg_b200_ctx *ctx, size_t chunk_index) { ctx->cur = &ctx->lanes[chunk_index % lanes_in_use(ctx)]; }
static int lanes_end(kzg_b200_ctx *ctx) {
    if (lanes_in_use(ctx) > 1) {
        CU(cudaEventRecord(ctx->lanes[1].ev_done, ctx->lanes[1].stream));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->lanes[1].ev_done, 0));
    }
    ctx->cur = &ctx->lanes[0];
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ context creation
static int build_table(kzg_b200_ctx *ctx, const uint8_t *g1_bytes) {
    const int n = ctx->n;
    ctx->cur = &ctx->lanes[0];
    uint8_t *d_bytes = nullptr;
    g1_affine_t *d_dec = nullptr;
    int32_t *d_st = nullptr;
    CU(cudaMalloc(&d_bytes, (size_t)n * 48));
    CU(cudaMalloc(&d_dec, (size_t)n * sizeof(g1_affine_t)));
    CU(cudaMalloc(&d_st, (size_t)n * sizeof(int32_t)));
    CU(cudaMemcpyAsync(d_bytes, g1_bytes, (size_t)n * 48, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_st, 0, (size_t)n * sizeof(int32_t), ctx->stream));
    // reference load_trusted_setup does not subgroup-check the G1 points (src/kzg.rs:859-872)
    k_decode_g1<<<blocks_for(n, 64), 64, 0, ctx->stream>>>(d_bytes, d_dec, d_st, n, 0, (uint32_t)n);
    ctx->launches++;
    std::vector<int32_t> st(n);
    CU(cudaMemcpyAsync(st.data(), d_st, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++)
        if (st[i] != 0) { cudaFree(d_bytes); cudaFree(d_dec); cudaFree(d_st); return KZG_B200_BAD_ARGS; }
    k_place_bases<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(d_dec, ctx->d_table, n, ctx->D);
    ctx->launches += 1;
    CU(cudaGetLastError());
    for (int L = 0; L + 1 < ctx->c; L++) {
        TableLevelPolicy pol{ctx->d_table, ctx->D, (uint32_t)L};
        RC(launch_batch_add(ctx, pol, (uint64_t)n << L));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bytes); cudaFree(d_dec); cudaFree(d_st);
    return KZG_B200_OK;
}

extern "C" int kzg_b200_ctx_create(const uint8_t *g1_lagrange, size_t n1, const uint8_t *g2_monomial, size_t n2,
                                   int device, int window_bits, kzg_b200_ctx **out) {
    if (!g1_lagrange || !g2_monomial || !out) return KZG_B200_BAD_ARGS;
    // reference src/kzg.rs:843-847 (n1 fixed per preset; both presets accepted here)
    if ((n1 != 4096 && n1 != 4) || n2 != KZG_B200_NUM_G2_POINTS) return KZG_B200_BAD_ARGS;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return KZG_B200_CUDA_ERROR;
    CU(cudaSetDevice(device));
    // G2 side of the setup on the host: decode all 65 points (reference src/kzg.rs:874-887) and
    // run the Lagrange-form sanity pairing (src/kzg.rs:802-830) -- it needs g1[0], g1[1] only.
    {
        int rc = host_check_setup(g1_lagrange, g2_monomial, n2);
        if (rc != KZG_B200_OK) return rc;
    }
    kzg_b200_ctx *ctx = new kzg_b200_ctx();
    ctx->device = device;
    ctx->n = (int)n1;
    memcpy(ctx->g2_tau, g2_monomial + 96, 96);
    ctx->tau_prepared = host_g2_prepare(ctx->g2_tau);
    if (!ctx->tau_prepared) { delete ctx; return KZG_B200_BAD_ARGS; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return KZG_B200_CUDA_ERROR; }
    ctx->sms = prop.multiProcessorCount;
    ctx->max_k = env_int("KZG_B200_BATCH_K", 4096);
    ctx->add_blocks = env_int("KZG_B200_ADD_BLOCKS", 3) >= 4 ? 4 : 3;
    ctx->grid_mult = std::max(1, env_int("KZG_B200_GRID_MULT", 1));
    ctx->dynamic = env_int("KZG_B200_DYNAMIC", 0) != 0;
    ctx->tree_blocks = std::min(std::max(0, env_int("KZG_B200_TREE_BLOCKS", 0)), 4);
    ctx->dyn_min_batch = std::max(1, env_int("KZG_B200_DYN_MIN_BATCH", 32));
    ctx->grid_blocks = std::max(1, env_int("KZG_B200_GRID_BLOCKS", ctx->add_blocks));
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return KZG_B200_CUDA_ERROR; }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { kzg_b200_ctx_destroy(ctx); return KZG_B200_CUDA_ERROR; }
    if (cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_side_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_side_join, cudaEventDisableTiming) != cudaSuccess) { kzg_b200_ctx_destroy(ctx); return KZG_B200_CUDA_ERROR; }
    ctx->nlanes = env_int("KZG_B200_LANES", 2) >= 2 ? 2 : 1;
    ctx->lanes[0].stream = ctx->stream;
    if (cudaStreamCreateWithFlags(&ctx->lanes[1].stream, cudaStreamNonBlocking) != cudaSuccess) { kzg_b200_ctx_destroy(ctx); return KZG_B200_CUDA_ERROR; }
    ctx->cur = &ctx->lanes[0];
    bool ev_ok = cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2; i++) ev_ok = ev_ok && cudaEventCreateWithFlags(&ctx->lanes[i].ev_done, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < KZG_SLOTS; i++)
        ev_ok = ev_ok && cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ev_ok) { kzg_b200_ctx_destroy(ctx); return KZG_B200_CUDA_ERROR; }
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    int c = window_bits > 0 ? window_bits : env_int("KZG_B200_WINDOW_BITS", 0);
    if (c <= 0) {
        // largest window whose table leaves room for a 4096-blob workspace and 16 GiB of caller data
        // (19 on an empty 180 GB B200: 14 windows, 103 GB of table)
        for (c = 19; c > 2; c--) {
            size_t tbl = n1 * ((size_t)1 << (c - 1)) * sizeof(g1_affine_t);
            if (tbl + ((size_t)44 << 30) <= free_b) break;
        }
    }
    if (c < 2 || c > 20) { kzg_b200_ctx_destroy(ctx); return KZG_B200_BAD_ARGS; }
    ctx->c = c;
    ctx->W = msm_num_windows(c);
    ctx->D = 1u << (c - 1);
    size_t tbl_bytes = n1 * (size_t)ctx->D * sizeof(g1_affine_t);
    if (cudaMalloc(&ctx->d_table, tbl_bytes) != cudaSuccess) { kzg_b200_ctx_destroy(ctx); return KZG_B200_CUDA_ERROR; }
    int rc = build_table(ctx, g1_lagrange);
    if (rc != KZG_B200_OK) { kzg_b200_ctx_destroy(ctx); return rc; }
    rc = fr_setup_roots(ctx->n, &ctx->d_roots, ctx->stream);
    if (rc != KZG_B200_OK) { kzg_b200_ctx_destroy(ctx); return rc; }
    // workspace: as many blobs per chunk as memory allows, capped (KZG_B200_CHUNK overrides)
    cudaMemGetInfo(&free_b, &total_b);
    size_t cap = (size_t)env_int("KZG_B200_CHUNK", 4096);
    size_t reserve = (size_t)12 << 30;  // leave room for the caller's device-resident blobs
    size_t usable = free_b > reserve + ((size_t)2 << 30) ? free_b - reserve : free_b / 2;
    size_t chunk = std::max<size_t>(1, std::min(cap, usable / per_blob_workspace(ctx)));
    rc = alloc_workspace(ctx, chunk);
    if (rc != KZG_B200_OK) { kzg_b200_ctx_destroy(ctx); return rc; }
    *out = ctx;
    return KZG_B200_OK;
}

static int hex_nibble(int ch) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
    if (ch >= 'A' && ch <= 'F') return ch - 'A' + 10;
    return -1;
}
// reference load_trusted_setup_file, src/kzg.rs:906-979
extern "C" int kzg_b200_ctx_create_from_file(const char *path, int device, int window_bits, kzg_b200_ctx **out) {
    FILE *f = fopen(path, "r");
    if (!f) return KZG_B200_INVALID_TRUSTED_SETUP;
    unsigned long n1 = 0, n2 = 0;
    if (fscanf(f, "%lu", &n1) != 1 || fscanf(f, "%lu", &n2) != 1 || (n1 != 4096 && n1 != 4) || n2 != 65) {
        fclose(f);
        return KZG_B200_INVALID_TRUSTED_SETUP;
    }
    std::vector<uint8_t> g1(n1 * 48), g2(n2 * 96);
    char tok[512];
    int rc = KZG_B200_OK;
    for (size_t i = 0; i < n1 + n2 && rc == KZG_B200_OK; i++) {
        size_t want = i < n1 ? 48 : 96;
        uint8_t *dst = i < n1 ? g1.data() + 48 * i : g2.data() + 96 * (i - n1);
        if (fscanf(f, "%511s", tok) != 1) { rc = KZG_B200_INVALID_TRUSTED_SETUP; break; }
        const char *h = tok;
        if (h[0] == '0' && h[1] == 'x') h += 2;
        if (strlen(h) != 2 * want) { rc = KZG_B200_INVALID_HEX_FORMAT; break; }
        for (size_t k = 0; k < want; k++) {
            int hi = hex_nibble(h[2 * k]), lo = hex_nibble(h[2 * k + 1]);
            if (hi < 0 || lo < 0) { rc = KZG_B200_INVALID_HEX_FORMAT; break; }
            dst[k] = (uint8_t)(hi << 4 | lo);
        }
    }
    fclose(f);
    if (rc != KZG_B200_OK) return rc;
    return kzg_b200_ctx_create(g1.data(), n1, g2.data(), n2, device, window_bits, out);
}

extern "C" void kzg_b200_ctx_destroy(kzg_b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->lanes[1].stream) cudaStreamSynchronize(ctx->lanes[1].stream);
    free_workspace(ctx);
    cudaFree(ctx->d_vb);
    cudaFree(ctx->d_z_all);
    cudaFree(ctx->d_sums_all);
    cudaFree(ctx->d_table);
    cudaFree(ctx->d_roots);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->ev_side_fork) cudaEventDestroy(ctx->ev_side_fork);
    if (ctx->ev_side_join) cudaEventDestroy(ctx->ev_side_join);
    if (ctx->lanes[1].stream) cudaStreamDestroy(ctx->lanes[1].stream);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    for (int i = 0; i < 2; i++)
        if (ctx->lanes[i].ev_done) cudaEventDestroy(ctx->lanes[i].ev_done);
    for (int i = 0; i < KZG_SLOTS; i++) {
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
    }
    stage_collect(ctx);
    host_g2_prepared_free(ctx->tau_prepared);
    delete ctx;
}
extern "C" size_t kzg_b200_field_elements_per_blob(const kzg_b200_ctx *ctx) { return ctx ? (size_t)ctx->n : 0; }
extern "C" int kzg_b200_window_bits(const kzg_b200_ctx *ctx) { return ctx ? ctx->c : 0; }
extern "C" void *kzg_b200_stream(kzg_b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" uint64_t kzg_b200_launch_count(const kzg_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int kzg_b200_synchronize(kzg_b200_ctx *ctx) {
    if (!ctx) return KZG_B200_BAD_ARGS;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    std::lock_guard<std::mutex> lock(ctx->mu);
    stage_collect(ctx);
    return KZG_B200_OK;
}
extern "C" int kzg_b200_profile_enable(kzg_b200_ctx *ctx, int on) {
    if (!ctx) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
    ctx->profile = on != 0;
    for (int i = 0; i < KZG_B200_NUM_STAGES; i++) { ctx->stage_ms[i] = 0; ctx->stage_launches[i] = 0; }
    return KZG_B200_OK;
}
extern "C" int kzg_b200_profile_read(kzg_b200_ctx *ctx, double *ms_out, uint64_t *launches_out) {
    if (!ctx) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
    for (int i = 0; i < KZG_B200_NUM_STAGES; i++) {
        if (ms_out) ms_out[i] = ctx->stage_ms[i];
        if (launches_out) launches_out[i] = ctx->stage_launches[i];
    }
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ blob_to_kzg_commitment
// one chunk, everything on the device: blobs -> digits -> MSM -> 48-byte commitments
// The Horner + compression pass is one thread per blob and ~1900 dependent products long: for a 4096-blob
// chunk it is 2.1 ms of latency on a mostly idle GPU (3.7 % of a step).  Device-resident calls that span
// several chunks therefore park the window sums of every chunk in one [window][blob] array of the whole call
// (88 MB for 65,536 blobs) and run the pass ONCE at the end, where 65,536 threads fill the machine.
struct DeferredCompress {
    g1_affine_t *sums = nullptr;  // nullptr: compress per chunk
    size_t n = 0;                 // blobs of the whole call = row pitch of `sums`
};
static int deferred_begin(kzg_b200_ctx *ctx, size_t n, DeferredCompress *dc) {
    dc->sums = nullptr;
    dc->n = n;
    if (n <= ctx->chunk) return KZG_B200_OK;
    const size_t elems = n * (size_t)ctx->W;
    if (elems > ctx->sums_all_elems) {
        if (ctx->d_sums_all) CU(cudaFree(ctx->d_sums_all));
        ctx->d_sums_all = nullptr;
        ctx->sums_all_elems = 0;
        CU(cudaMalloc(&ctx->d_sums_all, elems * sizeof(g1_affine_t)));
        ctx->sums_all_elems = elems;
    }
    dc->sums = ctx->d_sums_all;
    return KZG_B200_OK;
}
// after run_msm of one chunk (its sums are [window][count] in the lane's buffer): compress now, or park them
static int compress_or_park(kzg_b200_ctx *ctx, const g1_affine_t *res, size_t off, size_t count, const int32_t *d_status,
                            uint8_t *d_out, const DeferredCompress *dc) {
    cudaStream_t st = ctx->cur->stream;
    if (dc && dc->sums) {
        CU(cudaMemcpy2DAsync(dc->sums + off, dc->n * sizeof(g1_affine_t), res, count * sizeof(g1_affine_t),
                             count * sizeof(g1_affine_t), (size_t)ctx->W, cudaMemcpyDeviceToDevice, st));
        return KZG_B200_OK;
    }
    stage_begin(ctx, KZG_B200_STAGE_COMPRESS);
    k_horner_compress<<<blocks_for(count, 64), 64, 0, st>>>(res, ctx->c, ctx->W, d_status, d_out, (uint32_t)count);
    stage_end(ctx, 1);
    ctx->launches++;
    CU(cudaGetLastError());
    return KZG_B200_OK;
}
// once per call, on the caller-visible stream after the lanes have joined
static int deferred_finish(kzg_b200_ctx *ctx, const DeferredCompress *dc, const int32_t *d_status, uint8_t *d_out) {
    if (!dc->sums) return KZG_B200_OK;
    ctx->cur = &ctx->lanes[0];
    stage_begin(ctx, KZG_B200_STAGE_COMPRESS);
    k_horner_compress<<<blocks_for(dc->n, 64), 64, 0, ctx->stream>>>(dc->sums, ctx->c, ctx->W, d_status, d_out, (uint32_t)dc->n);
    stage_end(ctx, 1);
    ctx->launches++;
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

// d_out / d_status point at this chunk's slice; `off` is its first blob within the call (deferred compression)
static int commit_chunk(kzg_b200_ctx *ctx, const uint8_t *d_blobs, size_t count, uint8_t *d_out, int32_t *d_status,
                        size_t off = 0, const DeferredCompress *dc = nullptr) {
    const uint64_t elems = (uint64_t)count * ctx->n;
    cudaStream_t st = ctx->cur->stream;
    CU(cudaMemsetAsync(d_status, 0, count * sizeof(int32_t), st));
    stage_begin(ctx, KZG_B200_STAGE_DIGITS);
    k_blob_digits<<<blocks_for(elems, 256), 256, 0, st>>>(d_blobs, elems, ctx->n, ctx->c, ctx->W, ctx->cur->d_digits, d_status);
    stage_end(ctx, 1);
    ctx->launches++;
    const g1_affine_t *res = nullptr;
    RC(run_msm(ctx, count, &res));
    return compress_or_park(ctx, res, off, count, d_status, d_out, dc);
}

extern "C" int kzg_b200_blob_to_kzg_commitment_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, size_t n, uint8_t *d_out,
                                                      int32_t *d_status) {
    if (!ctx || (n && (!d_blobs || !d_out || !d_status))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t bpb = (size_t)ctx->n * 32;
    DeferredCompress dc;
    RC(deferred_begin(ctx, n, &dc));
    RC(lanes_begin(ctx));
    size_t i = 0;
    for (size_t off = 0; off < n; off += ctx->chunk, i++) {
        size_t cnt = std::min(ctx->chunk, n - off);
        lane_select(ctx, i);
        RC(commit_chunk(ctx, d_blobs + off * bpb, cnt, d_out + off * 48, d_status + off, off, &dc));
    }
    RC(lanes_end(ctx));
    return deferred_finish(ctx, &dc, d_status, d_out);
}

// Host-buffer calls run the chunks through KZG_SLOTS staging slots: the uploads of the next chunks
// (on copy_stream) overlap the kernels of the current ones (one chunk per lane in flight).
// upload(slot, off, cnt) enqueues the H2D copies of one chunk on copy_stream; run(slot, off, cnt)
// enqueues kernels + D2H on the current lane's stream.
template <class Upload, class Run>
static int staged_chunks(kzg_b200_ctx *ctx, size_t n, Upload upload, Run run) {
    const size_t nchunks = (n + ctx->chunk - 1) / ctx->chunk;
    const size_t ahead = KZG_SLOTS - 1;
    auto enqueue_upload = [&](size_t i) -> int {
        int slot = (int)(i % KZG_SLOTS);
        size_t off = i * ctx->chunk, cnt = std::min(ctx->chunk, n - off);
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[slot], 0));  // the chunk that used this slot is done
        RC(upload(slot, off, cnt));
        CU(cudaEventRecord(ctx->ev_h2d[slot], ctx->copy_stream));
        return KZG_B200_OK;
    };
    RC(lanes_begin(ctx));
    // the slots may still be in use by an earlier asynchronous call on `stream`
    for (int sl = 0; sl < KZG_SLOTS; sl++) CU(cudaEventRecord(ctx->ev_free[sl], ctx->stream));
    for (size_t i = 0; i < std::min(ahead, nchunks); i++) RC(enqueue_upload(i));
    for (size_t i = 0; i < nchunks; i++) {
        int slot = (int)(i % KZG_SLOTS);
        size_t off = i * ctx->chunk, cnt = std::min(ctx->chunk, n - off);
        if (i + ahead < nchunks) RC(enqueue_upload(i + ahead));
        lane_select(ctx, i);
        CU(cudaStreamWaitEvent(ctx->cur->stream, ctx->ev_h2d[slot], 0));
        RC(run(slot, off, cnt));
        CU(cudaEventRecord(ctx->ev_free[slot], ctx->cur->stream));
    }
    RC(lanes_end(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
    return KZG_B200_OK;
}

extern "C" int kzg_b200_blob_to_kzg_commitment_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out,
                                                     int32_t *status) {
    if (!ctx || (n && (!blobs || !out || !status))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t bpb = (size_t)ctx->n * 32, ch = ctx->chunk;
    return staged_chunks(
        ctx, n,
        [&](int slot, size_t off, size_t cnt) -> int {
            CU(cudaMemcpyAsync(ctx->d_stage_in + slot * ch * bpb, blobs + off * bpb, cnt * bpb, cudaMemcpyHostToDevice, ctx->copy_stream));
            return KZG_B200_OK;
        },
        [&](int slot, size_t off, size_t cnt) -> int {
            uint8_t *d_out = ctx->d_stage_out + slot * ch * 96;
            int32_t *d_st = ctx->d_status + slot * ch;
            RC(commit_chunk(ctx, ctx->d_stage_in + slot * ch * bpb, cnt, d_out, d_st));
            CU(cudaMemcpyAsync(out + off * 48, d_out, cnt * 48, cudaMemcpyDeviceToHost, ctx->cur->stream));
            CU(cudaMemcpyAsync(status + off, d_st, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->cur->stream));
            return KZG_B200_OK;
        });
}

// ------------------------------------------------------------------ roofline micro-benchmarks
extern "C" int kzg_b200_measure_peaks(kzg_b200_ctx *ctx, double *imad_per_s, double *imad_wide_per_s,
                                      double *fp_mul_per_s) {
    if (!ctx) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float ms = 0;
    {
        const int blocks = ctx->sms * 8, tpb = 256, iters = 4096;
        uint32_t *d = nullptr;
        CU(cudaMalloc(&d, (size_t)blocks * tpb * 4));
        k_peak_imad<<<blocks, tpb, 0, ctx->stream>>>(d, 64);
        CU(cudaEventRecord(e0, ctx->stream));
        k_peak_imad<<<blocks, tpb, 0, ctx->stream>>>(d, iters);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (imad_per_s) *imad_per_s = (double)blocks * tpb * iters * 64.0 / (ms * 1e-3);
        cudaFree(d);
    }
    {
        const int blocks = ctx->sms * 8, tpb = 256, iters = 4096;
        uint64_t *d = nullptr;
        CU(cudaMalloc(&d, (size_t)blocks * tpb * 8));
        k_peak_imad_wide<<<blocks, tpb, 0, ctx->stream>>>(d, 64);
        CU(cudaEventRecord(e0, ctx->stream));
        k_peak_imad_wide<<<blocks, tpb, 0, ctx->stream>>>(d, iters);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (imad_wide_per_s) *imad_wide_per_s = (double)blocks * tpb * iters * 32.0 / (ms * 1e-3);
        cudaFree(d);
    }
    {
        const int blocks = ctx->sms * 4, tpb = 128, iters = 2048;
        fp_t *d = nullptr;
        CU(cudaMalloc(&d, (size_t)blocks * tpb * sizeof(fp_t)));
        k_peak_fpmul<<<blocks, tpb, 0, ctx->stream>>>(d, 16);
        CU(cudaEventRecord(e0, ctx->stream));
        k_peak_fpmul<<<blocks, tpb, 0, ctx->stream>>>(d, iters);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (fp_mul_per_s) *fp_mul_per_s = (double)blocks * tpb * iters * 2.0 / (ms * 1e-3);
        cudaFree(d);
    }
    ctx->launches += 6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return KZG_B200_OK;
}

// debugging aid (not declared in the public header): raw table entries, affine Montgomery limbs
#ifdef KZG_TRACE
extern "C" int kzg_b200_debug_set_trace(void *d_buf) {
    unsigned long long *p = (unsigned long long *)d_buf;
    return cudaMemcpyToSymbol(kzg::g_kzg_trace, &p, sizeof(p)) == cudaSuccess ? 0 : KZG_B200_CUDA_ERROR;
}
#endif
extern "C" int kzg_b200_debug_table(kzg_b200_ctx *ctx, uint64_t first, uint64_t count, void *out) {
    if (!ctx) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpy(out, ctx->d_table + first, count * sizeof(g1_affine_t), cudaMemcpyDeviceToHost));
    return KZG_B200_OK;
}

// debugging / unit-test aid: device field operations on arrays (op 0: Fp mul, 1: Fp inverse (b ignored),
// 2: Fr mul, 3: Fp add, 4: Fp sub, 5: two-pipe Fp mul (fp_hybrid.cuh), 6: two-pipe Fp square (b ignored),
// 7: lazy Fp mul, 8: lazy Fp sub -- operands and results in [0, 2p)).  Operands are raw limbs (12 or 8 words each).
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
    else if (op == 5) fp_mul_hybrid(z, x, y);
    else if (op == 6) fp_sqr_hybrid(z, x);
    else if (op == 7) fe_mul_lazy(z, x, y);
    else if (op == 8) fe_sub_lazy(z, x, y);
    else fe_sub(z, x, y);
    for (int k = 0; k < 12; k++) out[12 * i + k] = z.l[k];
}
extern "C" int kzg_b200_debug_field_op(kzg_b200_ctx *ctx, int op, const uint32_t *a, const uint32_t *b, uint32_t *out,
                                       uint64_t count) {
    if (!ctx) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    size_t w = op == 2 ? 8 : 12;
    uint32_t *d = nullptr;
    CU(cudaMalloc(&d, 3 * count * w * 4));
    CU(cudaMemcpy(d, a, count * w * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + count * w, b, count * w * 4, cudaMemcpyHostToDevice));
    k_debug_field_op<<<blocks_for(count, 128), 128, 0, ctx->stream>>>(op, d, d + count * w, d + 2 * count * w, count);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, d + 2 * count * w, count * w * 4, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return KZG_B200_OK;
}

#include "proof_verify.inl"

// msm.cu -- host side of the MSM (kernels and policies: msm.cuh; recoding: blobpath.cuh).
//
//   msm_build_table          setup bytes -> bases (bit-reversal permuted) -> the comb table, on the device
//   msm_digits_from_*        scalars -> sign words -> comb digits [group][bit position][blob]
//   msm_run                  gather level + tree levels -> the 255 sums S_j of every blob
#include <algorithm>

#include "internal.h"
#include "msm.cuh"

using namespace kzg;

// ------------------------------------------------------------------ kernels
// bases[i] = decoded[bitrev(i)] for i < n, infinity for the padding (reference bit_reversal_permutation, src/kzg.rs:717-731)
__global__ void k_place_bases(const g1_affine_t *decoded, g1_affine_t *bases, uint32_t n, uint32_t n_pad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    g1_affine_t p;
    if (i < n) {
        uint32_t r = 0, v = i;
        for (uint32_t o = n; o > 1; o >>= 1) { r = (r << 1) | (v & 1); v >>= 1; }
        p = decoded[r];
    } else {
        g1a_set_inf(p);
    }
    bases[i] = p;
}
// T[q][0] = P_{q,0}
__global__ void k_seed_table(const g1_affine_t *bases, g1_affine_t *table, uint32_t G, uint32_t g) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= G) return;
    table[(uint64_t)q << (g - 1)] = bases[(uint64_t)q * g];
}
__global__ void k_blob_sign_words(const uint8_t *blobs, uint64_t total, int n, int n_pad, uint32_t *sign_words, int32_t *status) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    blob_sign_words_thread(blobs, e, n, n_pad, sign_words, status);
}
__global__ void k_fr_sign_words(const fr_t *scalars, uint64_t total, int n, int n_pad, uint32_t *sign_words) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    fr_sign_words_thread(scalars, e, n, n_pad, sign_words);
}
// thread = (blob b, group q), b fastest: every store of the 255 digits is coalesced over the blobs
__global__ void __launch_bounds__(128) k_comb_index(const uint32_t *__restrict__ sign_words, uint32_t count, uint32_t G, int g,
                                                     int n_pad, uint32_t *__restrict__ digits) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)count * G) return;
    uint32_t q = (uint32_t)(t / count), b = (uint32_t)(t - (uint64_t)q * count);
    comb_index_thread(sign_words, b, q, g, n_pad, count, digits);
}
// the odd group out at the gather level: out[r] = +-T[q][idx(q, r)] as a lazy residue pair
__global__ void k_gather_copy(const g1_affine_t *__restrict__ table, const uint32_t *__restrict__ digits, uint32_t q, uint64_t E,
                              uint64_t R, g1_affine_t *__restrict__ out) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const uint32_t d = digits[(uint64_t)q * R + r];
    g1_affine_t p;
    const g1_affine_t *src = table + (uint64_t)q * E + (d & 0x7fffffffu);
    ld_fp(p.x, &src->x);
    ld_fp(p.y, &src->y);
    if ((d >> 31) && !g1a_is_inf(p)) fe_neg_lazy(p.y, p.y);  // y != 0 on this curve
    st_fp(&out[r].x, p.x);
    st_fp(&out[r].y, p.y);
}

// Small batches: one WARP per sum S_j of a blob (R = count x 255 warps).  The tree of batched affine additions is eight
// launches, each around one shared inversion (~80 us on one thread): latency that a batch of a few blobs cannot hide.
// Here lane l adds the entries of groups l, l + 32, ... in Jacobian coordinates (complete mixed additions, no inversion;
// the next entry is loaded under the current addition), five shuffle rounds join the 32 partial sums, and the Horner pass
// takes the Jacobian sums as they are: ~0.1 ms for the 255 sums of a blob where gather + tree levels take 0.85 ms.
// More multiplications per addition (11 - 16 instead of 6): for batches that do not fill the GPU only.
__global__ void __launch_bounds__(128) k_comb_rows_warp(const g1_affine_t *__restrict__ table, const uint32_t *__restrict__ digits,
                                                         uint32_t G, uint64_t E, uint32_t R, g1_jac_t *__restrict__ out) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * 4 + warp;
    if (r >= R) return;  // the whole warp
    auto load = [&](uint32_t q, g1_affine_t &p) {
        const uint32_t d = digits[(uint64_t)q * R + r];
        const g1_affine_t *src = table + (uint64_t)q * E + (d & 0x7fffffffu);
        ld_fp(p.x, &src->x);
        ld_fp(p.y, &src->y);
        if ((d >> 31) && !g1a_is_inf(p)) fe_neg(p.y, p.y);  // table entries are canonical
    };
    g1_jac_t acc;
    g1j_set_inf(acc);
    g1_affine_t cur;
    g1a_set_inf(cur);
    if (lane < G) load(lane, cur);
#pragma unroll 1
    for (uint32_t q = lane; q < G; q += 32) {
        g1_affine_t nxt;
        g1a_set_inf(nxt);
        if (q + 32 < G) load(q + 32, nxt);
        if (!g1a_is_inf(cur)) g1j_add_affine(acc, acc, cur.x, cur.y);
        cur = nxt;
    }
#pragma unroll 1
    for (int s = 16; s > 0; s >>= 1) {
        g1_jac_t o;
        o.x = fp_shfl(0xffffffffu, acc.x, (int)lane + s);
        o.y = fp_shfl(0xffffffffu, acc.y, (int)lane + s);
        o.z = fp_shfl(0xffffffffu, acc.z, (int)lane + s);
        if ((int)lane < s) g1j_add(acc, acc, o);
    }
    if (lane == 0) out[r] = acc;
}

// Mid-size batches: the last levels of the addition tree (12 -> 6 -> 3 -> 2 -> 1 rows) are four launches of few additions, each
// around one shared inversion (~0.15 ms apiece whatever the size).  For chunks of up to KZG_JAC_TAIL_MAX blobs one thread per
// sum adds the remaining rows in Jacobian coordinates instead (11 mixed additions, no inversion) and the warp Horner pass
// takes the Jacobian sums.  in[q*R + r], q < rows: affine lazy residues.
__global__ void __launch_bounds__(128) k_tail_rows_jac(const g1_affine_t *__restrict__ in, uint32_t rows, uint64_t R, g1_jac_t *__restrict__ out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    g1_jac_t acc;
    g1j_set_inf(acc);
    g1_affine_t cur;
    ld_fp(cur.x, &in[r].x);
    ld_fp(cur.y, &in[r].y);
#pragma unroll 1
    for (uint32_t q = 0; q < rows; q++) {
        g1_affine_t nxt;
        g1a_set_inf(nxt);
        if (q + 1 < rows) {
            ld_fp(nxt.x, &in[(uint64_t)(q + 1) * R + r].x);
            ld_fp(nxt.y, &in[(uint64_t)(q + 1) * R + r].y);
        }
        if (!g1a_is_inf(cur)) {
            fe_canonical(cur.x);
            fe_canonical(cur.y);
            g1j_add_affine(acc, acc, cur.x, cur.y);
        }
        cur = nxt;
    }
    out[r] = acc;
}

// ---- the latency comb (internal.h): virtual points, digits, one block per sum
// out[t*n + i] = 2^(64 t) bases[i], affine canonical
__global__ void k_latency_bases(const g1_affine_t *__restrict__ bases, uint32_t n, g1_affine_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const g1_affine_t p = bases[i];
    out[i] = p;
    g1_jac_t q;
    g1j_from_affine(q, p);
#pragma unroll 1
    for (uint32_t t = 1; t < KZG_LAT_T; t++) {
#pragma unroll 1
        for (int d = 0; d < 256 / KZG_LAT_T; d++) g1j_dbl(q, q);
        g1_affine_t a;
        g1j_to_affine(a, q);
        out[(uint64_t)t * n + i] = a;
    }
}
// Thread = (blob b, group q of 16 virtual points 2^(64 t) G_i, i0 <= i < i0 + 16): the signs of virtual point (t, i) at
// row j are bits 64 t + j of the 256-digit form of scalar i's sign words -- s = sum_{j < 256} e_j 2^j with e_255 = +1,
// which differs from the 255-digit form (blobpath.cuh) in the top two bits only: e_254 flips and e_255 = the old e_254.
//   digits[(q*64 + j)*count + b]
__global__ void __launch_bounds__(128) k_comb_index_latency(const uint32_t *__restrict__ sign_words, uint32_t count, uint32_t n,
                                                             int n_pad, uint32_t *__restrict__ digits) {
    const uint32_t groups_per_t = n / KZG_LAT_G, G = KZG_LAT_T * groups_per_t;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t)count * G) return;
    const uint32_t q = (uint32_t)(tid / count), b = (uint32_t)(tid - (uint64_t)q * count);
    const uint32_t t = q / groups_per_t, i0 = (q - t * groups_per_t) * KZG_LAT_G;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        const int w = 2 * (int)t + h;
        const uint32_t *row = sign_words + ((uint64_t)b * 8 + w) * (uint64_t)n_pad + i0;
        uint32_t a[32];
#pragma unroll
        for (int k = 0; k < 32; k++) {
            uint32_t v = k < KZG_LAT_G ? row[k] : 0u;
            if (w == 7) v = (v & 0x3fffffffu) | ((~v) & 0x40000000u) | ((v & 0x40000000u) << 1);
            a[k] = v;
        }
        transpose32(a);  // a[tt] bit k = sign of point k at bit position 32 w + tt
#pragma unroll
        for (int tt = 0; tt < 32; tt++) {
            const int j = 32 * h + tt;
            digits[((uint64_t)q * KZG_LAT_ROWS + j) * count + b] = comb_digit(a[tt], KZG_LAT_G);
        }
    }
}
// One BLOCK of four warps per sum (R = count x 64 blocks): thread l adds the entries of groups l, l + 128, ... in Jacobian
// coordinates, shuffle rounds join the lanes of a warp, thread 0 the four warps.
__global__ void __launch_bounds__(128) k_comb_rows_block(const g1_affine_t *__restrict__ table, const uint32_t *__restrict__ digits,
                                                          uint32_t G, uint64_t E, uint32_t R, g1_jac_t *__restrict__ out) {
    __shared__ g1_jac_t part[4];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x;
    auto load = [&](uint32_t q, g1_affine_t &p) {
        const uint32_t d = digits[(uint64_t)q * R + r];
        const g1_affine_t *src = table + (uint64_t)q * E + (d & 0x7fffffffu);
        ld_fp(p.x, &src->x);
        ld_fp(p.y, &src->y);
        if ((d >> 31) && !g1a_is_inf(p)) fe_neg(p.y, p.y);
    };
    g1_jac_t acc;
    g1j_set_inf(acc);
    g1_affine_t cur;
    g1a_set_inf(cur);
    if (threadIdx.x < G) load(threadIdx.x, cur);
#pragma unroll 1
    for (uint32_t q = threadIdx.x; q < G; q += 128) {
        g1_affine_t nxt;
        g1a_set_inf(nxt);
        if (q + 128 < G) load(q + 128, nxt);
        if (!g1a_is_inf(cur)) g1j_add_affine(acc, acc, cur.x, cur.y);
        cur = nxt;
    }
#pragma unroll 1
    for (int s = 16; s > 0; s >>= 1) {
        g1_jac_t o;
        o.x = fp_shfl(0xffffffffu, acc.x, (int)lane + s);
        o.y = fp_shfl(0xffffffffu, acc.y, (int)lane + s);
        o.z = fp_shfl(0xffffffffu, acc.z, (int)lane + s);
        if ((int)lane < s) g1j_add(acc, acc, o);
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll 1
        for (int w = 1; w < 4; w++) {
            g1_jac_t o = part[w];
            g1j_add(acc, acc, o);
        }
        out[r] = acc;
    }
}

// ------------------------------------------------------------------ launches
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

// One level of additions: `total` independent additions described by `pol`, spread over the whole GPU.
// A thread takes up to max_k additions per batch (one shared inversion per block and batch).
template <class Policy>
static int launch_batch_add(kzg_b200_ctx *ctx, const Policy &pol, uint64_t total) {
    if (total == 0) return KZG_B200_OK;
    const unsigned tpb = KZG_ADD_THREADS;
    const uint64_t t_max = (uint64_t)ctx->sms * ctx->grid_blocks * tpb;
    uint64_t T;
    int k;
    if (total <= t_max) {
        T = (total + tpb - 1) / tpb * tpb;
        k = 1;
    } else {
        T = t_max;
        uint64_t need = (total + T - 1) / T;
        // equal batches: ceil(need / ceil(need / max_k))
        uint64_t batches = (need + ctx->max_k - 1) / ctx->max_k;
        k = (int)((need + batches - 1) / batches);
    }
    RC(ensure_scratch(ctx, (size_t)(T * k)));
    cudaStream_t st = ctx->cur->stream;
    fp_t *scr = ctx->cur->d_scratch;
    const unsigned grid = (unsigned)(T / tpb);
    if (ctx->add_blocks >= 4) batch_add_kernel<Policy, 4><<<grid, tpb, 0, st>>>(pol, total, scr, k);
    else if (ctx->add_blocks == 2) batch_add_kernel<Policy, 2><<<grid, tpb, 0, st>>>(pol, total, scr, k);
    else batch_add_kernel<Policy, 3><<<grid, tpb, 0, st>>>(pol, total, scr, k);
    ctx->launches++;
    CU(cudaGetLastError());
    return KZG_B200_OK;
}

int msm_build_table(kzg_b200_ctx *ctx, const uint8_t *g1_bytes) {
    const int n = ctx->n, g = ctx->g;
    ctx->cur = &ctx->lanes[0];
    cudaStream_t st = ctx->stream;
    DeviceBuf bytes, dec, dst;
    CU(bytes.alloc((size_t)n * 48));
    CU(dec.alloc((size_t)n * sizeof(g1_affine_t)));
    CU(dst.alloc((size_t)n * sizeof(int32_t)));
    CU(cudaMemcpyAsync(bytes.p, g1_bytes, (size_t)n * 48, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dst.p, 0, (size_t)n * sizeof(int32_t), st));
    // reference load_trusted_setup does not subgroup-check the G1 points (src/kzg.rs:859-872)
    RC(g1_launch_decode(st, bytes.as<uint8_t>(), dec.as<g1_affine_t>(), dst.as<int32_t>(), n, 0, n));
    ctx->launches++;
    std::vector<int32_t> status(n);
    CU(cudaMemcpyAsync(status.data(), dst.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++)
        if (status[i] != 0) return KZG_B200_BAD_ARGS;
    k_place_bases<<<blocks_for(ctx->n_pad, 128), 128, 0, st>>>(dec.as<g1_affine_t>(), ctx->d_bases, n, ctx->n_pad);
    k_seed_table<<<blocks_for(ctx->G, 128), 128, 0, st>>>(ctx->d_bases, ctx->d_table, ctx->G, g);
    ctx->launches += 2;
    CU(cudaGetLastError());
    for (int m = 1; m < g; m++) {
        const uint64_t total = (uint64_t)ctx->G << (m - 1);
        // the new slots first, then the old ones in place
        RC(launch_batch_add(ctx, CombLevelPolicy{ctx->d_table, ctx->d_bases, (uint32_t)g, (uint32_t)m, 1u}, total));
        RC(launch_batch_add(ctx, CombLevelPolicy{ctx->d_table, ctx->d_bases, (uint32_t)g, (uint32_t)m, 0u}, total));
    }
    CU(cudaStreamSynchronize(st));
    return KZG_B200_OK;
}

static int comb_index(kzg_b200_ctx *ctx, size_t count) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    const uint64_t threads = (uint64_t)count * ctx->G;
    k_comb_index<<<blocks_for(threads, 128), 128, 0, ln->stream>>>(ln->d_sign_words, (uint32_t)count, (uint32_t)ctx->G, ctx->g,
                                                                   ctx->n_pad, ln->d_digits);
    ctx->launches++;
    CU(cudaGetLastError());
    return KZG_B200_OK;
}
int msm_digits_from_blobs(kzg_b200_ctx *ctx, const uint8_t *d_blobs, size_t count, int32_t *d_status, bool signs_only) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    const uint64_t total = (uint64_t)count * ctx->n_pad;
    stage_begin(ctx, KZG_B200_STAGE_DIGITS);
    k_blob_sign_words<<<blocks_for(total, 256), 256, 0, ln->stream>>>(d_blobs, total, ctx->n, ctx->n_pad, ln->d_sign_words, d_status);
    ctx->launches++;
    int rc = signs_only ? KZG_B200_OK : comb_index(ctx, count);  // the latency comb indexes the sign words itself
    stage_end(ctx, signs_only ? 1 : 2);
    return rc;
}
int msm_digits_from_scalars(kzg_b200_ctx *ctx, const fr_t *d_scalars, size_t count, bool signs_only) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    const uint64_t total = (uint64_t)count * ctx->n_pad;
    stage_begin(ctx, KZG_B200_STAGE_DIGITS);
    k_fr_sign_words<<<blocks_for(total, 256), 256, 0, ln->stream>>>(d_scalars, total, ctx->n, ctx->n_pad, ln->d_sign_words);
    ctx->launches++;
    int rc = signs_only ? KZG_B200_OK : comb_index(ctx, count);
    stage_end(ctx, signs_only ? 1 : 2);
    return rc;
}

int msm_run(kzg_b200_ctx *ctx, size_t count, const g1_affine_t **out, const g1_jac_t **out_jac) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    cudaStream_t st = ln->stream;
    const uint64_t R = (uint64_t)count * ctx->W;  // (bit position, blob) pairs = points per row of a level
    uint32_t rows = (uint32_t)ctx->G;
    if (R * ((rows + 1) / 2) >= (1ull << 32) || R * rows >= (1ull << 32)) return KZG_B200_BAD_ARGS;  // chunk sizes keep every level below 2^32
    const FastDiv fd = FastDiv::make((uint32_t)R);
    // level 0: table entries of groups 2p and 2p + 1; an odd last group is copied
    stage_begin(ctx, KZG_B200_STAGE_MSM_GATHER);
    uint64_t launches = 0;
    if (rows / 2) {
        GatherPolicy gp{ctx->d_table, ln->d_digits, ln->d_buf_a, fd, (uint32_t)ctx->E};
        RC(launch_batch_add(ctx, gp, R * (rows / 2)));
        launches++;
    }
    if (rows & 1) {
        k_gather_copy<<<blocks_for(R, 256), 256, 0, st>>>(ctx->d_table, ln->d_digits, rows - 1, ctx->E, R, ln->d_buf_a + (uint64_t)(rows / 2) * R);
        ctx->launches++;
        launches++;
        CU(cudaGetLastError());
    }
    stage_end(ctx, launches);
    rows = (rows + 1) / 2;
    g1_affine_t *in = ln->d_buf_a, *o = ln->d_buf_b;
    stage_begin(ctx, KZG_B200_STAGE_MSM_TREE);
    launches = 0;
    // the Jacobian sums (144 B each) go into a tree buffer: the smaller one holds ceil(ceil(G / 2) / 2) rows of chunk x 255 points
    const uint64_t small_buf = (uint64_t)ctx->chunk * (((uint64_t)(ctx->G + 1) / 2 + 1) / 2) * sizeof(g1_affine_t);
    const bool jac_tail = out_jac != nullptr && count <= KZG_JAC_TAIL_MAX && (uint64_t)count * sizeof(g1_jac_t) <= small_buf &&
                          env_int("KZG_B200_JAC_TAIL", 1) != 0;
    if (out_jac) *out_jac = nullptr;
    while (rows > 1) {
        if (jac_tail && rows <= KZG_JAC_TAIL_ROWS) {
            g1_jac_t *sums = reinterpret_cast<g1_jac_t *>(o);  // R x 144 B in a buffer of at least 45 R x 96 B
            k_tail_rows_jac<<<blocks_for(R, 128), 128, 0, st>>>(in, rows, R, sums);
            launches++;
            ctx->launches++;
            CU(cudaGetLastError());
            stage_end(ctx, launches);
            *out = nullptr;
            *out_jac = sums;
            return KZG_B200_OK;
        }
        PairPolicy tp{in, o, fd};
        RC(launch_batch_add(ctx, tp, R * (rows / 2)));
        launches++;
        if (rows & 1)
            CU(cudaMemcpyAsync(o + (uint64_t)(rows / 2) * R, in + (uint64_t)(rows - 1) * R, R * sizeof(g1_affine_t), cudaMemcpyDeviceToDevice, st));
        rows = (rows + 1) / 2;
        std::swap(in, o);
    }
    stage_end(ctx, launches);
    *out = in;
    return KZG_B200_OK;
}

// the small-batch form (k_comb_rows_warp): the 255 sums of every blob as Jacobian points, (*out)[j*count + b]
bool msm_small_fits(const kzg_b200_ctx *ctx, size_t count) {  // the Jacobian sums live in the lane's first tree buffer
    return count * sizeof(g1_jac_t) <= ctx->chunk * (((size_t)ctx->G + 1) / 2) * sizeof(g1_affine_t);
}
int msm_run_small(kzg_b200_ctx *ctx, size_t count, const g1_jac_t **out) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    const uint64_t R = (uint64_t)count * ctx->W;
    if (!msm_small_fits(ctx, count)) return KZG_B200_BAD_ARGS;
    g1_jac_t *sums = reinterpret_cast<g1_jac_t *>(ln->d_buf_a);
    stage_begin(ctx, KZG_B200_STAGE_MSM_GATHER);
    k_comb_rows_warp<<<blocks_for(R, 4), 128, 0, ln->stream>>>(ctx->d_table, ln->d_digits, (uint32_t)ctx->G, ctx->E, (uint32_t)R, sums);
    stage_end(ctx, 1);
    ctx->launches++;
    CU(cudaGetLastError());
    *out = sums;
    return KZG_B200_OK;
}

size_t msm_latency_table_bytes(const kzg_b200_ctx *ctx) {
    return (size_t)KZG_LAT_T * ctx->n / KZG_LAT_G * ((size_t)1 << (KZG_LAT_G - 1)) * sizeof(g1_affine_t);
}
int msm_build_latency_table(kzg_b200_ctx *ctx) {
    const uint32_t n = (uint32_t)ctx->n, G = KZG_LAT_T * n / KZG_LAT_G;
    ctx->cur = &ctx->lanes[0];
    cudaStream_t st = ctx->stream;
    k_latency_bases<<<blocks_for(n, 64), 64, 0, st>>>(ctx->d_bases, n, ctx->d_bases_lat);
    k_seed_table<<<blocks_for(G, 128), 128, 0, st>>>(ctx->d_bases_lat, ctx->d_table_lat, G, KZG_LAT_G);
    ctx->launches += 2;
    CU(cudaGetLastError());
    for (int m = 1; m < KZG_LAT_G; m++) {
        const uint64_t total = (uint64_t)G << (m - 1);
        RC(launch_batch_add(ctx, CombLevelPolicy{ctx->d_table_lat, ctx->d_bases_lat, (uint32_t)KZG_LAT_G, (uint32_t)m, 1u}, total));
        RC(launch_batch_add(ctx, CombLevelPolicy{ctx->d_table_lat, ctx->d_bases_lat, (uint32_t)KZG_LAT_G, (uint32_t)m, 0u}, total));
    }
    CU(cudaStreamSynchronize(st));
    return KZG_B200_OK;
}
bool msm_latency_ok(const kzg_b200_ctx *ctx, size_t count) {  // table built; the digits and sums fit the lane's buffers
    if (!ctx->d_table_lat || !msm_small_fits(ctx, count)) return false;
    const uint64_t G = (uint64_t)KZG_LAT_T * ctx->n / KZG_LAT_G;
    return (uint64_t)count * KZG_LAT_ROWS * G <= (uint64_t)ctx->chunk * ctx->W * ctx->G;
}
int msm_run_latency(kzg_b200_ctx *ctx, size_t count, const g1_jac_t **out) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    if (!msm_latency_ok(ctx, count)) return KZG_B200_BAD_ARGS;
    const uint32_t n = (uint32_t)ctx->n, G = KZG_LAT_T * n / KZG_LAT_G;
    const uint64_t R = (uint64_t)count * KZG_LAT_ROWS;
    g1_jac_t *sums = reinterpret_cast<g1_jac_t *>(ln->d_buf_a);
    stage_begin(ctx, KZG_B200_STAGE_MSM_GATHER);
    k_comb_index_latency<<<blocks_for((uint64_t)count * G, 128), 128, 0, ln->stream>>>(ln->d_sign_words, (uint32_t)count, n, ctx->n_pad,
                                                                                       ln->d_digits);
    k_comb_rows_block<<<(unsigned)R, 128, 0, ln->stream>>>(ctx->d_table_lat, ln->d_digits, G, (uint64_t)1 << (KZG_LAT_G - 1), (uint32_t)R, sums);
    stage_end(ctx, 2);
    ctx->launches += 2;
    CU(cudaGetLastError());
    *out = sums;
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ workspace
static size_t rows_after_gather(const kzg_b200_ctx *ctx) { return ((size_t)ctx->G + 1) / 2; }
size_t msm_workspace_per_blob(const kzg_b200_ctx *ctx) {
    const size_t W = (size_t)ctx->W, ra = rows_after_gather(ctx), rb = (ra + 1) / 2;
    return 8 * (size_t)ctx->n_pad * 4 + W * ctx->G * 4 + W * (ra + rb) * sizeof(g1_affine_t);
}
int msm_alloc_lane(kzg_b200_ctx *ctx, kzg_b200_ctx::Lane &ln, size_t chunk) {
    const size_t W = (size_t)ctx->W, ra = rows_after_gather(ctx), rb = (ra + 1) / 2;
    CU(cudaMalloc(&ln.d_sign_words, chunk * 8 * (size_t)ctx->n_pad * sizeof(uint32_t)));
    CU(cudaMalloc(&ln.d_digits, chunk * W * ctx->G * sizeof(uint32_t)));
    CU(cudaMalloc(&ln.d_buf_a, chunk * W * ra * sizeof(g1_affine_t)));
    CU(cudaMalloc(&ln.d_buf_b, chunk * W * rb * sizeof(g1_affine_t)));
    return KZG_B200_OK;
}
void msm_free_lane(kzg_b200_ctx::Lane &ln) {
    cudaFree(ln.d_sign_words); cudaFree(ln.d_digits); cudaFree(ln.d_buf_a); cudaFree(ln.d_buf_b); cudaFree(ln.d_scratch);
    ln.d_sign_words = ln.d_digits = nullptr;
    ln.d_buf_a = ln.d_buf_b = nullptr;
    ln.d_scratch = nullptr;
    ln.scratch_elems = 0;
}

// kzg_b200.cu -- context and C ABI of the B200 blob path (see include/kzg_b200.h; kernels: msm.cu, g1ops.cu, frops.cu).
//
// Host code here only sequences kernels and moves bytes; every arithmetic step of the
// path runs on the GPU.  There is deliberately no CPU fallback: if CUDA is unusable the
// entry points return KZG_B200_CUDA_ERROR.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>

#include "internal.h"

using namespace kzg;

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// ------------------------------------------------------------------ stage timing
void stage_begin(kzg_b200_ctx *ctx, int stage) {
    if (!ctx->profile) return;
    kzg_b200_ctx::StageRec r;
    r.stage = stage;
    if (cudaEventCreate(&r.a) != cudaSuccess) return;
    if (cudaEventCreate(&r.b) != cudaSuccess) { cudaEventDestroy(r.a); return; }
    cudaEventRecord(r.a, ctx->cur->stream);
    ctx->pending.push_back(r);
    ctx->stage_open = true;
}
void stage_end(kzg_b200_ctx *ctx, uint64_t launches) {
    if (!ctx->profile || !ctx->stage_open) return;
    ctx->stage_open = false;
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
    ctx->stage_open = false;
}

// ------------------------------------------------------------------ workspace
static size_t per_blob_workspace(const kzg_b200_ctx *ctx) {
    size_t lane = msm_workspace_per_blob(ctx) + 2 * (size_t)ctx->n * sizeof(fr_t) /*poly, inv*/ + 64 + sizeof(fr_t) +
                  2 * sizeof(g1_affine_t);
    return ctx->nlanes * lane + KZG_SLOTS * ((size_t)ctx->n * 32 + 96 * 2 + 4);
}
// the batches of the addition kernel keep their prefix products here (grown on demand up to this)
static size_t scratch_bytes(const kzg_b200_ctx *ctx) {
    return (size_t)ctx->nlanes * ctx->sms * ctx->grid_blocks * 128 * (size_t)ctx->max_k * sizeof(fp_t);
}
static void free_workspace(kzg_b200_ctx *ctx) {
    for (auto &ln : ctx->lanes) {
        msm_free_lane(ln);
        cudaFree(ln.d_poly); cudaFree(ln.d_inv); cudaFree(ln.d_z); cudaFree(ln.d_zy); cudaFree(ln.d_pts); cudaFree(ln.d_sha_state);
        ln.d_poly = ln.d_inv = ln.d_z = nullptr;
        ln.d_sha_state = nullptr;
        ln.d_zy = nullptr;
        ln.d_pts = nullptr;
    }
    cudaFree(ctx->d_stage_in); cudaFree(ctx->d_stage_aux); cudaFree(ctx->d_stage_out); cudaFree(ctx->d_status);
    ctx->d_stage_in = ctx->d_stage_aux = ctx->d_stage_out = nullptr;
    ctx->d_status = nullptr;
    ctx->chunk = 0;
}
static int alloc_workspace(kzg_b200_ctx *ctx, size_t chunk) {
    free_workspace(ctx);
    for (int l = 0; l < ctx->nlanes; l++) {
        kzg_b200_ctx::Lane &ln = ctx->lanes[l];
        RC(msm_alloc_lane(ctx, ln, chunk));
        CU(cudaMalloc(&ln.d_poly, chunk * (size_t)ctx->n * sizeof(fr_t)));
        CU(cudaMalloc(&ln.d_inv, chunk * (size_t)ctx->n * sizeof(fr_t)));
        CU(cudaMalloc(&ln.d_z, chunk * sizeof(fr_t)));
        CU(cudaMalloc(&ln.d_sha_state, chunk * 8 * sizeof(uint32_t)));
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
static size_t table_bytes(size_t n, int g) { return ((n + g - 1) / g) * ((size_t)1 << (g - 1)) * sizeof(g1_affine_t); }

// Everything after `new kzg_b200_ctx` fails through kzg_b200_ctx_destroy, which releases whatever exists.
static int ctx_init(kzg_b200_ctx *ctx, const uint8_t *g1_lagrange, size_t n1, const uint8_t *g2_monomial, int device, int comb_width) {
    ctx->device = device;
    ctx->n = (int)n1;
    memcpy(ctx->g2_tau, g2_monomial + 96, 96);
    ctx->tau_prepared = host_g2_prepare(ctx->g2_tau);
    if (!ctx->tau_prepared) return KZG_B200_BAD_ARGS;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    ctx->sms = prop.multiProcessorCount;
    ctx->max_k = std::max(1, env_int("KZG_B200_BATCH_K", 1024));
    ctx->add_blocks = std::min(4, std::max(2, env_int("KZG_B200_ADD_BLOCKS", 3)));
    ctx->grid_blocks = std::max(1, env_int("KZG_B200_GRID_BLOCKS", ctx->add_blocks));
    ctx->nlanes = env_int("KZG_B200_LANES", 2) >= 2 ? 2 : 1;
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    ctx->lanes[0].stream = ctx->stream;
    CU(cudaStreamCreateWithFlags(&ctx->lanes[1].stream, cudaStreamNonBlocking));
    ctx->cur = &ctx->lanes[0];
    CU(cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&ctx->lanes[i].ev_done, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&ctx->lanes[i].side_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->lanes[i].ev_side_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->lanes[i].ev_side_join, cudaEventDisableTiming));
    }
    for (int i = 0; i < KZG_SLOTS; i++) {
        CU(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_aux[i], cudaEventDisableTiming));
        for (int k = 0; k < KZG_HASH_SLICES; k++) CU(cudaEventCreateWithFlags(&ctx->ev_slice[i][k], cudaEventDisableTiming));
    }
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    // KZG_B200_MEM_BUDGET_GB: plan as if only this much device memory were free (a GPU shared with other tenants; tests)
    const size_t budget = (size_t)std::max(0, env_int("KZG_B200_MEM_BUDGET_GB", 0)) << 30;
    if (budget && budget < free_b) free_b = budget;
    const size_t chunk_cap = (size_t)std::max(1, env_int("KZG_B200_CHUNK", 4096));
    const size_t reserve = std::min((size_t)8 << 30, free_b / 8);  // left for the caller's device-resident blobs
    int g = comb_width > 0 ? comb_width : env_int("KZG_B200_COMB_WIDTH", 0);
    if (g <= 0) {
        // The widest comb whose table (a) takes at most half of the free device memory and (b) still leaves room for
        // a 256-blob workspace and the reserve; every step of g halves the table and costs about 4.5 % more additions
        // (g = 23 on an empty 180 GB B200: 72 GB of table, 178 additions per bit position).  Floor 8 (6 MB).
        for (g = KZG_COMB_MAX_WIDTH; g > 8; g--) {
            ctx->g = g;
            ctx->G = ((int)n1 + g - 1) / g;
            ctx->n_pad = ctx->G * g;
            const size_t tbl = table_bytes(n1, g);
            if (tbl <= free_b / 2 && tbl + 256 * per_blob_workspace(ctx) + scratch_bytes(ctx) + reserve <= free_b) break;
        }
    }
    g = std::min<int>(g, (int)n1);  // one group holds the whole minimal preset
    if (g < 1 || g > KZG_COMB_MAX_WIDTH) return KZG_B200_BAD_ARGS;
    ctx->g = g;
    ctx->G = ((int)n1 + g - 1) / g;
    ctx->n_pad = ctx->G * g;
    ctx->E = (uint64_t)1 << (g - 1);
    CU(cudaMalloc(&ctx->d_table, table_bytes(n1, g)));
    CU(cudaMalloc(&ctx->d_bases, (size_t)ctx->n_pad * sizeof(g1_affine_t)));
    RC(msm_build_table(ctx, g1_lagrange));
    RC(fr_setup_roots_device(ctx->n, &ctx->d_roots, ctx->stream));
    // workspace: as many blobs per chunk as memory allows, capped (KZG_B200_CHUNK overrides)
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (budget) free_b = std::min(free_b, budget > table_bytes(n1, g) ? budget - table_bytes(n1, g) : (size_t)0);
    const size_t fixed = scratch_bytes(ctx) + reserve;
    const size_t usable = free_b > fixed + ((size_t)1 << 30) ? free_b - fixed : free_b / 2;
    const size_t chunk = std::max<size_t>(1, std::min(chunk_cap, usable / per_blob_workspace(ctx)));
    RC(alloc_workspace(ctx, chunk));
    // the latency comb of small calls (internal.h), when it fits beside everything else: not fatal if it does not
    if (n1 % KZG_LAT_G == 0 && n1 >= 64 && env_int("KZG_B200_LATENCY_TABLE", 1) != 0) {
        const size_t lat = msm_latency_table_bytes(ctx) + (size_t)KZG_LAT_T * n1 * sizeof(g1_affine_t);
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t used = table_bytes(n1, g) + chunk * per_blob_workspace(ctx) + scratch_bytes(ctx);
        if (lat + reserve <= free_b && (!budget || used + lat + reserve <= budget)) {
            if (cudaMalloc(&ctx->d_table_lat, msm_latency_table_bytes(ctx)) == cudaSuccess &&
                cudaMalloc(&ctx->d_bases_lat, (size_t)KZG_LAT_T * n1 * sizeof(g1_affine_t)) == cudaSuccess) {
                RC(msm_build_latency_table(ctx));
            } else {
                (void)cudaGetLastError();
                cudaFree(ctx->d_table_lat);
                ctx->d_table_lat = nullptr;
            }
        }
    }
    return KZG_B200_OK;
}

extern "C" int kzg_b200_ctx_create(const uint8_t *g1_lagrange, size_t n1, const uint8_t *g2_monomial, size_t n2,
                                   int device, int comb_width, kzg_b200_ctx **out) {
    if (!g1_lagrange || !g2_monomial || !out) return KZG_B200_BAD_ARGS;
    *out = nullptr;
    // reference src/kzg.rs:843-847 (n1 fixed per preset; both presets accepted here)
    if ((n1 != 4096 && n1 != 4) || n2 != KZG_B200_NUM_G2_POINTS) return KZG_B200_BAD_ARGS;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return KZG_B200_CUDA_ERROR;
    CU(cudaSetDevice(device));
    // G2 side of the setup on the host: decode all 65 points (reference src/kzg.rs:874-887) and
    // run the Lagrange-form sanity pairing (src/kzg.rs:802-830) -- it needs g1[0], g1[1] only.
    RC(host_check_setup(g1_lagrange, g2_monomial, n2));
    kzg_b200_ctx *ctx = new kzg_b200_ctx();
    int rc = ctx_init(ctx, g1_lagrange, n1, g2_monomial, device, comb_width);
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
extern "C" int kzg_b200_ctx_create_from_file(const char *path, int device, int comb_width, kzg_b200_ctx **out) {
    if (!path || !out) return KZG_B200_BAD_ARGS;
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
    return kzg_b200_ctx_create(g1.data(), n1, g2.data(), n2, device, comb_width, out);
}

extern "C" void kzg_b200_ctx_destroy(kzg_b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->lanes[1].stream) cudaStreamSynchronize(ctx->lanes[1].stream);
    free_workspace(ctx);
    cudaFree(ctx->d_vb);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (cudaEvent_t e : ctx->ev_chunks) cudaEventDestroy(e);
    cudaFree(ctx->d_z_all);
    cudaFree(ctx->d_sums_all);
    cudaFree(ctx->d_table_lat);
    cudaFree(ctx->d_bases_lat);
    cudaFree(ctx->d_table);
    cudaFree(ctx->d_bases);
    cudaFree(ctx->d_roots);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->lanes[1].stream) cudaStreamDestroy(ctx->lanes[1].stream);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    for (int i = 0; i < 2; i++) {
        kzg_b200_ctx::Lane &ln = ctx->lanes[i];
        if (ln.side_stream) { cudaStreamSynchronize(ln.side_stream); cudaStreamDestroy(ln.side_stream); }
        if (ln.ev_done) cudaEventDestroy(ln.ev_done);
        if (ln.ev_side_fork) cudaEventDestroy(ln.ev_side_fork);
        if (ln.ev_side_join) cudaEventDestroy(ln.ev_side_join);
    }
    for (int i = 0; i < KZG_SLOTS; i++) {
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
        if (ctx->ev_aux[i]) cudaEventDestroy(ctx->ev_aux[i]);
        for (int k = 0; k < KZG_HASH_SLICES; k++)
            if (ctx->ev_slice[i][k]) cudaEventDestroy(ctx->ev_slice[i][k]);
    }
    stage_collect(ctx);
    host_g2_prepared_free(ctx->tau_prepared);
    delete ctx;
}
extern "C" size_t kzg_b200_field_elements_per_blob(const kzg_b200_ctx *ctx) { return ctx ? (size_t)ctx->n : 0; }
extern "C" int kzg_b200_comb_width(const kzg_b200_ctx *ctx) { return ctx ? ctx->g : 0; }
extern "C" size_t kzg_b200_table_bytes(const kzg_b200_ctx *ctx) { return ctx ? table_bytes((size_t)ctx->n, ctx->g) : 0; }
extern "C" size_t kzg_b200_chunk_blobs(const kzg_b200_ctx *ctx) { return ctx ? ctx->chunk : 0; }
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
// one chunk, everything on the device: blobs -> comb digits -> MSM -> 48-byte commitments
// The Horner + compression pass is one thread per blob and ~4600 dependent products long (254 doublings and
// additions): for a single chunk it is milliseconds of latency on a mostly idle GPU.  Device-resident calls that
// span several chunks therefore park the sums of every chunk in one [bit position][blob] array of the whole call
// (1.6 GB for 65,536 blobs) and run the pass ONCE at the end, where 65,536 threads fill the machine.
struct DeferredCompress {
    g1_affine_t *sums = nullptr;  // nullptr: compress per chunk
    size_t n = 0;                 // blobs of the whole call = row pitch of `sums`
    uint8_t *out = nullptr;       // n x 48 B and n x int32 of the whole call, for host-buffer calls (they come after `sums`)
    int32_t *status = nullptr;
};
// several_pieces: the call runs as more than one chunk although it would fit one (host calls cut by msm_piece_plan)
static int deferred_begin(kzg_b200_ctx *ctx, size_t n, DeferredCompress *dc, bool several_pieces = false) {
    dc->sums = nullptr;
    dc->n = n;
    if (n <= ctx->chunk && !several_pieces) return KZG_B200_OK;
    const size_t elems = n * (size_t)ctx->W + n;  // + n x 96 B: room for the call's outputs (48 B) and status words
    if (elems > ctx->sums_all_elems) {
        if (ctx->d_sums_all) CU(cudaFree(ctx->d_sums_all));
        ctx->d_sums_all = nullptr;
        ctx->sums_all_elems = 0;
        CU(cudaMalloc(&ctx->d_sums_all, elems * sizeof(g1_affine_t)));
        ctx->sums_all_elems = elems;
    }
    dc->sums = ctx->d_sums_all;
    dc->out = reinterpret_cast<uint8_t *>(ctx->d_sums_all + n * (size_t)ctx->W);
    dc->status = reinterpret_cast<int32_t *>(dc->out + 48 * n);
    return KZG_B200_OK;
}
// after msm_run of one chunk (its sums are [bit position][count] in the lane's buffer): compress now, or park them
static int compress_or_park(kzg_b200_ctx *ctx, const g1_affine_t *res, size_t off, size_t count, const int32_t *d_status,
                            uint8_t *d_out, const DeferredCompress *dc) {
    cudaStream_t st = ctx->cur->stream;
    if (dc && dc->sums) {
        CU(cudaMemcpy2DAsync(dc->sums + off, dc->n * sizeof(g1_affine_t), res, count * sizeof(g1_affine_t),
                             count * sizeof(g1_affine_t), (size_t)ctx->W, cudaMemcpyDeviceToDevice, st));
        return KZG_B200_OK;
    }
    stage_begin(ctx, KZG_B200_STAGE_COMPRESS);
    int rc = g1_launch_horner_compress(st, res, count, ctx->W, d_status, d_out, count);
    stage_end(ctx, 1);
    ctx->launches++;
    return rc;
}
// once per call, on the caller-visible stream after the lanes have joined
static int deferred_finish(kzg_b200_ctx *ctx, const DeferredCompress *dc, const int32_t *d_status, uint8_t *d_out) {
    if (!dc->sums) return KZG_B200_OK;
    ctx->cur = &ctx->lanes[0];
    stage_begin(ctx, KZG_B200_STAGE_COMPRESS);
    int rc = g1_launch_horner_compress(ctx->stream, dc->sums, dc->n, ctx->W, d_status, d_out, dc->n);
    stage_end(ctx, 1);
    ctx->launches++;
    return rc;
}

// The MSM of one chunk from the digits of the current lane, down to the compressed points (or the parked sums).
// Small single-chunk calls take the one-warp-per-sum form (msm_run_small; KZG_B200_MSM_SMALL_MAX blobs, 0 = never).
// which form the MSM of a chunk takes: 0 = batched affine levels, 1 = one warp per sum, 2 = the latency comb
static int msm_form(const kzg_b200_ctx *ctx, size_t count, const DeferredCompress *dc) {
    const size_t small_max = (size_t)std::max(0, std::min(KZG_MSM_SMALL_CAP, env_int("KZG_B200_MSM_SMALL_MAX", KZG_MSM_SMALL_MAX)));
    if (count > small_max || (dc && dc->sums) || !msm_small_fits(ctx, count)) return 0;
    return msm_latency_ok(ctx, count) && env_int("KZG_B200_LATENCY_TABLE", 1) != 0 ? 2 : 1;
}
static int msm_and_compress(kzg_b200_ctx *ctx, size_t off, size_t count, const int32_t *d_status, uint8_t *d_out,
                            const DeferredCompress *dc, cudaEvent_t join = nullptr) {
    cudaStream_t st = ctx->cur->stream;
    const int form = msm_form(ctx, count, dc);
    if (form != 0) {
        const g1_jac_t *sums = nullptr;
        // with the latency comb (internal.h): 64 sums and 63 doublings per blob instead of 255 and 254
        const bool lat = form == 2;
        if (lat) RC(msm_run_latency(ctx, count, &sums));
        else RC(msm_run_small(ctx, count, &sums));
        if (join) CU(cudaStreamWaitEvent(st, join, 0));
        stage_begin(ctx, KZG_B200_STAGE_COMPRESS);
        int rc = g1_launch_horner_compress_jac(st, sums, count, lat ? KZG_LAT_ROWS : ctx->W, d_status, d_out, count);
        stage_end(ctx, 1);
        ctx->launches++;
        return rc;
    }
    const g1_affine_t *res = nullptr;
    const g1_jac_t *jsums = nullptr;
    RC(msm_run(ctx, count, &res, (dc && dc->sums) ? nullptr : &jsums));  // parked sums are affine
    if (join) CU(cudaStreamWaitEvent(st, join, 0));
    if (jsums) {
        stage_begin(ctx, KZG_B200_STAGE_COMPRESS);
        int rc = g1_launch_horner_compress_jac(st, jsums, count, ctx->W, d_status, d_out, count);
        stage_end(ctx, 1);
        ctx->launches++;
        return rc;
    }
    return compress_or_park(ctx, res, off, count, d_status, d_out, dc);
}

// d_out / d_status point at this chunk's slice; `off` is its first blob within the call (deferred compression)
static int commit_chunk(kzg_b200_ctx *ctx, const uint8_t *d_blobs, size_t count, uint8_t *d_out, int32_t *d_status,
                        size_t off = 0, const DeferredCompress *dc = nullptr) {
    CU(cudaMemsetAsync(d_status, 0, count * sizeof(int32_t), ctx->cur->stream));
    RC(msm_digits_from_blobs(ctx, d_blobs, count, d_status, msm_form(ctx, count, dc) == 2));
    return msm_and_compress(ctx, off, count, d_status, d_out, dc);
}

// device pointers handed to the *_device entry points are read with 128-bit loads
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int kzg_b200_blob_to_kzg_commitment_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, size_t n, uint8_t *d_out,
                                                      int32_t *d_status) {
    if (!ctx || (n && (!d_blobs || !d_out || !d_status))) return KZG_B200_BAD_ARGS;
    if (!aligned16(d_blobs) || !aligned16(d_out)) return KZG_B200_BAD_ARGS;
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
// finish() runs once after the lanes have joined, on the caller-visible stream, before the call waits for it.
// piece: blobs per chunk of this call (at most ctx->chunk, the size of a staging slot); first: blobs of the first chunk
// (0 = like the others) -- nothing overlaps the upload of the first chunk, so it should be short.
struct PiecePlan { size_t first, piece; };
// Commitment / proof calls over host buffers.  A call that fits one chunk is cut in two halves from 512 blobs on (the second
// half uploads under the first one's kernels and the two lanes overlap: 1,024 blobs 18.3 -> 16.6 ms, 4,096 blobs 67.8 -> 58 ms,
// profiles/mid_chunks_r2as.txt); a larger call starts with a 1,024-blob chunk (2.5 ms of exposed upload instead of 10).
static PiecePlan msm_piece_plan(const kzg_b200_ctx *ctx, size_t n) {
    const size_t ch = ctx->chunk;
    if (env_int("KZG_B200_PLAIN_PIECES", 0)) return {0, ch};
    if (n <= ch) return n >= 512 ? PiecePlan{0, (n + 1) / 2} : PiecePlan{0, ch};
    return {std::min<size_t>(ch, 1024), ch};
}
template <class Upload, class Run, class Finish>
static int staged_chunks(kzg_b200_ctx *ctx, size_t n, size_t piece, Upload upload, Run run, Finish finish, size_t first = 0) {
    piece = std::max<size_t>(1, std::min(piece, ctx->chunk));
    if (first == 0 || first > piece) first = piece;
    first = std::min(first, n);
    const size_t nchunks = n <= first ? 1 : 1 + (n - first + piece - 1) / piece;
    auto piece_off = [&](size_t i) { return i == 0 ? (size_t)0 : first + (i - 1) * piece; };
    auto piece_cnt = [&](size_t i) { return i == 0 ? first : std::min(piece, n - piece_off(i)); };
    const size_t ahead = KZG_SLOTS - 1;
    ctx->call_blobs = n;
    const bool trace = env_int("KZG_B200_TRACE", 0) == 2;  // host wall clock of the enqueue / wait phases on stderr
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[kzg_b200 trace]   staged: %-10s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    auto enqueue_upload = [&](size_t i) -> int {
        int slot = (int)(i % KZG_SLOTS);
        size_t off = piece_off(i), cnt = piece_cnt(i);
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[slot], 0));  // the chunk that used this slot is done
        ctx->aux_recorded[slot] = false;
        ctx->slices_recorded[slot] = 0;
        RC(upload(slot, off, cnt));
        CU(cudaEventRecord(ctx->ev_h2d[slot], ctx->copy_stream));
        return KZG_B200_OK;
    };
    RC(lanes_begin(ctx));
    // the slots may still be in use by an earlier asynchronous call on `stream`
    for (int sl = 0; sl < KZG_SLOTS; sl++) CU(cudaEventRecord(ctx->ev_free[sl], ctx->stream));
    for (size_t i = 0; i < std::min(ahead, nchunks); i++) RC(enqueue_upload(i));
    lap("uploads");
    for (size_t i = 0; i < nchunks; i++) {
        int slot = (int)(i % KZG_SLOTS);
        size_t off = piece_off(i), cnt = piece_cnt(i);
        if (i + ahead < nchunks) RC(enqueue_upload(i + ahead));
        lane_select(ctx, i);
        // a sliced upload is waited for slice by slice by the chunk itself (verify_chunk_a); its last slice event is ev_h2d's equal
        ctx->cur_slices = ctx->slices_recorded[slot];
        ctx->cur_slice_ev = ctx->ev_slice[slot];
        if (!ctx->cur_slices) CU(cudaStreamWaitEvent(ctx->cur->stream, ctx->ev_h2d[slot], 0));
        ctx->aux_ready = ctx->aux_recorded[slot] ? ctx->ev_aux[slot] : nullptr;
        int run_rc = run(slot, off, cnt);
        ctx->aux_ready = nullptr;
        ctx->cur_slices = 0;
        RC(run_rc);
        CU(cudaEventRecord(ctx->ev_free[slot], ctx->cur->stream));
    }
    lap("enqueue");
    RC(lanes_end(ctx));
    RC(finish());
    CU(cudaStreamSynchronize(ctx->stream));
    lap("wait");
    stage_collect(ctx);
    return KZG_B200_OK;
}
template <class Upload, class Run>
static int staged_chunks(kzg_b200_ctx *ctx, size_t n, size_t piece, Upload upload, Run run, size_t first = 0) {
    return staged_chunks(ctx, n, piece, upload, run, []() -> int { return KZG_B200_OK; }, first);
}
// grow-only pinned host buffer of the context (results that the host reads while the GPU works on the next chunk)
static int ensure_pinned(kzg_b200_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->h_pin_bytes) return KZG_B200_OK;
    if (ctx->h_pin) CU(cudaFreeHost(ctx->h_pin));
    ctx->h_pin = nullptr;
    ctx->h_pin_bytes = 0;
    CU(cudaHostAlloc((void **)&ctx->h_pin, bytes + bytes / 2, cudaHostAllocDefault));
    ctx->h_pin_bytes = bytes + bytes / 2;
    return KZG_B200_OK;
}

extern "C" int kzg_b200_blob_to_kzg_commitment_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out,
                                                     int32_t *status) {
    if (!ctx || (n && (!blobs || !out || !status))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t bpb = (size_t)ctx->n * 32, ch = ctx->chunk;
    // calls over several chunks run the Horner + compression pass once at the end, like the device-resident form
    DeferredCompress dc;
    const PiecePlan plan = msm_piece_plan(ctx, n);
    RC(deferred_begin(ctx, n, &dc, n > std::min(plan.piece, ctx->chunk)));
    return staged_chunks(
        ctx, n, plan.piece,
        [&](int slot, size_t off, size_t cnt) -> int {
            CU(cudaMemcpyAsync(ctx->d_stage_in + slot * ch * bpb, blobs + off * bpb, cnt * bpb, cudaMemcpyHostToDevice, ctx->copy_stream));
            return KZG_B200_OK;
        },
        [&](int slot, size_t off, size_t cnt) -> int {
            uint8_t *d_out = dc.sums ? dc.out + off * 48 : ctx->d_stage_out + slot * ch * 96;
            int32_t *d_st = dc.sums ? dc.status + off : ctx->d_status + slot * ch;
            RC(commit_chunk(ctx, ctx->d_stage_in + slot * ch * bpb, cnt, d_out, d_st, off, &dc));
            if (!dc.sums) {
                CU(cudaMemcpyAsync(out + off * 48, d_out, cnt * 48, cudaMemcpyDeviceToHost, ctx->cur->stream));
                CU(cudaMemcpyAsync(status + off, d_st, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->cur->stream));
            }
            return KZG_B200_OK;
        },
        [&]() -> int {
            if (!dc.sums) return KZG_B200_OK;
            RC(deferred_finish(ctx, &dc, dc.status, dc.out));
            CU(cudaMemcpyAsync(out, dc.out, n * 48, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(status, dc.status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            return KZG_B200_OK;
        },
        plan.first);
}

// ------------------------------------------------------------------ roofline micro-benchmarks, test aids
extern "C" int kzg_b200_measure_peaks(kzg_b200_ctx *ctx, double *imad_per_s, double *imad_wide_per_s,
                                      double *fp_mul_per_s) {
    if (!ctx) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    return g1_measure_peaks(ctx, imad_per_s, imad_wide_per_s, fp_mul_per_s);
}
// raw table entries, affine Montgomery limbs: entry `first + k` of the flat table (group q, index idx -> q * 2^(g-1) + idx)
extern "C" int kzg_b200_debug_table(kzg_b200_ctx *ctx, uint64_t first, uint64_t count, void *out) {
    if (!ctx || !out) return KZG_B200_BAD_ARGS;
    if (first + count > (uint64_t)ctx->G * ctx->E) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, ctx->d_table + first, count * sizeof(g1_affine_t), cudaMemcpyDeviceToHost));
    return KZG_B200_OK;
}
extern "C" int kzg_b200_debug_field_op(kzg_b200_ctx *ctx, int op, const uint32_t *a, const uint32_t *b, uint32_t *out,
                                       uint64_t count) {
    if (!ctx || !a || !b || !out) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    return g1_debug_field_op(ctx, op, a, b, out, count);
}

// see include/kzg_b200.h: commitment_i == [p_i(tau)] G1 for every blob, tau known (testing setups only)
extern "C" int kzg_b200_debug_check_tau_identity(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments,
                                                 size_t n, const uint8_t tau[32], int32_t *d_ok) {
    if (!ctx || !tau || (n && (!d_blobs || !d_commitments || !d_ok))) return KZG_B200_BAD_ARGS;
    if (!aligned16(d_blobs)) return KZG_B200_BAD_ARGS;
    fr_t t;
    scalar_from_be32(t, tau);
    if (!fr_is_canonical(t)) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    kzg_b200_ctx::Lane *ln = ctx->cur = &ctx->lanes[0];
    const size_t bpb = (size_t)ctx->n * 32;
    std::vector<fr_t> zs(std::min(n, ctx->chunk), t);
    for (size_t off = 0; off < n; off += ctx->chunk) {
        const size_t cnt = std::min(ctx->chunk, n - off);
        CU(cudaMemcpyAsync(ln->d_z, zs.data(), cnt * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_status, 0, cnt * sizeof(int32_t), ctx->stream));
        RC(fr_launch_eval(ctx->stream, 0, d_blobs + off * bpb, ln->d_z, ctx->d_roots, ctx->n, ln->d_inv, ln->d_poly, ln->d_zy,
                          ctx->d_status, cnt));
        RC(g1_launch_tau_identity(ctx->stream, d_commitments + off * 48, ln->d_zy, cnt, d_ok + off));
        ctx->launches += 2;
        CU(cudaStreamSynchronize(ctx->stream));  // zs is reused by the next chunk's copy
    }
    return KZG_B200_OK;
}

#include "host_sha256.h"
#include "proof_verify.inl"

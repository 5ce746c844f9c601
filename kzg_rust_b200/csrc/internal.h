// internal.h -- what the translation units of libkzg_b200.so share: the context, the error macros and the
// host-side launch functions of the kernels (a kernel is only launched from the file that defines it).
//
//   kzg_b200.cu   context, C ABI, sequencing of the commitment / proof / verification paths
//   msm.cu        the MSM: comb table, scalar recoding, batched affine additions        (msm.cuh)
//   g1ops.cu      point decoding + subgroup checks, Horner + compression, micro-benchmarks (g1.cuh, blobpath.cuh)
//   frops.cu      challenges, evaluation + quotient, r-power terms of batch verification  (frpath.cuh)
//   host_pairing.cpp   the final pairing check on the host
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/kzg_b200.h"
#include "blobpath.cuh"
#include "host_pairing.h"

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

// A temporary device allocation that is released on every way out of a function.
struct DeviceBuf {
    void *p = nullptr;
    DeviceBuf() = default;
    DeviceBuf(const DeviceBuf &) = delete;
    DeviceBuf &operator=(const DeviceBuf &) = delete;
    ~DeviceBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

#define KZG_SLOTS 3
#define KZG_HASH_SLICES 8
struct kzg_b200_ctx {
    int device = 0;
    int n = 0;            // FIELD_ELEMENTS_PER_BLOB
    int g = 0;            // comb width: setup points per table group
    int G = 0;            // groups: ceil(n / g)
    int n_pad = 0;        // G * g (the padding points of the last group are points at infinity)
    int W = KZG_COMB_WINDOWS;  // bit positions of a scalar = sums per blob
    uint64_t E = 0;       // table entries per group: 2^(g-1)
    int sms = 0;
    int max_k = 512;      // additions per thread per batch (one shared inversion per block and batch)
    int add_blocks = 3;   // resident blocks per SM the addition kernel is compiled for (KZG_B200_ADD_BLOCKS: 2, 3 or 4)
    int grid_blocks = 3;  // blocks per SM a launch asks for (KZG_B200_GRID_BLOCKS)
    kzg::g1_affine_t *d_table = nullptr;
    kzg::g1_affine_t *d_bases = nullptr;  // the n_pad setup points, bit-reversal permuted (src/kzg.rs:895-896)
    // The latency comb (mainnet, optional): the Horner pass of a commitment is 255 dependent doublings -- most of what a
    // small call takes.  A second, small table over the 4 n "virtual" points 2^(64 t) G_i (t < 4; groups of 16, 2^15 entries
    // each, 3.2 GB) lets bit positions j, j + 64, j + 128, j + 192 share one sum: 64 sums of 1024 entries and 63 doublings
    // per commitment.  Used by single-chunk calls of up to KZG_MSM_SMALL_MAX blobs (msm_run_latency).
    kzg::g1_affine_t *d_table_lat = nullptr;
    kzg::g1_affine_t *d_bases_lat = nullptr;  // [t][i] = 2^(64 t) bases[i]
    kzg::fr_t *d_roots = nullptr;         // roots of unity, Montgomery form, bit-reversed (src/kzg.rs:764-799)
    uint8_t g2_tau[96];                   // [tau]G2 = g2_values[1]
    // Work is cut into chunks of `chunk` blobs.  Two chunks are in flight at a time, each on its
    // own lane (stream + workspace), so the latency-bound end of one chunk (the small levels of the
    // addition tree, the Horner pass) runs under the big levels of the next.  Lane 0 launches on
    // `stream`, the stream callers synchronise with; `cur` is the lane of the chunk being enqueued
    // (calls on a context are serialised by `mu`).
    struct Lane {
        cudaStream_t stream = nullptr;
        uint32_t *d_sign_words = nullptr;  // chunk x 8 x n_pad: the 255 signs of every scalar
        uint32_t *d_digits = nullptr;      // [group][bit position][blob]: table index + sign, what the gather level reads
        kzg::g1_affine_t *d_buf_a = nullptr, *d_buf_b = nullptr;
        kzg::fp_t *d_scratch = nullptr;
        size_t scratch_elems = 0;
        kzg::fr_t *d_poly = nullptr;       // chunk x n evaluations (proof / verify paths)
        kzg::fr_t *d_inv = nullptr;        // chunk x n prefix products, then 1/(z - w_i), then the quotient
        kzg::fr_t *d_z = nullptr;          // chunk challenges / evaluation points (canonical)
        uint32_t *d_sha_state = nullptr;   // chunk x 8 words: SHA-256 states between the slice launches of a chunk's hash
        uint8_t *d_zy = nullptr;           // chunk x 64 B: z || y big-endian
        kzg::g1_affine_t *d_pts = nullptr; // chunk x 2 decoded commitments / proofs
        cudaEvent_t ev_done = nullptr;
        // point validation runs beside the challenge hash of the same chunk (both are latency-bound kernels)
        cudaStream_t side_stream = nullptr;
        cudaEvent_t ev_side_fork = nullptr, ev_side_join = nullptr;
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
    cudaEvent_t ev_h2d[KZG_SLOTS] = {nullptr, nullptr, nullptr}, ev_free[KZG_SLOTS] = {nullptr, nullptr, nullptr};
    // a slot's commitments / proofs are uploaded BEFORE its blobs and get their own event: their validation starts while the
    // blobs are still on their way (and so never starts at the same instant as the hash, see verify_chunk_a)
    cudaEvent_t ev_aux[KZG_SLOTS] = {nullptr, nullptr, nullptr};
    bool aux_recorded[KZG_SLOTS] = {false, false, false};
    cudaEvent_t aux_ready = nullptr;  // ev_aux of the chunk being enqueued, when its upload recorded one
    // Host verification uploads the blobs of a chunk in KZG_HASH_SLICES column slices (slice k = bytes [k W, (k+1) W) of EVERY
    // blob of the chunk, one 2-D copy), each behind its own event: a blob's SHA-256 is sequential over its bytes, so the hash
    // of slice k runs while slice k+1 is on the wire and ends one slice after the upload instead of one whole hash after it.
    cudaEvent_t ev_slice[KZG_SLOTS][KZG_HASH_SLICES] = {};
    int slices_recorded[KZG_SLOTS] = {0, 0, 0};  // slices the upload of the slot was cut into (0: one plain copy)
    int cur_slices = 0;                          // of the chunk being enqueued
    cudaEvent_t *cur_slice_ev = nullptr;
    host_g2_prepared *tau_prepared = nullptr;  // Miller-loop lines of [tau]G2
    kzg::g1_affine_t *d_sums_all = nullptr; // sums of a whole device-resident call, [bit position][blob] (grow-only)
    size_t sums_all_elems = 0;
    kzg::fr_t *d_z_all = nullptr;          // challenges of a whole device-resident call (grow-only)
    size_t z_all_elems = 0;
    uint8_t *h_pin = nullptr;         // pinned host buffer of a device-resident verification call (grow-only)
    size_t h_pin_bytes = 0;
    std::vector<cudaEvent_t> ev_chunks;  // "this chunk's (z, y) records are on the host", one per chunk of such a call
    uint8_t *d_vb = nullptr;          // buffers of one verification call (grow-only)
    size_t vb_bytes = 0;
    // what the last kzg_b200_verify_phase_a validated and left decoded in d_vb: SHA-256 over (n, commitments, proofs).
    // kzg_b200_verify_phase_b on the same bytes reuses the decoded points instead of validating them again.
    bool va_valid = false;
    size_t va_n = 0;
    uint8_t va_digest[32];
    cudaStream_t stream = nullptr;
    size_t call_blobs = 0;            // blobs of the call being enqueued (set by the entry points, under `mu`)
    uint64_t launches = 0;
    // optional per-stage device timing (CUDA events on `stream`), see kzg_b200_profile_*
    bool profile = false;
    struct StageRec { int stage; cudaEvent_t a, b; };
    std::vector<StageRec> pending;
    bool stage_open = false;
    double stage_ms[KZG_B200_NUM_STAGES] = {0};
    uint64_t stage_launches[KZG_B200_NUM_STAGES] = {0};
    std::mutex mu;
};

int env_int(const char *name, int dflt);
void stage_begin(kzg_b200_ctx *ctx, int stage);
void stage_end(kzg_b200_ctx *ctx, uint64_t launches);
static inline unsigned blocks_for(uint64_t total, unsigned tpb) { return (unsigned)((total + tpb - 1) / tpb); }

// ---- msm.cu
// decode + bit-reverse the setup points into ctx->d_bases and build ctx->d_table (both allocated by the caller)
int msm_build_table(kzg_b200_ctx *ctx, const uint8_t *g1_bytes);
// scalars of `count` blobs -> comb digits in the current lane: from blob bytes (with the canonical check, status[b]
// = BAD_ARGS for a blob with an element >= r) or from canonical limbs scalars[b*n + i]
// signs_only: stop after the sign words (the latency comb of small calls builds its own digits from them)
int msm_digits_from_blobs(kzg_b200_ctx *ctx, const uint8_t *d_blobs, size_t count, int32_t *d_status, bool signs_only = false);
int msm_digits_from_scalars(kzg_b200_ctx *ctx, const kzg::fr_t *d_scalars, size_t count, bool signs_only = false);
// the 255 sums S_j of `count` blobs from the digits of the current lane: (*out)[j*count + b] (lazy residues)
// out_jac != nullptr: chunks of up to KZG_JAC_TAIL_MAX blobs may end with the Jacobian tail (k_tail_rows_jac, msm.cu): then
// *out is null and *out_jac holds the sums as Jacobian points
#define KZG_JAC_TAIL_MAX 256
#define KZG_JAC_TAIL_ROWS 12
int msm_run(kzg_b200_ctx *ctx, size_t count, const kzg::g1_affine_t **out, const kzg::g1_jac_t **out_jac = nullptr);
// the same sums for a small batch, one warp per sum, as Jacobian points (no inversions on the way): see k_comb_rows_warp
// default threshold: above ~16 blobs the sums no longer fit one wave of warps (255 per blob, eight resident per SM) and the
// batched affine levels win (64 blobs: 2.97 ms against 1.79 ms; 1 blob: 0.16 against 0.85 ms); KZG_B200_MSM_SMALL_MAX moves
// it up to the cap
#define KZG_MSM_SMALL_MAX 16
#define KZG_MSM_SMALL_CAP 64
int msm_run_small(kzg_b200_ctx *ctx, size_t count, const kzg::g1_jac_t **out);
bool msm_small_fits(const kzg_b200_ctx *ctx, size_t count);
// the latency comb: KZG_LAT_T virtual points per setup point, groups of KZG_LAT_G, KZG_LAT_ROWS sums per blob
#define KZG_LAT_G 16
#define KZG_LAT_T 4
#define KZG_LAT_ROWS 64
size_t msm_latency_table_bytes(const kzg_b200_ctx *ctx);
// builds ctx->d_table_lat / d_bases_lat (both allocated by the caller) from ctx->d_bases; needs the lane workspace
int msm_build_latency_table(kzg_b200_ctx *ctx);
// the KZG_LAT_ROWS sums of every blob from the sign words of the current lane, Jacobian, (*out)[j*count + b]
int msm_run_latency(kzg_b200_ctx *ctx, size_t count, const kzg::g1_jac_t **out);
bool msm_latency_ok(const kzg_b200_ctx *ctx, size_t count);
// bytes of lane workspace the MSM needs per blob of a chunk
size_t msm_workspace_per_blob(const kzg_b200_ctx *ctx);
int msm_alloc_lane(kzg_b200_ctx *ctx, kzg_b200_ctx::Lane &ln, size_t chunk);
void msm_free_lane(kzg_b200_ctx::Lane &ln);

// ---- g1ops.cu
// status slot of point i is i % status_mod
int g1_launch_decode(cudaStream_t st, const uint8_t *d_in, kzg::g1_affine_t *d_out, int32_t *d_status, size_t count,
                     int check_subgroup, size_t status_mod);
// `count` commitments and `count` proofs in one launch; a bad commitment or proof of blob i marks status[i]
int g1_launch_decode2(cudaStream_t st, const uint8_t *d_commitments, const uint8_t *d_proofs, kzg::g1_affine_t *d_cpts,
                      kzg::g1_affine_t *d_ppts, int32_t *d_status, size_t count, int check_subgroup);
// the subgroup check of points already decompressed (a failure marks status[i] of commitment i / proof i)
int g1_launch_subgroup2(cudaStream_t st, const kzg::g1_affine_t *d_cpts, const kzg::g1_affine_t *d_ppts, int32_t *d_status, size_t count);
// sums[j*stride + i], j < W  ->  48-byte compressed sum_j 2^j sums[j] of blob i (zeros where status[i] != 0)
int g1_launch_horner_compress(cudaStream_t st, const kzg::g1_affine_t *d_sums, size_t stride, int W, const int32_t *d_status,
                              uint8_t *d_out, size_t count);
// the same from Jacobian sums (msm_run_small), small batches
int g1_launch_horner_compress_jac(cudaStream_t st, const kzg::g1_jac_t *d_sums, size_t stride, int W, const int32_t *d_status,
                                  uint8_t *d_out, size_t count);
int g1_measure_peaks(kzg_b200_ctx *ctx, double *imad_per_s, double *imad_wide_per_s, double *fp_mul_per_s);
int g1_debug_field_op(kzg_b200_ctx *ctx, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, uint64_t count);
// C_i == [p_i(tau)] G1 for every i (test aid, see kzg_b200_debug_check_tau_identity)
int g1_launch_tau_identity(cudaStream_t st, const uint8_t *d_commitments, const uint8_t *d_zy, size_t count, int32_t *d_ok);

// ---- frops.cu
int fr_setup_roots_device(int n, kzg::fr_t **d_roots, cudaStream_t stream);
// lanes per blob the hash of `count` blobs will use (1 = one thread per blob: the batch is too large for hash and validation
// to have a scheduler per warp)
int fr_challenge_form(size_t count, size_t call_blobs, int sms);
// call_blobs: blobs of the whole call this launch is a chunk of (picks the form of the hash kernel)
int fr_launch_challenge(cudaStream_t st, const uint8_t *d_blobs, const uint8_t *d_commitments, size_t count, int n, kzg::fr_t *d_z,
                        size_t call_blobs, int sms);
// blocks [blk_begin, blk_end) of the hash only, state carried in d_state (count x 8 words); g = fr_challenge_form(..) >= 2
int fr_launch_challenge_range(cudaStream_t st, const uint8_t *d_blobs, const uint8_t *d_commitments, size_t count, int n, kzg::fr_t *d_z,
                              int g, uint32_t blk_begin, uint32_t blk_end, uint32_t *d_state);
int fr_launch_load_scalars(cudaStream_t st, const uint8_t *d_in, size_t count, kzg::fr_t *d_out, int32_t *d_status);
// y_i = p_i(z_i) (+ the quotient (p_i(X) - y_i)/(X - z_i) as canonical scalars in d_inv when quotient != 0)
int fr_launch_eval(cudaStream_t st, int quotient, const uint8_t *d_blobs, const kzg::fr_t *d_z, const kzg::fr_t *d_roots, int n,
                   kzg::fr_t *d_inv, kzg::fr_t *d_poly, uint8_t *d_zy, int32_t *d_status, size_t count);
int fr_launch_verify_terms(cudaStream_t st, const kzg::g1_affine_t *d_cpts, const kzg::g1_affine_t *d_ppts, const uint8_t *d_zy,
                           const kzg::fr_t &r_canon, uint64_t first, size_t count, kzg::g1_jac_t *d_terms, kzg::fr_t *d_sy);
// the same three sums by the bucket method (pippenger.cuh), down to the 224-byte partial record: 9 kernels
#define KZG_PIP_LAUNCHES 9
size_t fr_pip_workspace_bytes(size_t count);
int fr_launch_verify_pippenger(cudaStream_t st, const kzg::g1_affine_t *d_cpts, const kzg::g1_affine_t *d_ppts, const uint8_t *d_zy,
                               const kzg::fr_t &r_canon, uint64_t first, size_t count, int force_c, uint8_t *d_ws, kzg::fr_t *d_sy,
                               kzg::g1_affine_t *d_sums, kzg::fr_t *d_sy_total, uint8_t *d_partial);
#define KZG_VERIFY_SUM_BLOCKS 148
int fr_launch_verify_sums(cudaStream_t st, const kzg::g1_jac_t *d_terms, const kzg::fr_t *d_sy, size_t count,
                          kzg::g1_affine_t *d_sums, kzg::fr_t *d_sy_total, uint8_t *d_partial, kzg::g1_jac_t *d_partials);

/*
 * kzg_b200.h -- C ABI of the B200 blob path (libkzg_b200.so).
 *
 * Drop-in boundary for the data-parallel part of pawanjay176/kzg_rust.  The reference has no
 * FFI seam of its own for this path (its only foreign calls are into blst); the entry points
 * below are what `impl Kzg` (reference src/kzg.rs:983-1079) binds to once its bodies are
 * re-routed, plus batched forms.  Each declaration cites the reference item it replaces.
 * INTEGRATION.md shows the Rust side (`extern "C"` block + the `Kzg` methods calling it).
 *
 * Conventions
 *   - All buffers are caller-owned, contiguous, fixed-size records: blobs n x BYTES_PER_BLOB,
 *     commitments / proofs n x 48, field elements n x 32 (big-endian, as in the reference).
 *     Byte-length validation stays in the host wrapper types (reference `Blob::from_bytes`
 *     src/kzg.rs:160-173, `Bytes48::from_bytes` :130-141, `Bytes32::from_bytes` :107-117).
 *   - Return value: KZG_B200_OK or an error code that mirrors reference `enum Error`
 *     (src/kzg.rs:10-22) plus one CUDA/runtime code.  Batched calls also fill status[i] per
 *     blob, so one bad blob does not poison the batch; the single-blob Rust methods map a
 *     non-zero status[0] to `Err`.
 *   - A context plays the role of `KzgSettings` (src/kzg.rs:27-40) and owns all device memory
 *     (the precomputed signed sums of the Lagrange points, roots of unity, workspaces).  It is
 *     bound to one GPU; multi-GPU use is one context (and one process) per GPU with the blob
 *     range sharded by the caller.  Calls on one context are serialised internally, so it may
 *     be shared between host threads like `&KzgSettings`.
 *   - There is no CPU fallback: without a usable CUDA device ctx creation fails.
 */
#ifndef KZG_B200_H
#define KZG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/consts.rs:5-37 */
#define KZG_B200_BYTES_PER_FIELD_ELEMENT 32
#define KZG_B200_BYTES_PER_COMMITMENT 48
#define KZG_B200_BYTES_PER_PROOF 48
#define KZG_B200_BYTES_PER_G1 48
#define KZG_B200_BYTES_PER_G2 96
#define KZG_B200_NUM_G2_POINTS 65

/* reference `enum Error`, src/kzg.rs:10-22 (+ KZG_B200_CUDA_ERROR) */
enum {
    KZG_B200_OK = 0,
    KZG_B200_BAD_ARGS = 1,
    KZG_B200_INTERNAL_ERROR = 2,
    KZG_B200_INVALID_BYTES_LENGTH = 3,
    KZG_B200_INVALID_HEX_FORMAT = 4,
    KZG_B200_INVALID_TRUSTED_SETUP = 5,
    KZG_B200_CUDA_ERROR = 6
};

typedef struct kzg_b200_ctx kzg_b200_ctx;

/*
 * Replaces `Kzg::load_trusted_setup` (src/kzg.rs:1005-1010 -> load_trusted_setup :833-899).
 *   g1_lagrange: n1 x 48 B compressed G1 points, Lagrange form, FILE order (the library
 *                applies the bit-reversal permutation of src/kzg.rs:895-896 itself)
 *   g2_monomial: n2 x 96 B compressed G2 points; n2 must be 65
 *   n1:          FIELD_ELEMENTS_PER_BLOB of the preset: 4096 (kzg_mainnet) or 4 (kzg_minimal)
 *   device:      CUDA device ordinal
 *   comb_width:  g, the number of setup points whose signed sums one table group holds (ceil(n1/g) groups of
 *                2^(g-1) affine points of 96 B); a commitment costs 255 x (ceil(n1/g) - 1) point additions.
 *                1..24; 0 = the widest comb whose table takes at most half of the device's free memory and
 *                leaves room for the workspace (23 on an empty 180 GB B200: 72 GB of table; 22: 38 GB, 4.5 % more
 *                additions; 20: 10 GB, 15 % more)
 * A mainnet context also builds a second, 3.2 GB table for calls of up to 16 blobs (64 sums and 63 doublings per
 * commitment instead of 255 and 254: a single blob_to_kzg_commitment in 0.75 ms) when it fits beside the first one, the
 * workspace and the reserve; KZG_B200_LATENCY_TABLE=0 leaves it out.
 */
int kzg_b200_ctx_create(const uint8_t *g1_lagrange, size_t n1, const uint8_t *g2_monomial, size_t n2,
                        int device, int comb_width, kzg_b200_ctx **out);

/* Replaces `Kzg::load_trusted_setup_file` (src/kzg.rs:995-999 -> :906-979): text file
 * "n1\nn2\n" followed by n1 + n2 hex lines. */
int kzg_b200_ctx_create_from_file(const char *path, int device, int comb_width, kzg_b200_ctx **out);

void kzg_b200_ctx_destroy(kzg_b200_ctx *ctx);

/* FIELD_ELEMENTS_PER_BLOB of the context's preset (src/consts.rs:13). */
size_t kzg_b200_field_elements_per_blob(const kzg_b200_ctx *ctx);
/* Comb width g the context's table was built with, the table's size in bytes, and the number of blobs the
 * context works on at a time (batched calls of any size are cut into chunks of this many). */
int kzg_b200_comb_width(const kzg_b200_ctx *ctx);
size_t kzg_b200_table_bytes(const kzg_b200_ctx *ctx);
size_t kzg_b200_chunk_blobs(const kzg_b200_ctx *ctx);

/*
 * Replaces `Kzg::blob_to_kzg_commitment` (src/kzg.rs:1013-1018 -> :401-406), batched.
 * HOST pointers.  out: n x 48 B.  status: n x int32 (KZG_B200_OK / KZG_B200_BAD_ARGS for a
 * non-canonical field element, src/utils.rs:262-275).
 */
int kzg_b200_blob_to_kzg_commitment_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out,
                                          int32_t *status);

/*
 * Replaces `Kzg::compute_blob_kzg_proof` (src/kzg.rs:1030-1036 -> :533-544), batched.
 * status[i] = KZG_B200_BAD_ARGS for a non-canonical blob element or an invalid commitment
 * (bad encoding, not on the curve, not in G1; src/utils.rs:282-315).
 */
int kzg_b200_compute_blob_kzg_proof_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments,
                                          size_t n, uint8_t *proofs_out, int32_t *status);

/*
 * Replaces `Kzg::compute_kzg_proof` (src/kzg.rs:1021-1027 -> :446-457), batched: proof of
 * evaluation at the caller's z (n x 32 B big-endian, must be canonical).  y_out: n x 32 B.
 */
int kzg_b200_compute_kzg_proof_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *z, size_t n,
                                     uint8_t *proofs_out, uint8_t *y_out, int32_t *status);

/*
 * Replaces `Kzg::verify_blob_kzg_proof_batch` (src/kzg.rs:1066-1078 -> :637-693) and, with
 * n == 1, `Kzg::verify_blob_kzg_proof` (:1050-1063 -> :547-569).  n == 0 -> *ok = 1.
 * Returns KZG_B200_BAD_ARGS (and leaves *ok = 0) if any blob, commitment or proof is
 * malformed, like the `?` short-circuit at src/kzg.rs:671-683.
 */
int kzg_b200_verify_blob_kzg_proof_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments,
                                         const uint8_t *proofs, size_t n, int *ok);

/* reference `Kzg::verify_kzg_proof` (src/kzg.rs:1039-1047 -> :429-445, :409-426): the proof that the
 * polynomial behind `commitment` takes the value y at z.  BAD_ARGS for a z or y that is not a canonical
 * field element, or a commitment / proof that is not a point of G1 (infinity is accepted). */
int kzg_b200_verify_kzg_proof(kzg_b200_ctx *ctx, const uint8_t commitment[48], const uint8_t z[32],
                              const uint8_t y[32], const uint8_t proof[48], int *ok);

/*
 * Two-phase form of the batch verification for multi-GPU sharding.
 * Phase A (per shard): validate, z_i and y_i.  zy_out: n x 64 B (z_i || y_i, big-endian).
 * The caller concatenates the shards' records in blob order, computes r with
 * kzg_b200_compute_r (reference `compute_r_powers`, src/utils.rs:426-474), then
 * Phase B (per shard): partial sums with r^(first_index + i):
 *   out_partial = A_g (96 B uncompressed affine x||y or 0x40.. for infinity) || B_g (96 B) || s_g (32 B)
 *   with A_g = sum r^i proof_i, B_g = sum r^i C_i + sum r^i z_i proof_i, s_g = sum r^i y_i.
 * kzg_b200_verify_finish adds the partial sums of all shards and runs the final check
 * e(A, [tau]G2) == e(B - [s]G1, G2) on the host (reference src/kzg.rs:618-625).
 * Phase B trusts nothing it is handed: commitments and proofs are decompressed and subgroup-checked again, z_i and
 * y_i must be canonical field elements (KZG_B200_BAD_ARGS otherwise), so it is safe on re-fetched data or on another
 * context than phase A's.
 */
int kzg_b200_verify_phase_a(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments,
                            const uint8_t *proofs, size_t n, uint8_t *zy_out);
int kzg_b200_compute_r(const kzg_b200_ctx *ctx, const uint8_t *commitments, const uint8_t *zy,
                       const uint8_t *proofs, size_t n_total, uint8_t r_out[32]);
int kzg_b200_verify_phase_b(kzg_b200_ctx *ctx, const uint8_t *commitments, const uint8_t *zy,
                            const uint8_t *proofs, size_t n, const uint8_t r[32], uint64_t first_index,
                            uint8_t partial_out[224]);
int kzg_b200_verify_finish(const kzg_b200_ctx *ctx, const uint8_t *partials, size_t n_partials, int *ok);

/*
 * Device-resident forms (inputs and outputs already in this context's GPU memory; used by
 * pipelines that produce blobs on the device and by bench.py's HBM-resident measurement).
 * Asynchronous on the context's stream; kzg_b200_synchronize waits for completion.
 * The device pointers must be 16-byte aligned (they are read with 128-bit loads): KZG_B200_BAD_ARGS otherwise.
 */
int kzg_b200_blob_to_kzg_commitment_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, size_t n, uint8_t *d_out,
                                           int32_t *d_status);
int kzg_b200_compute_blob_kzg_proof_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments,
                                           size_t n, uint8_t *d_proofs_out, int32_t *d_status);
int kzg_b200_synchronize(kzg_b200_ctx *ctx);
/* `verify_blob_kzg_proof_batch` over device-resident inputs.  Synchronous: 160 bytes per blob return to the host for
 * the sequential hash of compute_r_powers and the final pairing check runs there.  All validations and all
 * Fiat-Shamir challenges of the call run in one launch each. */
int kzg_b200_verify_blob_kzg_proof_batch_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments,
                                                const uint8_t *d_proofs, size_t n, int *ok);
/* Phase A of the two-phase form (kzg_b200_verify_phase_a above) for a shard that is already in this context's GPU memory:
 * zy_out (host, n x 64 B) as there; commitments_out / proofs_out (host, n x 48 B each, optional) receive the shard's
 * compressed points, which the caller needs on the host for the exchange and for kzg_b200_compute_r.  Synchronous.
 * kzg_b200_verify_phase_b on the same bytes reuses the points this call validated. */
int kzg_b200_verify_phase_a_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments,
                                   const uint8_t *d_proofs, size_t n, uint8_t *zy_out, uint8_t *commitments_out,
                                   uint8_t *proofs_out);
/*
 * Per-stage device timing of the calls on this context, measured with CUDA events on the
 * context's stream (bench.py's roofline numbers come from here).  enable(1) resets the
 * accumulators; read() waits for the stream and returns milliseconds / kernel launches per
 * stage since then (arrays of KZG_B200_NUM_STAGES).
 */
enum {
    KZG_B200_STAGE_DIGITS = 0,       /* blob bytes -> canonical check -> sign words -> comb digits */
    KZG_B200_STAGE_MSM_GATHER = 1,   /* first level of the MSM: table gather + batched affine additions */
    KZG_B200_STAGE_MSM_TREE = 2,     /* remaining levels of the addition tree */
    KZG_B200_STAGE_COMPRESS = 3,     /* 48-byte compression */
    KZG_B200_STAGE_CHALLENGE = 4,    /* per-blob SHA-256 Fiat-Shamir challenge */
    KZG_B200_STAGE_EVAL = 5,         /* barycentric evaluation (+ quotient, digits) */
    KZG_B200_STAGE_VALIDATE = 6,     /* G1 decompression + subgroup checks */
    KZG_B200_STAGE_VERIFY_TERMS = 7, /* r-power linear combinations of batch verification (bucket-method MSM) */
    KZG_B200_NUM_STAGES = 8
};
int kzg_b200_profile_enable(kzg_b200_ctx *ctx, int on);
int kzg_b200_profile_read(kzg_b200_ctx *ctx, double *ms_out, uint64_t *launches_out);
/* the CUDA stream (cudaStream_t) the context launches on, for event timing */
void *kzg_b200_stream(kzg_b200_ctx *ctx);
/* number of kernel launches issued by this context so far (bench.py reports the delta) */
uint64_t kzg_b200_launch_count(const kzg_b200_ctx *ctx);

/* Host final check used by verify (stand-in for blst's pairing, reference src/utils.rs:189-214):
 * e(a1, a2) == e(b1, b2) on compressed points.  Exposed for tests. */
int kzg_b200_pairings_verify(const uint8_t a1[48], const uint8_t a2[96], const uint8_t b1[48],
                             const uint8_t b2[96], int *ok);

/* Micro-benchmarks used by bench.py to measure the integer-multiply roofline on the device
 * the context is bound to: issue rate of independent 32-bit multiply-adds (mad.lo.u32 ->
 * IMAD, ops/s), of independent 32x32+64 multiply-adds (mad.wide.u32 -> IMAD.WIDE, ops/s)
 * and of dependent full Fp Montgomery multiplications (mul/s).  */
int kzg_b200_measure_peaks(kzg_b200_ctx *ctx, double *imad_per_s, double *imad_wide_per_s, double *fp_mul_per_s);

/* Test aids (tests/test_gpu_field.py, tests/test_gpu_commit.py): device field arithmetic on arrays of raw
 * limbs -- op 0: Fp mul, 1: Fp inverse, 2: Fr mul (8 words per element, all others 12), 3: Fp add, 4: Fp sub,
 * 7 / 8: Fp mul / sub on lazy residues in [0, 2p) (what the MSM levels use) -- and a read-back of precomputed
 * table entries (flat index: group q, entry idx -> q * 2^(g-1) + idx; 96 B each, affine Montgomery limbs). */
int kzg_b200_debug_field_op(kzg_b200_ctx *ctx, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, uint64_t count);
int kzg_b200_debug_table(kzg_b200_ctx *ctx, uint64_t first, uint64_t count, void *out);
/*
 * Whole-batch identity check for setups whose secret is public (the bundled testing setup: tau = 1337, SURVEY.md
 * section 8c): d_ok[i] = (commitment_i == [p_i(tau)] G1), where p_i(tau) comes from the evaluation kernel and
 * [y]G1 from a plain double-and-add ladder on the generator -- no part of the MSM is on that side.  Device
 * pointers (16-byte aligned), n x int32 out; synchronous.  bench.py checks every commitment of the timed batch with it.
 */
int kzg_b200_debug_check_tau_identity(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments, size_t n,
                                      const uint8_t tau[32], int32_t *d_ok);

#ifdef __cplusplus
}
#endif
#endif

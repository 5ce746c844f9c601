//! `extern "C"` declarations of libkzg_b200.so -- GENERATED from include/kzg_b200.h by tools/gen_ffi_rs.py,
//! do not edit (tests/test_ffi_matches_header.py compares this file with the header).
#![allow(dead_code)]
use std::os::raw::{c_char, c_int, c_void};

/// Opaque context (`kzg_b200_ctx`): plays the role of the reference's `KzgSettings` and owns all device memory.
#[repr(C)]
pub struct KzgB200Ctx {
    _private: [u8; 0],
}

pub const KZG_B200_BYTES_PER_FIELD_ELEMENT: usize = 32;
pub const KZG_B200_BYTES_PER_COMMITMENT: usize = 48;
pub const KZG_B200_BYTES_PER_PROOF: usize = 48;
pub const KZG_B200_BYTES_PER_G1: usize = 48;
pub const KZG_B200_BYTES_PER_G2: usize = 96;
pub const KZG_B200_NUM_G2_POINTS: usize = 65;
pub const KZG_B200_OK: c_int = 0;
pub const KZG_B200_BAD_ARGS: c_int = 1;
pub const KZG_B200_INTERNAL_ERROR: c_int = 2;
pub const KZG_B200_INVALID_BYTES_LENGTH: c_int = 3;
pub const KZG_B200_INVALID_HEX_FORMAT: c_int = 4;
pub const KZG_B200_INVALID_TRUSTED_SETUP: c_int = 5;
pub const KZG_B200_CUDA_ERROR: c_int = 6;
pub const KZG_B200_STAGE_DIGITS: c_int = 0;
pub const KZG_B200_STAGE_MSM_GATHER: c_int = 1;
pub const KZG_B200_STAGE_MSM_TREE: c_int = 2;
pub const KZG_B200_STAGE_COMPRESS: c_int = 3;
pub const KZG_B200_STAGE_CHALLENGE: c_int = 4;
pub const KZG_B200_STAGE_EVAL: c_int = 5;
pub const KZG_B200_STAGE_VALIDATE: c_int = 6;
pub const KZG_B200_STAGE_VERIFY_TERMS: c_int = 7;
pub const KZG_B200_NUM_STAGES: usize = 8;

#[link(name = "kzg_b200")]
extern "C" {
    pub fn kzg_b200_ctx_create(g1_lagrange: *const u8, n1: usize, g2_monomial: *const u8, n2: usize, device: c_int, comb_width: c_int, out: *mut *mut KzgB200Ctx) -> c_int;
    pub fn kzg_b200_ctx_create_from_file(path: *const c_char, device: c_int, comb_width: c_int, out: *mut *mut KzgB200Ctx) -> c_int;
    pub fn kzg_b200_ctx_destroy(ctx: *mut KzgB200Ctx);
    pub fn kzg_b200_field_elements_per_blob(ctx: *const KzgB200Ctx) -> usize;
    pub fn kzg_b200_comb_width(ctx: *const KzgB200Ctx) -> c_int;
    pub fn kzg_b200_table_bytes(ctx: *const KzgB200Ctx) -> usize;
    pub fn kzg_b200_chunk_blobs(ctx: *const KzgB200Ctx) -> usize;
    pub fn kzg_b200_blob_to_kzg_commitment_batch(ctx: *mut KzgB200Ctx, blobs: *const u8, n: usize, out: *mut u8, status: *mut i32) -> c_int;
    pub fn kzg_b200_compute_blob_kzg_proof_batch(ctx: *mut KzgB200Ctx, blobs: *const u8, commitments: *const u8, n: usize, proofs_out: *mut u8, status: *mut i32) -> c_int;
    pub fn kzg_b200_compute_kzg_proof_batch(ctx: *mut KzgB200Ctx, blobs: *const u8, z: *const u8, n: usize, proofs_out: *mut u8, y_out: *mut u8, status: *mut i32) -> c_int;
    pub fn kzg_b200_verify_blob_kzg_proof_batch(ctx: *mut KzgB200Ctx, blobs: *const u8, commitments: *const u8, proofs: *const u8, n: usize, ok: *mut c_int) -> c_int;
    pub fn kzg_b200_verify_kzg_proof(ctx: *mut KzgB200Ctx, commitment: *const u8, z: *const u8, y: *const u8, proof: *const u8, ok: *mut c_int) -> c_int;
    pub fn kzg_b200_verify_phase_a(ctx: *mut KzgB200Ctx, blobs: *const u8, commitments: *const u8, proofs: *const u8, n: usize, zy_out: *mut u8) -> c_int;
    pub fn kzg_b200_compute_r(ctx: *const KzgB200Ctx, commitments: *const u8, zy: *const u8, proofs: *const u8, n_total: usize, r_out: *mut u8) -> c_int;
    pub fn kzg_b200_verify_phase_b(ctx: *mut KzgB200Ctx, commitments: *const u8, zy: *const u8, proofs: *const u8, n: usize, r: *const u8, first_index: u64, partial_out: *mut u8) -> c_int;
    pub fn kzg_b200_verify_finish(ctx: *const KzgB200Ctx, partials: *const u8, n_partials: usize, ok: *mut c_int) -> c_int;
    pub fn kzg_b200_blob_to_kzg_commitment_device(ctx: *mut KzgB200Ctx, d_blobs: *const u8, n: usize, d_out: *mut u8, d_status: *mut i32) -> c_int;
    pub fn kzg_b200_compute_blob_kzg_proof_device(ctx: *mut KzgB200Ctx, d_blobs: *const u8, d_commitments: *const u8, n: usize, d_proofs_out: *mut u8, d_status: *mut i32) -> c_int;
    pub fn kzg_b200_synchronize(ctx: *mut KzgB200Ctx) -> c_int;
    pub fn kzg_b200_verify_blob_kzg_proof_batch_device(ctx: *mut KzgB200Ctx, d_blobs: *const u8, d_commitments: *const u8, d_proofs: *const u8, n: usize, ok: *mut c_int) -> c_int;
    pub fn kzg_b200_verify_phase_a_device(ctx: *mut KzgB200Ctx, d_blobs: *const u8, d_commitments: *const u8, d_proofs: *const u8, n: usize, zy_out: *mut u8, commitments_out: *mut u8, proofs_out: *mut u8) -> c_int;
    pub fn kzg_b200_profile_enable(ctx: *mut KzgB200Ctx, on: c_int) -> c_int;
    pub fn kzg_b200_profile_read(ctx: *mut KzgB200Ctx, ms_out: *mut f64, launches_out: *mut u64) -> c_int;
    pub fn kzg_b200_stream(ctx: *mut KzgB200Ctx) -> *mut c_void;
    pub fn kzg_b200_launch_count(ctx: *const KzgB200Ctx) -> u64;
    pub fn kzg_b200_pairings_verify(a1: *const u8, a2: *const u8, b1: *const u8, b2: *const u8, ok: *mut c_int) -> c_int;
    pub fn kzg_b200_measure_peaks(ctx: *mut KzgB200Ctx, imad_per_s: *mut f64, imad_wide_per_s: *mut f64, fp_mul_per_s: *mut f64) -> c_int;
    pub fn kzg_b200_debug_field_op(ctx: *mut KzgB200Ctx, op: c_int, a: *const u32, b: *const u32, out: *mut u32, count: u64) -> c_int;
    pub fn kzg_b200_debug_table(ctx: *mut KzgB200Ctx, first: u64, count: u64, out: *mut c_void) -> c_int;
    pub fn kzg_b200_debug_check_tau_identity(ctx: *mut KzgB200Ctx, d_blobs: *const u8, d_commitments: *const u8, n: usize, tau: *const u8, d_ok: *mut i32) -> c_int;
}

//! `kzg_rust` with its data-parallel blob path on a B200.
//!
//! Same public surface as the reference crate (reference `src/lib.rs:7-12`, `README.md:8-32`): the byte newtypes,
//! `Error`, and per preset a `Kzg` struct with the eight methods of `impl Kzg` (reference `src/kzg.rs:983-1079`),
//! `Blob`, `KzgSettings` and `TrustedSetup`.  The bodies no longer call blst: every method forwards to the C ABI of
//! `include/kzg_b200.h` (`ffi.rs`, generated from that header), where the work runs as hand-written sm_100a CUDA.
//! New here are the `*_batch` methods, which hand whole slices of blobs to the GPU in one call.
//!
//! The presets are the two modules the reference's README names: `kzg_mainnet` (4096 field elements per blob) and
//! `kzg_minimal` (4).  The crate root re-exports the mainnet items, as the reference does.
mod bytes;
mod ffi;
#[macro_use]
mod preset;

pub use bytes::{
    hex_to_bytes, Bytes32, Bytes48, Error, KzgCommitment, KzgProof, BYTES_PER_COMMITMENT, BYTES_PER_FIELD_ELEMENT,
    BYTES_PER_G1, BYTES_PER_G2, BYTES_PER_PROOF, TRUSTED_SETUP_NUM_G2_POINTS,
};

preset_module!(
    /// Mainnet preset: `FIELD_ELEMENTS_PER_BLOB = 4096` (reference `src/consts.rs:13`).
    kzg_mainnet,
    4096
);
preset_module!(
    /// Minimal preset: `FIELD_ELEMENTS_PER_BLOB = 4`.
    kzg_minimal,
    4
);

pub use kzg_mainnet::{Blob, Kzg, KzgSettings, TrustedSetup, BYTES_PER_BLOB, FIELD_ELEMENTS_PER_BLOB};

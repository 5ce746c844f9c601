//! The reference's six vector suites (reference `src/lib.rs:30-203`, formats `src/test_formats/*.rs`) against the
//! GPU-backed crate.  Run from a checkout of the reference's `tests/` tree and `trusted_setup.txt`:
//!
//!     KZG_VECTORS=/path/to/kzg_rust cargo test --release
//!
//! (default: the directory above this crate).  Every case is `{input: {...}, output: ...}` in `data.yaml`; a case
//! whose inputs do not parse, or whose call fails, must have `output: null`.
use std::fs;
use std::path::PathBuf;

use kzg_rust::{Blob, Bytes32, Bytes48, Error, Kzg, KzgCommitment, KzgProof, KzgSettings};
use serde::Deserialize;

fn root() -> PathBuf {
    std::env::var("KZG_VECTORS").map(PathBuf::from).unwrap_or_else(|_| PathBuf::from(".."))
}

fn settings() -> KzgSettings {
    Kzg::load_trusted_setup_file(root().join("trusted_setup.txt")).unwrap()
}

fn cases<T: for<'de> Deserialize<'de>>(suite: &str) -> Vec<T> {
    let pattern = root().join("tests").join(suite).join("*/*/data.yaml");
    let files: Vec<PathBuf> = glob::glob(pattern.to_str().unwrap()).unwrap().map(Result::unwrap).collect();
    assert!(!files.is_empty(), "no vectors under {}", pattern.display());
    files
        .into_iter()
        .map(|f| serde_yaml::from_str(&fs::read_to_string(f).unwrap()).unwrap())
        .collect()
}

#[derive(Deserialize)]
struct Case<I, O> {
    input: I,
    output: Option<O>,
}

#[derive(Deserialize)]
struct BlobIn {
    blob: String,
}

#[derive(Deserialize)]
struct BlobZIn {
    blob: String,
    z: String,
}

#[derive(Deserialize)]
struct BlobCommitmentIn {
    blob: String,
    commitment: String,
}

#[derive(Deserialize)]
struct PointIn {
    commitment: String,
    z: String,
    y: String,
    proof: String,
}

#[derive(Deserialize)]
struct BlobProofIn {
    blob: String,
    commitment: String,
    proof: String,
}

#[derive(Deserialize)]
struct BatchIn {
    blobs: Vec<String>,
    commitments: Vec<String>,
    proofs: Vec<String>,
}

fn commitment(hex: &str) -> Result<KzgCommitment, Error> {
    Ok(KzgCommitment(Bytes48::from_hex(hex)?))
}

fn proof(hex: &str) -> Result<KzgProof, Error> {
    Ok(KzgProof(Bytes48::from_hex(hex)?))
}

#[test]
fn test_blob_to_kzg_commitment() {
    let s = settings();
    for t in cases::<Case<BlobIn, String>>("blob_to_kzg_commitment") {
        let got = Blob::from_hex(&t.input.blob).and_then(|b| Kzg::blob_to_kzg_commitment(&b, &s));
        match got {
            Ok(c) => assert_eq!(c.to_bytes(), *Bytes48::from_hex(&t.output.unwrap()).unwrap()),
            Err(_) => assert!(t.output.is_none()),
        }
    }
}

#[test]
fn test_compute_kzg_proof() {
    let s = settings();
    for t in cases::<Case<BlobZIn, (String, String)>>("compute_kzg_proof") {
        let got = Blob::from_hex(&t.input.blob)
            .and_then(|b| Bytes32::from_hex(&t.input.z).map(|z| (b, z)))
            .and_then(|(b, z)| Kzg::compute_kzg_proof(&b, &z, &s));
        match got {
            Ok((p, y)) => {
                let (ep, ey) = t.output.unwrap();
                assert_eq!(p.to_bytes(), *Bytes48::from_hex(&ep).unwrap());
                assert_eq!(*y, *Bytes32::from_hex(&ey).unwrap());
            }
            Err(_) => assert!(t.output.is_none()),
        }
    }
}

#[test]
fn test_compute_blob_kzg_proof() {
    let s = settings();
    for t in cases::<Case<BlobCommitmentIn, String>>("compute_blob_kzg_proof") {
        let got = Blob::from_hex(&t.input.blob)
            .and_then(|b| commitment(&t.input.commitment).map(|c| (b, c)))
            .and_then(|(b, c)| Kzg::compute_blob_kzg_proof(&b, &c, &s));
        match got {
            Ok(p) => assert_eq!(p.to_bytes(), *Bytes48::from_hex(&t.output.unwrap()).unwrap()),
            Err(_) => assert!(t.output.is_none()),
        }
    }
}

#[test]
fn test_verify_kzg_proof() {
    let s = settings();
    for t in cases::<Case<PointIn, bool>>("verify_kzg_proof") {
        let got = (|| {
            let c = commitment(&t.input.commitment)?;
            let z = Bytes32::from_hex(&t.input.z)?;
            let y = Bytes32::from_hex(&t.input.y)?;
            let p = proof(&t.input.proof)?;
            Kzg::verify_kzg_proof(&c, &z, &y, &p, &s)
        })();
        match got {
            Ok(ok) => assert_eq!(Some(ok), t.output),
            Err(_) => assert!(t.output.is_none()),
        }
    }
}

#[test]
fn test_verify_blob_kzg_proof() {
    let s = settings();
    for t in cases::<Case<BlobProofIn, bool>>("verify_blob_kzg_proof") {
        let got = (|| {
            let b = Blob::from_hex(&t.input.blob)?;
            let c = commitment(&t.input.commitment)?;
            let p = proof(&t.input.proof)?;
            Kzg::verify_blob_kzg_proof(&b, &c, &p, &s)
        })();
        match got {
            Ok(ok) => assert_eq!(Some(ok), t.output),
            Err(_) => assert!(t.output.is_none()),
        }
    }
}

#[test]
fn test_verify_blob_kzg_proof_batch() {
    let s = settings();
    for t in cases::<Case<BatchIn, bool>>("verify_blob_kzg_proof_batch") {
        let got = (|| {
            let blobs = t.input.blobs.iter().map(|h| Blob::from_hex(h)).collect::<Result<Vec<_>, _>>()?;
            let cs = t.input.commitments.iter().map(|h| commitment(h)).collect::<Result<Vec<_>, _>>()?;
            let ps = t.input.proofs.iter().map(|h| proof(h)).collect::<Result<Vec<_>, _>>()?;
            Kzg::verify_blob_kzg_proof_batch(&blobs, &cs, &ps, &s)
        })();
        match got {
            Ok(ok) => assert_eq!(Some(ok), t.output),
            Err(_) => assert!(t.output.is_none()),
        }
    }
}

/// The batched entry points agree with the single-blob ones on the valid commitment vectors.
#[test]
fn test_batched_entry_points_match_single_calls() {
    let s = settings();
    let blobs: Vec<Blob> = cases::<Case<BlobIn, String>>("blob_to_kzg_commitment")
        .into_iter()
        .filter(|t| t.output.is_some())
        .map(|t| Blob::from_hex(&t.input.blob).unwrap())
        .collect();
    let commitments = Kzg::blob_to_kzg_commitment_batch(&blobs, &s).unwrap();
    for (b, c) in blobs.iter().zip(commitments.iter()) {
        assert_eq!(*c, Kzg::blob_to_kzg_commitment(b, &s).unwrap());
    }
    let proofs = Kzg::compute_blob_kzg_proof_batch(&blobs, &commitments, &s).unwrap();
    for ((b, c), p) in blobs.iter().zip(commitments.iter()).zip(proofs.iter()) {
        assert_eq!(*p, Kzg::compute_blob_kzg_proof(b, c, &s).unwrap());
    }
    assert!(Kzg::verify_blob_kzg_proof_batch(&blobs, &commitments, &proofs, &s).unwrap());
}

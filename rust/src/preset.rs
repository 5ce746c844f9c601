//! One module per preset (`kzg_mainnet`, `kzg_minimal`): everything of the reference's API that depends on
//! `FIELD_ELEMENTS_PER_BLOB` -- `Blob`, `KzgSettings`, `TrustedSetup`, `Kzg` (reference `README.md:8-32`).
//!
//! The bodies forward to the C ABI (`ffi.rs`).  Single-blob methods are the batched entry points with n = 1 and turn a
//! non-zero `status[0]` into the `Err` the reference returns; the `*_batch` methods are new and return the first
//! error in blob order, like the `?` short-circuit at reference `src/kzg.rs:671-683`.

macro_rules! preset_module {
    ($(#[$doc:meta])* $name:ident, $n:expr) => {
        $(#[$doc])*
        pub mod $name {
            use std::ffi::CString;
            use std::ops::{Deref, DerefMut};
            use std::os::raw::c_int;
            use std::path::Path;
            use std::ptr;

            use serde::de::{self, Deserializer, Visitor};
            use serde::{Deserialize, Serialize};

            use crate::bytes::{check, Bytes32, Error, KzgCommitment, KzgProof};
            use crate::bytes::{BYTES_PER_FIELD_ELEMENT, BYTES_PER_G1, BYTES_PER_G2, TRUSTED_SETUP_NUM_G2_POINTS};
            use crate::ffi;

            /// reference `src/consts.rs:13`.
            pub const FIELD_ELEMENTS_PER_BLOB: usize = $n;
            /// Number of bytes in a blob (reference `src/consts.rs:16`).
            pub const BYTES_PER_BLOB: usize = FIELD_ELEMENTS_PER_BLOB * BYTES_PER_FIELD_ELEMENT;

            /// reference `src/kzg.rs:153-178`.
            #[derive(Debug, Clone, PartialEq)]
            pub struct Blob {
                bytes: Box<[u8; BYTES_PER_BLOB]>,
            }

            impl Blob {
                pub fn from_bytes(bytes: &[u8]) -> Result<Self, Error> {
                    if bytes.len() != BYTES_PER_BLOB {
                        return Err(Error::InvalidBytesLength(format!(
                            "Invalid byte length. Expected {} got {}",
                            BYTES_PER_BLOB,
                            bytes.len(),
                        )));
                    }
                    let mut new_bytes = Box::new([0u8; BYTES_PER_BLOB]);
                    new_bytes.copy_from_slice(bytes);
                    Ok(Self { bytes: new_bytes })
                }

                pub fn from_hex(hex_str: &str) -> Result<Self, Error> {
                    Self::from_bytes(&crate::bytes::hex_to_bytes(hex_str)?)
                }
            }

            impl From<[u8; BYTES_PER_BLOB]> for Blob {
                fn from(value: [u8; BYTES_PER_BLOB]) -> Self {
                    Self { bytes: Box::new(value) }
                }
            }

            impl Deref for Blob {
                type Target = [u8; BYTES_PER_BLOB];
                fn deref(&self) -> &Self::Target {
                    &self.bytes
                }
            }

            impl DerefMut for Blob {
                fn deref_mut(&mut self) -> &mut Self::Target {
                    &mut self.bytes
                }
            }

            /// Stores the setup and parameters needed for computing KZG proofs (reference `src/kzg.rs:27-40`).
            ///
            /// Here it owns a `kzg_b200_ctx`: the comb table of the Lagrange-basis points, the roots of unity and the
            /// workspaces live in the memory of one GPU.  The context serialises calls internally, so `&KzgSettings`
            /// can be shared between threads like the reference's plain-data struct.
            #[derive(Debug)]
            pub struct KzgSettings {
                ctx: *mut ffi::KzgB200Ctx,
            }

            unsafe impl Send for KzgSettings {}
            unsafe impl Sync for KzgSettings {}

            impl Drop for KzgSettings {
                fn drop(&mut self) {
                    unsafe { ffi::kzg_b200_ctx_destroy(self.ctx) }
                }
            }

            impl KzgSettings {
                /// Initializes a trusted setup from `FIELD_ELEMENTS_PER_BLOB` g1 points
                /// and 65 g2 points in byte format (reference `src/kzg.rs:45-79`), on GPU 0 with the automatic
                /// comb width.
                pub fn load_trusted_setup(
                    g1_bytes: Vec<[u8; BYTES_PER_G1]>,
                    g2_bytes: Vec<[u8; BYTES_PER_G2]>,
                ) -> Result<Self, Error> {
                    Self::load_trusted_setup_on(g1_bytes, g2_bytes, 0, 0)
                }

                /// The same with the CUDA device ordinal and the comb width of the table chosen by the caller
                /// (`comb_width` 0 = automatic, see `include/kzg_b200.h`).  Multi-GPU use is one `KzgSettings` per GPU.
                pub fn load_trusted_setup_on(
                    g1_bytes: Vec<[u8; BYTES_PER_G1]>,
                    g2_bytes: Vec<[u8; BYTES_PER_G2]>,
                    device: i32,
                    comb_width: i32,
                ) -> Result<Self, Error> {
                    if g1_bytes.len() != FIELD_ELEMENTS_PER_BLOB {
                        return Err(Error::InvalidTrustedSetup(format!(
                            "Invalid number of g1 points in trusted setup. Expected {} got {}",
                            FIELD_ELEMENTS_PER_BLOB,
                            g1_bytes.len()
                        )));
                    }
                    if g2_bytes.len() != TRUSTED_SETUP_NUM_G2_POINTS {
                        return Err(Error::InvalidTrustedSetup(format!(
                            "Invalid number of g2 points in trusted setup. Expected {} got {}",
                            TRUSTED_SETUP_NUM_G2_POINTS,
                            g2_bytes.len()
                        )));
                    }
                    let g1_points: Vec<u8> = g1_bytes.iter().flat_map(|p| p.iter().copied()).collect();
                    let g2_points: Vec<u8> = g2_bytes.iter().flat_map(|p| p.iter().copied()).collect();
                    let mut ctx: *mut ffi::KzgB200Ctx = ptr::null_mut();
                    let rc = unsafe {
                        ffi::kzg_b200_ctx_create(
                            g1_points.as_ptr(),
                            g1_bytes.len(),
                            g2_points.as_ptr(),
                            g2_bytes.len(),
                            device as c_int,
                            comb_width as c_int,
                            &mut ctx,
                        )
                    };
                    check(rc, "load_trusted_setup")?;
                    Ok(Self { ctx })
                }

                /// reference `load_trusted_setup_file`, `src/kzg.rs:906-979`.
                pub fn load_trusted_setup_file_on<P: AsRef<Path>>(
                    trusted_setup_file: P,
                    device: i32,
                    comb_width: i32,
                ) -> Result<Self, Error> {
                    let path = trusted_setup_file.as_ref().to_str().ok_or_else(|| {
                        Error::InvalidTrustedSetup("trusted setup path is not valid unicode".to_string())
                    })?;
                    let c_path = CString::new(path)
                        .map_err(|_| Error::InvalidTrustedSetup("trusted setup path contains a NUL byte".to_string()))?;
                    let mut ctx: *mut ffi::KzgB200Ctx = ptr::null_mut();
                    let rc = unsafe {
                        ffi::kzg_b200_ctx_create_from_file(c_path.as_ptr(), device as c_int, comb_width as c_int, &mut ctx)
                    };
                    check(rc, "load_trusted_setup_file")?;
                    let s = Self { ctx };
                    if s.field_elements_per_blob() != FIELD_ELEMENTS_PER_BLOB {
                        return Err(Error::InvalidTrustedSetup(format!(
                            "Invalid number of g1 points in trusted setup. Expected {} got {}",
                            FIELD_ELEMENTS_PER_BLOB,
                            s.field_elements_per_blob()
                        )));
                    }
                    Ok(s)
                }

                fn field_elements_per_blob(&self) -> usize {
                    unsafe { ffi::kzg_b200_field_elements_per_blob(self.ctx) }
                }

                /// Comb width of the precomputed table, and its size in bytes of device memory.
                pub fn comb_width(&self) -> usize {
                    unsafe { ffi::kzg_b200_comb_width(self.ctx) as usize }
                }

                pub fn table_bytes(&self) -> usize {
                    unsafe { ffi::kzg_b200_table_bytes(self.ctx) }
                }
            }

            /// Wrapper over a BLS G1 point's byte representation (reference `src/trusted_setup.rs:5-7`).
            #[derive(Debug, Clone, PartialEq)]
            struct G1Point([u8; BYTES_PER_G1]);

            /// Wrapper over a BLS G2 point's byte representation (reference `src/trusted_setup.rs:9-11`).
            #[derive(Debug, Clone, PartialEq)]
            struct G2Point([u8; BYTES_PER_G2]);

            /// Contains the trusted setup parameters that are required to instantiate a `KzgSettings` object, in the
            /// json format of the ethereum consensus specs (reference `src/trusted_setup.rs:13-44`).
            #[derive(Debug, Clone, PartialEq, Serialize, Deserialize)]
            pub struct TrustedSetup {
                #[serde(rename = "setup_G1_lagrange")]
                #[serde(deserialize_with = "deserialize_g1_points")]
                g1_points: Vec<G1Point>,
                #[serde(rename = "setup_G2")]
                g2_points: Vec<G2Point>,
            }

            impl TrustedSetup {
                pub fn g1_points(&self) -> Vec<[u8; BYTES_PER_G1]> {
                    self.g1_points.iter().map(|p| p.0).collect()
                }

                pub fn g2_points(&self) -> Vec<[u8; BYTES_PER_G2]> {
                    self.g2_points.iter().map(|p| p.0).collect()
                }

                pub fn g1_len(&self) -> usize {
                    self.g1_points.len()
                }

                pub fn g2_len(&self) -> usize {
                    self.g2_points.len()
                }
            }

            fn strip_prefix(s: &str) -> &str {
                s.strip_prefix("0x").unwrap_or(s)
            }

            fn decode_point<E: de::Error, const N: usize>(v: &str, what: &str) -> Result<[u8; N], E> {
                let point = hex::decode(strip_prefix(v))
                    .map_err(|e| de::Error::custom(format!("Failed to decode {} point: {}", what, e)))?;
                if point.len() != N {
                    return Err(de::Error::custom(format!(
                        "{} point has invalid length. Expected {} got {}",
                        what,
                        N,
                        point.len()
                    )));
                }
                let mut res = [0u8; N];
                res.copy_from_slice(&point);
                Ok(res)
            }

            impl Serialize for G1Point {
                fn serialize<S: serde::Serializer>(&self, serializer: S) -> Result<S::Ok, S::Error> {
                    serializer.serialize_str(&hex::encode(self.0))
                }
            }

            impl Serialize for G2Point {
                fn serialize<S: serde::Serializer>(&self, serializer: S) -> Result<S::Ok, S::Error> {
                    serializer.serialize_str(&hex::encode(self.0))
                }
            }

            impl<'de> Deserialize<'de> for G1Point {
                fn deserialize<D: Deserializer<'de>>(deserializer: D) -> Result<Self, D::Error> {
                    struct G1PointVisitor;
                    impl<'de> Visitor<'de> for G1PointVisitor {
                        type Value = G1Point;
                        fn expecting(&self, formatter: &mut std::fmt::Formatter) -> std::fmt::Result {
                            formatter.write_str("A 48 byte hex encoded string")
                        }
                        fn visit_str<E: de::Error>(self, v: &str) -> Result<Self::Value, E> {
                            Ok(G1Point(decode_point::<E, BYTES_PER_G1>(v, "G1")?))
                        }
                    }
                    deserializer.deserialize_str(G1PointVisitor)
                }
            }

            impl<'de> Deserialize<'de> for G2Point {
                fn deserialize<D: Deserializer<'de>>(deserializer: D) -> Result<Self, D::Error> {
                    struct G2PointVisitor;
                    impl<'de> Visitor<'de> for G2PointVisitor {
                        type Value = G2Point;
                        fn expecting(&self, formatter: &mut std::fmt::Formatter) -> std::fmt::Result {
                            formatter.write_str("A 96 byte hex encoded string")
                        }
                        fn visit_str<E: de::Error>(self, v: &str) -> Result<Self::Value, E> {
                            Ok(G2Point(decode_point::<E, BYTES_PER_G2>(v, "G2")?))
                        }
                    }
                    deserializer.deserialize_str(G2PointVisitor)
                }
            }

            /// Minimal and mainnet trusted setup parameters differ only by the number of G1 points they contain: the
            /// list is truncated to this preset's `FIELD_ELEMENTS_PER_BLOB` (reference `src/trusted_setup.rs:138-153`).
            fn deserialize_g1_points<'de, D: Deserializer<'de>>(deserializer: D) -> Result<Vec<G1Point>, D::Error> {
                let mut decoded: Vec<G1Point> = Deserialize::deserialize(deserializer)?;
                decoded.truncate(FIELD_ELEMENTS_PER_BLOB);
                Ok(decoded)
            }

            fn first_error(status: &[i32], what: &str) -> Result<(), Error> {
                match status.iter().find(|&&st| st != ffi::KZG_B200_OK) {
                    Some(&st) => Err(Error::from_code(st, what)),
                    None => Ok(()),
                }
            }

            fn flatten_blobs(blobs: &[Blob]) -> Vec<u8> {
                let mut out = Vec::with_capacity(blobs.len() * BYTES_PER_BLOB);
                for b in blobs {
                    out.extend_from_slice(&b.bytes[..]);
                }
                out
            }

            fn flatten48<T: Deref<Target = [u8; 48]>>(items: &[T]) -> Vec<u8> {
                let mut out = Vec::with_capacity(items.len() * 48);
                for x in items {
                    out.extend_from_slice(x.deref());
                }
                out
            }

            fn split48(bytes: &[u8]) -> Vec<[u8; 48]> {
                bytes
                    .chunks_exact(48)
                    .map(|c| {
                        let mut a = [0u8; 48];
                        a.copy_from_slice(c);
                        a
                    })
                    .collect()
            }

            /// A wrapper struct that exposes the interface functions as struct methods (reference
            /// `pub struct Kzg`, `src/kzg.rs:983-1079`).
            pub struct Kzg;

            impl Kzg {
                /// Loads a trusted setup in the text format of `trusted_setup.txt` (reference `src/kzg.rs:995-999`).
                pub fn load_trusted_setup_file<P: AsRef<Path>>(trusted_setup_file: P) -> Result<KzgSettings, Error> {
                    KzgSettings::load_trusted_setup_file_on(trusted_setup_file, 0, 0)
                }

                /// Loads a trusted setup and returns a `KzgSettings` struct (reference `src/kzg.rs:1005-1010`).
                pub fn load_trusted_setup(
                    g1_bytes: Vec<[u8; BYTES_PER_G1]>,
                    g2_bytes: Vec<[u8; BYTES_PER_G2]>,
                ) -> Result<KzgSettings, Error> {
                    KzgSettings::load_trusted_setup(g1_bytes, g2_bytes)
                }

                /// Return the `KzgCommitment` corresponding to the `Blob` (reference `src/kzg.rs:1013-1018`).
                pub fn blob_to_kzg_commitment(blob: &Blob, s: &KzgSettings) -> Result<KzgCommitment, Error> {
                    let mut out = [0u8; 48];
                    let mut status = [0i32; 1];
                    let rc = unsafe {
                        ffi::kzg_b200_blob_to_kzg_commitment_batch(s.ctx, blob.as_ptr(), 1, out.as_mut_ptr(), status.as_mut_ptr())
                    };
                    check(rc, "blob_to_kzg_commitment")?;
                    first_error(&status, "blob_to_kzg_commitment")?;
                    Ok(KzgCommitment::from(out))
                }

                /// Compute the `KzgProof` given the `Blob` at the point corresponding to field element `z`
                /// (reference `src/kzg.rs:1021-1027`).
                pub fn compute_kzg_proof(blob: &Blob, z_bytes: &Bytes32, s: &KzgSettings) -> Result<(KzgProof, Bytes32), Error> {
                    let mut proof = [0u8; 48];
                    let mut y = [0u8; 32];
                    let mut status = [0i32; 1];
                    let rc = unsafe {
                        ffi::kzg_b200_compute_kzg_proof_batch(
                            s.ctx,
                            blob.as_ptr(),
                            z_bytes.as_ptr(),
                            1,
                            proof.as_mut_ptr(),
                            y.as_mut_ptr(),
                            status.as_mut_ptr(),
                        )
                    };
                    check(rc, "compute_kzg_proof")?;
                    first_error(&status, "compute_kzg_proof")?;
                    Ok((KzgProof::from(proof), Bytes32::from(y)))
                }

                /// Compute the `KzgProof` given the `Blob` and `KzgCommitment` (reference `src/kzg.rs:1030-1036`).
                pub fn compute_blob_kzg_proof(blob: &Blob, commitment_bytes: &KzgCommitment, s: &KzgSettings) -> Result<KzgProof, Error> {
                    let mut proof = [0u8; 48];
                    let mut status = [0i32; 1];
                    let rc = unsafe {
                        ffi::kzg_b200_compute_blob_kzg_proof_batch(
                            s.ctx,
                            blob.as_ptr(),
                            commitment_bytes.as_ptr(),
                            1,
                            proof.as_mut_ptr(),
                            status.as_mut_ptr(),
                        )
                    };
                    check(rc, "compute_blob_kzg_proof")?;
                    first_error(&status, "compute_blob_kzg_proof")?;
                    Ok(KzgProof::from(proof))
                }

                /// Verify a KZG proof claiming that `p(z) == y` (reference `src/kzg.rs:1039-1047`).
                pub fn verify_kzg_proof(
                    commitment_bytes: &KzgCommitment,
                    z_bytes: &Bytes32,
                    y_bytes: &Bytes32,
                    proof_bytes: &KzgProof,
                    s: &KzgSettings,
                ) -> Result<bool, Error> {
                    let mut ok: c_int = 0;
                    let rc = unsafe {
                        ffi::kzg_b200_verify_kzg_proof(
                            s.ctx,
                            commitment_bytes.as_ptr(),
                            z_bytes.as_ptr(),
                            y_bytes.as_ptr(),
                            proof_bytes.as_ptr(),
                            &mut ok,
                        )
                    };
                    check(rc, "verify_kzg_proof")?;
                    Ok(ok != 0)
                }

                /// Given a blob and its proof, verify that it corresponds to the provided commitment
                /// (reference `src/kzg.rs:1050-1063`).
                pub fn verify_blob_kzg_proof(
                    blob: &Blob,
                    commitment_bytes: &KzgCommitment,
                    proof_bytes: &KzgProof,
                    s: &KzgSettings,
                ) -> Result<bool, Error> {
                    let mut ok: c_int = 0;
                    let rc = unsafe {
                        ffi::kzg_b200_verify_blob_kzg_proof_batch(
                            s.ctx,
                            blob.as_ptr(),
                            commitment_bytes.as_ptr(),
                            proof_bytes.as_ptr(),
                            1,
                            &mut ok,
                        )
                    };
                    check(rc, "verify_blob_kzg_proof")?;
                    Ok(ok != 0)
                }

                /// Given a list of blobs and blob KZG proofs, verify that they correspond to the
                /// provided commitments (reference `src/kzg.rs:1066-1078` -> `:637-693`).
                pub fn verify_blob_kzg_proof_batch(
                    blobs: &[Blob],
                    commitment_bytes: &[KzgCommitment],
                    proof_bytes: &[KzgProof],
                    s: &KzgSettings,
                ) -> Result<bool, Error> {
                    if blobs.len() != commitment_bytes.len() || blobs.len() != proof_bytes.len() {
                        return Err(Error::BadArgs(format!(
                            "Inconsistent lengths, blobs: {}, commitments: {}, proofs: {}",
                            blobs.len(),
                            commitment_bytes.len(),
                            proof_bytes.len()
                        )));
                    }
                    if blobs.is_empty() {
                        return Ok(true);
                    }
                    let b = flatten_blobs(blobs);
                    let c = flatten48(commitment_bytes);
                    let p = flatten48(proof_bytes);
                    let mut ok: c_int = 0;
                    let rc = unsafe {
                        ffi::kzg_b200_verify_blob_kzg_proof_batch(s.ctx, b.as_ptr(), c.as_ptr(), p.as_ptr(), blobs.len(), &mut ok)
                    };
                    check(rc, "verify_blob_kzg_proof_batch")?;
                    Ok(ok != 0)
                }

                // ---- batched entry points (new with the GPU path)

                /// `blob_to_kzg_commitment` for every blob of the slice in one call.
                pub fn blob_to_kzg_commitment_batch(blobs: &[Blob], s: &KzgSettings) -> Result<Vec<KzgCommitment>, Error> {
                    let b = flatten_blobs(blobs);
                    let mut out = vec![0u8; 48 * blobs.len()];
                    let mut status = vec![0i32; blobs.len()];
                    let rc = unsafe {
                        ffi::kzg_b200_blob_to_kzg_commitment_batch(s.ctx, b.as_ptr(), blobs.len(), out.as_mut_ptr(), status.as_mut_ptr())
                    };
                    check(rc, "blob_to_kzg_commitment_batch")?;
                    first_error(&status, "blob_to_kzg_commitment_batch")?;
                    Ok(split48(&out).into_iter().map(KzgCommitment::from).collect())
                }

                /// `compute_blob_kzg_proof` for every (blob, commitment) pair in one call.
                pub fn compute_blob_kzg_proof_batch(
                    blobs: &[Blob],
                    commitment_bytes: &[KzgCommitment],
                    s: &KzgSettings,
                ) -> Result<Vec<KzgProof>, Error> {
                    if blobs.len() != commitment_bytes.len() {
                        return Err(Error::BadArgs(format!(
                            "Inconsistent lengths, blobs: {}, commitments: {}",
                            blobs.len(),
                            commitment_bytes.len()
                        )));
                    }
                    let b = flatten_blobs(blobs);
                    let c = flatten48(commitment_bytes);
                    let mut out = vec![0u8; 48 * blobs.len()];
                    let mut status = vec![0i32; blobs.len()];
                    let rc = unsafe {
                        ffi::kzg_b200_compute_blob_kzg_proof_batch(
                            s.ctx,
                            b.as_ptr(),
                            c.as_ptr(),
                            blobs.len(),
                            out.as_mut_ptr(),
                            status.as_mut_ptr(),
                        )
                    };
                    check(rc, "compute_blob_kzg_proof_batch")?;
                    first_error(&status, "compute_blob_kzg_proof_batch")?;
                    Ok(split48(&out).into_iter().map(KzgProof::from).collect())
                }

                /// `compute_kzg_proof` for every (blob, z) pair in one call.
                pub fn compute_kzg_proof_batch(
                    blobs: &[Blob],
                    z_bytes: &[Bytes32],
                    s: &KzgSettings,
                ) -> Result<Vec<(KzgProof, Bytes32)>, Error> {
                    if blobs.len() != z_bytes.len() {
                        return Err(Error::BadArgs(format!(
                            "Inconsistent lengths, blobs: {}, z: {}",
                            blobs.len(),
                            z_bytes.len()
                        )));
                    }
                    let b = flatten_blobs(blobs);
                    let mut z = Vec::with_capacity(32 * blobs.len());
                    for x in z_bytes {
                        z.extend_from_slice(x.deref());
                    }
                    let mut proofs = vec![0u8; 48 * blobs.len()];
                    let mut ys = vec![0u8; 32 * blobs.len()];
                    let mut status = vec![0i32; blobs.len()];
                    let rc = unsafe {
                        ffi::kzg_b200_compute_kzg_proof_batch(
                            s.ctx,
                            b.as_ptr(),
                            z.as_ptr(),
                            blobs.len(),
                            proofs.as_mut_ptr(),
                            ys.as_mut_ptr(),
                            status.as_mut_ptr(),
                        )
                    };
                    check(rc, "compute_kzg_proof_batch")?;
                    first_error(&status, "compute_kzg_proof_batch")?;
                    let mut out = Vec::with_capacity(blobs.len());
                    for (p, y) in split48(&proofs).into_iter().zip(ys.chunks_exact(32)) {
                        let mut ya = [0u8; 32];
                        ya.copy_from_slice(y);
                        out.push((KzgProof::from(p), Bytes32::from(ya)));
                    }
                    Ok(out)
                }
            }
        }
    };
}

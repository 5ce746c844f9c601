//! `Error`, the fixed-size byte newtypes and `hex_to_bytes` -- unchanged from the reference
//! (reference `src/kzg.rs:10-22, 82-86, 101-151, 180-279`); they do not depend on the preset.
use std::ops::Deref;

/// The number of bytes in a BLS scalar field element (reference `src/consts.rs:5`).
pub const BYTES_PER_FIELD_ELEMENT: usize = 32;
/// The number of bytes in a KZG commitment (reference `src/consts.rs:8`).
pub const BYTES_PER_COMMITMENT: usize = 48;
/// The number of bytes in a KZG proof (reference `src/consts.rs:11`).
pub const BYTES_PER_PROOF: usize = 48;
/// The number of bytes in a g1 point (reference `src/consts.rs:31`).
pub const BYTES_PER_G1: usize = 48;
/// The number of bytes in a g2 point (reference `src/consts.rs:34`).
pub const BYTES_PER_G2: usize = 96;
/// The number of g2 points in a trusted setup (reference `src/consts.rs:37`).
pub const TRUSTED_SETUP_NUM_G2_POINTS: usize = 65;

/// reference `enum Error`, `src/kzg.rs:10-22`.
#[derive(Debug)]
pub enum Error {
    /// The supplied data is invalid in some way.
    BadArgs(String),
    /// Internal error - this should never occur.  Also reported when the CUDA runtime fails
    /// (`KZG_B200_CUDA_ERROR`): there is no CPU fallback to mask it.
    InternalError,
    /// The provided bytes are of incorrect length.
    InvalidBytesLength(String),
    /// Error when converting from hex to bytes.
    InvalidHexFormat(String),
    /// The provided trusted setup params are invalid.
    InvalidTrustedSetup(String),
}

impl Error {
    /// Status code of `include/kzg_b200.h` (which mirrors this enum in declaration order) -> `Error`.
    pub(crate) fn from_code(code: i32, what: &str) -> Error {
        match code {
            crate::ffi::KZG_B200_BAD_ARGS => Error::BadArgs(what.to_string()),
            crate::ffi::KZG_B200_INVALID_BYTES_LENGTH => Error::InvalidBytesLength(what.to_string()),
            crate::ffi::KZG_B200_INVALID_HEX_FORMAT => Error::InvalidHexFormat(what.to_string()),
            crate::ffi::KZG_B200_INVALID_TRUSTED_SETUP => Error::InvalidTrustedSetup(what.to_string()),
            _ => Error::InternalError,
        }
    }
}

/// `Ok(())` for `KZG_B200_OK`, the mapped error otherwise.
pub(crate) fn check(code: i32, what: &str) -> Result<(), Error> {
    if code == crate::ffi::KZG_B200_OK {
        Ok(())
    } else {
        Err(Error::from_code(code, what))
    }
}

/// Converts a hex string (with or without the 0x prefix) to bytes (reference `src/kzg.rs:82-86`).
pub fn hex_to_bytes(hex_str: &str) -> Result<Vec<u8>, Error> {
    let trimmed_str = hex_str.strip_prefix("0x").unwrap_or(hex_str);
    hex::decode(trimmed_str).map_err(|e| Error::InvalidHexFormat(format!("Failed to decode hex: {}", e)))
}

/// reference `src/kzg.rs:101-122`.
#[derive(Default, Debug, Copy, Clone, PartialEq)]
pub struct Bytes32 {
    pub(crate) bytes: [u8; 32],
}

impl Bytes32 {
    pub fn from_bytes(b: &[u8]) -> Result<Self, Error> {
        if b.len() != 32 {
            return Err(Error::BadArgs(format!("Bytes32 length error. Expected 32, got {}", b.len())));
        }
        let mut arr = [0; 32];
        arr.copy_from_slice(b);
        Ok(Bytes32 { bytes: arr })
    }

    pub fn from_hex(hex_str: &str) -> Result<Self, Error> {
        Self::from_bytes(&hex_to_bytes(hex_str)?)
    }
}

/// reference `src/kzg.rs:124-151`.
#[derive(Debug, Copy, Clone, PartialEq)]
pub struct Bytes48 {
    pub(crate) bytes: [u8; 48],
}

impl Bytes48 {
    pub fn from_bytes(bytes: &[u8]) -> Result<Self, Error> {
        if bytes.len() != 48 {
            return Err(Error::InvalidBytesLength(format!(
                "Invalid byte length. Expected {} got {}",
                48,
                bytes.len(),
            )));
        }
        let mut new_bytes = [0; 48];
        new_bytes.copy_from_slice(bytes);
        Ok(Self { bytes: new_bytes })
    }

    pub fn from_hex(hex_str: &str) -> Result<Self, Error> {
        Self::from_bytes(&hex_to_bytes(hex_str)?)
    }
}

impl Default for Bytes48 {
    fn default() -> Self {
        Self { bytes: [0; 48] }
    }
}

/// reference `src/kzg.rs:180-191`.
#[derive(Debug, Copy, Clone, PartialEq)]
pub struct KzgCommitment(pub Bytes48);

impl KzgCommitment {
    pub fn from_hex(hex_str: &str) -> Result<Self, Error> {
        Ok(Self(Bytes48::from_bytes(&hex_to_bytes(hex_str)?)?))
    }

    pub fn to_bytes(self) -> [u8; BYTES_PER_COMMITMENT] {
        self.0.bytes
    }
}

/// reference `src/kzg.rs:193-204`.
#[derive(Debug, Copy, Clone, PartialEq)]
pub struct KzgProof(pub Bytes48);

impl KzgProof {
    pub fn from_hex(hex_str: &str) -> Result<Self, Error> {
        Ok(Self(Bytes48::from_bytes(&hex_to_bytes(hex_str)?)?))
    }

    pub fn to_bytes(self) -> [u8; BYTES_PER_PROOF] {
        self.0.bytes
    }
}

impl From<[u8; BYTES_PER_COMMITMENT]> for KzgCommitment {
    fn from(value: [u8; BYTES_PER_COMMITMENT]) -> Self {
        Self(Bytes48 { bytes: value })
    }
}

impl From<[u8; BYTES_PER_PROOF]> for KzgProof {
    fn from(value: [u8; BYTES_PER_PROOF]) -> Self {
        Self(Bytes48 { bytes: value })
    }
}

impl From<[u8; 32]> for Bytes32 {
    fn from(value: [u8; 32]) -> Self {
        Self { bytes: value }
    }
}

impl From<[u8; 48]> for Bytes48 {
    fn from(value: [u8; 48]) -> Self {
        Self { bytes: value }
    }
}

impl Deref for Bytes32 {
    type Target = [u8; 32];
    fn deref(&self) -> &Self::Target {
        &self.bytes
    }
}

impl Deref for Bytes48 {
    type Target = [u8; 48];
    fn deref(&self) -> &Self::Target {
        &self.bytes
    }
}

impl Deref for KzgProof {
    type Target = [u8; BYTES_PER_PROOF];
    fn deref(&self) -> &Self::Target {
        &self.0.bytes
    }
}

impl Deref for KzgCommitment {
    type Target = [u8; BYTES_PER_COMMITMENT];
    fn deref(&self) -> &Self::Target {
        &self.0.bytes
    }
}

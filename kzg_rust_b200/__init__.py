"""kzg_rust_b200 -- B200-native blob path of pawanjay176/kzg_rust behind the crate's `Kzg` API.

`kzg_mainnet` / `kzg_minimal` (reference README.md:8-9) are the same code with
FIELD_ELEMENTS_PER_BLOB = 4096 / 4; the preset is fixed by the trusted setup a context is
created from."""
from .kzg import (BYTES_PER_COMMITMENT, BYTES_PER_FIELD_ELEMENT, BYTES_PER_G1, BYTES_PER_G2, BYTES_PER_PROOF,
                  TRUSTED_SETUP_NUM_G2_POINTS, BadArgs, Blob, BlobMinimal, Bytes32, Bytes48, CudaError, Error,
                  InternalError, InvalidBytesLength, InvalidHexFormat, InvalidTrustedSetup, Kzg, KzgCommitment,
                  KzgProof, KzgSettings, TrustedSetup, hex_to_bytes, load_library)

__all__ = [n for n in dir() if not n.startswith("_")]

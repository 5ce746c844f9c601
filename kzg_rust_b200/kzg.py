"""Host-side mirror of the reference crate's public API over the CUDA library.

Names, argument meaning and error behaviour follow `impl Kzg` and its byte newtypes
(reference src/kzg.rs:10-22, 101-279, 983-1079); every method forwards to the C ABI of
include/kzg_b200.h with n = 1, and the `*_batch` methods expose n > 1.  There is no CPU
path: importing works anywhere, but using `Kzg` needs libkzg_b200.so and a CUDA device.
"""
import ctypes
import os

import numpy as np

BYTES_PER_FIELD_ELEMENT = 32
BYTES_PER_COMMITMENT = 48
BYTES_PER_PROOF = 48
BYTES_PER_G1 = 48
BYTES_PER_G2 = 96
TRUSTED_SETUP_NUM_G2_POINTS = 65

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkzg_b200.so")


# ---- reference `enum Error` (src/kzg.rs:10-22)
class Error(Exception):
    pass


class BadArgs(Error):
    pass


class InternalError(Error):
    pass


class InvalidBytesLength(Error):
    pass


class InvalidHexFormat(Error):
    pass


class InvalidTrustedSetup(Error):
    pass


class CudaError(Error):
    """No reference counterpart: the CUDA runtime failed (there is no CPU fallback)."""


_ERRORS = {1: BadArgs, 2: InternalError, 3: InvalidBytesLength, 4: InvalidHexFormat,
           5: InvalidTrustedSetup, 6: CudaError}


def _raise(code, what=""):
    raise _ERRORS.get(code, InternalError)("%s (code %d)" % (what, code))


_lib = None


def load_library():
    """dlopen the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("KZG_B200_LIB", LIB_PATH)  # experiments load a differently built copy (tools/variant_sweep.sh)
    if not os.path.exists(path):
        raise CudaError("%s is missing; build it with `python -m kzg_rust_b200.build` "
                        "(there is no CPU fallback)" % os.path.basename(path))
    L = ctypes.CDLL(path)
    vp, cp, sz, ci = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int
    pvp, pci = ctypes.POINTER(vp), ctypes.POINTER(ci)
    L.kzg_b200_ctx_create.argtypes = [vp, sz, vp, sz, ci, ci, pvp]
    L.kzg_b200_ctx_create_from_file.argtypes = [cp, ci, ci, pvp]
    L.kzg_b200_ctx_destroy.argtypes = [vp]
    L.kzg_b200_ctx_destroy.restype = None
    L.kzg_b200_field_elements_per_blob.argtypes = [vp]
    L.kzg_b200_field_elements_per_blob.restype = sz
    L.kzg_b200_comb_width.argtypes = [vp]
    L.kzg_b200_table_bytes.argtypes = [vp]
    L.kzg_b200_table_bytes.restype = sz
    L.kzg_b200_chunk_blobs.argtypes = [vp]
    L.kzg_b200_chunk_blobs.restype = sz
    L.kzg_b200_debug_check_tau_identity.argtypes = [vp, vp, vp, sz, cp, vp]
    L.kzg_b200_blob_to_kzg_commitment_batch.argtypes = [vp, vp, sz, vp, vp]
    L.kzg_b200_compute_blob_kzg_proof_batch.argtypes = [vp, vp, vp, sz, vp, vp]
    L.kzg_b200_compute_kzg_proof_batch.argtypes = [vp, vp, vp, sz, vp, vp, vp]
    L.kzg_b200_verify_blob_kzg_proof_batch.argtypes = [vp, vp, vp, vp, sz, pci]
    L.kzg_b200_verify_phase_a.argtypes = [vp, vp, vp, vp, sz, vp]
    L.kzg_b200_compute_r.argtypes = [vp, vp, vp, vp, sz, vp]
    L.kzg_b200_verify_phase_b.argtypes = [vp, vp, vp, vp, sz, vp, ctypes.c_uint64, vp]
    L.kzg_b200_verify_finish.argtypes = [vp, vp, sz, pci]
    L.kzg_b200_blob_to_kzg_commitment_device.argtypes = [vp, vp, sz, vp, vp]
    L.kzg_b200_compute_blob_kzg_proof_device.argtypes = [vp, vp, vp, sz, vp, vp]
    L.kzg_b200_synchronize.argtypes = [vp]
    L.kzg_b200_verify_blob_kzg_proof_batch_device.argtypes = [vp, vp, vp, vp, sz, pci]
    L.kzg_b200_verify_phase_a_device.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp]
    L.kzg_b200_stream.argtypes = [vp]
    L.kzg_b200_stream.restype = vp
    L.kzg_b200_verify_kzg_proof.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
    L.kzg_b200_launch_count.argtypes = [vp]
    L.kzg_b200_launch_count.restype = ctypes.c_uint64
    L.kzg_b200_profile_enable.argtypes = [vp, ci]
    L.kzg_b200_profile_read.argtypes = [vp, vp, vp]
    L.kzg_b200_pairings_verify.argtypes = [cp, cp, cp, cp, pci]
    L.kzg_b200_measure_peaks.argtypes = [vp] + [ctypes.POINTER(ctypes.c_double)] * 3
    _lib = L
    return L


def hex_to_bytes(hex_str):
    """reference src/kzg.rs:82-86 (with or without the 0x prefix)."""
    s = hex_str[2:] if hex_str.startswith("0x") else hex_str
    try:
        return bytes.fromhex(s)
    except ValueError as e:
        raise InvalidHexFormat("Failed to decode hex: %s" % e)


class _FixedBytes:
    SIZE = 0
    LENGTH_ERROR = InvalidBytesLength

    def __init__(self, b):
        b = bytes(b)
        if len(b) != self.SIZE:
            raise self.LENGTH_ERROR("Invalid byte length. Expected %d got %d" % (self.SIZE, len(b)))
        self.bytes = b

    @classmethod
    def from_bytes(cls, b):
        return cls(b)

    @classmethod
    def from_hex(cls, hex_str):
        return cls(hex_to_bytes(hex_str))

    def to_bytes(self):
        return self.bytes

    def __bytes__(self):
        return self.bytes

    def __eq__(self, other):
        return type(self) is type(other) and self.bytes == other.bytes

    def __hash__(self):
        return hash((type(self).__name__, self.bytes))

    def __repr__(self):
        return "%s(0x%s)" % (type(self).__name__, self.bytes[:8].hex() + ("..." if self.SIZE > 8 else ""))


class Bytes32(_FixedBytes):
    """reference src/kzg.rs:101-122 (a wrong length is `BadArgs` there)."""
    SIZE = 32
    LENGTH_ERROR = BadArgs


class Bytes48(_FixedBytes):
    """reference src/kzg.rs:124-151."""
    SIZE = 48


class KzgCommitment(Bytes48):
    """reference src/kzg.rs:180-191."""


class KzgProof(Bytes48):
    """reference src/kzg.rs:193-204."""


def make_blob_type(n):
    class Blob(_FixedBytes):
        """reference src/kzg.rs:153-178."""
        SIZE = 32 * n
    return Blob


Blob = make_blob_type(4096)          # kzg_mainnet
BlobMinimal = make_blob_type(4)      # kzg_minimal


def _raw(x):
    return x.bytes if isinstance(x, _FixedBytes) else bytes(x)


class KzgSettings:
    """Plays the role of reference `KzgSettings` (src/kzg.rs:27-40): owns the device-resident
    tables built from the trusted setup."""

    def __init__(self, handle):
        self._h = handle
        L = load_library()
        self.field_elements_per_blob = L.kzg_b200_field_elements_per_blob(handle)
        self.bytes_per_blob = 32 * self.field_elements_per_blob
        self.comb_width = L.kzg_b200_comb_width(handle)
        self.table_bytes = L.kzg_b200_table_bytes(handle)
        self.chunk_blobs = L.kzg_b200_chunk_blobs(handle)

    @classmethod
    def load_trusted_setup(cls, g1_bytes, g2_bytes, device=0, comb_width=0):
        """reference src/kzg.rs:45-79.  g1_bytes / g2_bytes: lists of 48 / 96 byte strings."""
        g1 = b"".join(bytes(x) for x in g1_bytes) if not isinstance(g1_bytes, (bytes, bytearray)) else bytes(g1_bytes)
        g2 = b"".join(bytes(x) for x in g2_bytes) if not isinstance(g2_bytes, (bytes, bytearray)) else bytes(g2_bytes)
        n1, n2 = len(g1) // BYTES_PER_G1, len(g2) // BYTES_PER_G2
        if n1 not in (4096, 4) or len(g1) % BYTES_PER_G1:
            raise InvalidTrustedSetup("Invalid number of g1 points in trusted setup. Expected 4096 or 4 got %d" % n1)
        if n2 != TRUSTED_SETUP_NUM_G2_POINTS or len(g2) % BYTES_PER_G2:
            raise InvalidTrustedSetup("Invalid number of g2 points in trusted setup. Expected 65 got %d" % n2)
        L = load_library()
        h = ctypes.c_void_p()
        b1 = ctypes.create_string_buffer(g1, len(g1))
        b2 = ctypes.create_string_buffer(g2, len(g2))
        rc = L.kzg_b200_ctx_create(ctypes.addressof(b1), n1, ctypes.addressof(b2), n2, device, comb_width,
                                   ctypes.byref(h))
        if rc:
            _raise(rc, "load_trusted_setup")
        return cls(h)

    @classmethod
    def load_trusted_setup_file(cls, path, device=0, comb_width=0):
        """reference src/kzg.rs:906-979."""
        L = load_library()
        h = ctypes.c_void_p()
        rc = L.kzg_b200_ctx_create_from_file(os.fsencode(path), device, comb_width, ctypes.byref(h))
        if rc:
            _raise(rc, "load_trusted_setup_file")
        return cls(h)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.kzg_b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class TrustedSetup:
    """reference `TrustedSetup` (src/trusted_setup.rs:13-44, 138-153): the trusted setup in the json format of the
    ethereum consensus specs -- {"setup_G1_lagrange": [hex, ...], "setup_G2": [hex, ...]}, hex with or without 0x.
    The G1 list is truncated to the preset's FIELD_ELEMENTS_PER_BLOB after parsing, as the reference does."""

    def __init__(self, g1_points, g2_points):
        self._g1 = [bytes(p) for p in g1_points]
        self._g2 = [bytes(p) for p in g2_points]

    @staticmethod
    def _point(v, size, what):
        if not isinstance(v, str):
            raise InvalidTrustedSetup("A %d byte hex encoded string" % size)
        try:
            b = bytes.fromhex(v[2:] if v.startswith("0x") else v)
        except ValueError as e:
            raise InvalidTrustedSetup("Failed to decode %s point: %s" % (what, e))
        if len(b) != size:
            raise InvalidTrustedSetup("%s point has invalid length. Expected %d got %d" % (what, size, len(b)))
        return b

    @classmethod
    def from_json(cls, text, field_elements_per_blob=4096):
        import json
        try:
            doc = json.loads(text)
            g1, g2 = doc["setup_G1_lagrange"], doc["setup_G2"]
        except (ValueError, KeyError, TypeError) as e:
            raise InvalidTrustedSetup("not a trusted setup in json form: %s" % e)
        g1 = [cls._point(v, BYTES_PER_G1, "G1") for v in g1][:field_elements_per_blob]
        g2 = [cls._point(v, BYTES_PER_G2, "G2") for v in g2]
        return cls(g1, g2)

    @classmethod
    def from_json_file(cls, path, field_elements_per_blob=4096):
        with open(path) as fh:
            return cls.from_json(fh.read(), field_elements_per_blob)

    def to_json(self):
        import json
        return json.dumps({"setup_G1_lagrange": [p.hex() for p in self._g1], "setup_G2": [p.hex() for p in self._g2]})

    def g1_points(self):
        return list(self._g1)

    def g2_points(self):
        return list(self._g2)

    def g1_len(self):
        return len(self._g1)

    def g2_len(self):
        return len(self._g2)


def _np_u8(buf, nbytes):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.reshape(-1).view(np.uint8)
    if a.size != nbytes:
        raise InvalidBytesLength("Invalid byte length. Expected %d got %d" % (nbytes, a.size))
    return np.ascontiguousarray(a)


class Kzg:
    """reference `pub struct Kzg` (src/kzg.rs:983-1079)."""

    @staticmethod
    def load_trusted_setup_file(path, device=0, comb_width=0):
        return KzgSettings.load_trusted_setup_file(path, device, comb_width)

    @staticmethod
    def load_trusted_setup(g1_bytes, g2_bytes, device=0, comb_width=0):
        return KzgSettings.load_trusted_setup(g1_bytes, g2_bytes, device, comb_width)

    # ---- batched entry points (new): contiguous buffers in, numpy arrays out
    @staticmethod
    def blob_to_kzg_commitment_batch(blobs, s):
        """blobs: bytes / numpy uint8 of n * bytes_per_blob.  Returns (commitments[n,48], status[n])."""
        L = load_library()
        nbytes = blobs.nbytes if isinstance(blobs, np.ndarray) else len(blobs)
        if nbytes % s.bytes_per_blob:
            raise InvalidBytesLength("blobs buffer is not a multiple of %d bytes" % s.bytes_per_blob)
        n = nbytes // s.bytes_per_blob
        a = _np_u8(blobs, nbytes)
        out = np.zeros((n, 48), dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        rc = L.kzg_b200_blob_to_kzg_commitment_batch(s._h, a.ctypes.data, n, out.ctypes.data, status.ctypes.data)
        if rc:
            _raise(rc, "blob_to_kzg_commitment_batch")
        return out, status

    @staticmethod
    def compute_blob_kzg_proof_batch(blobs, commitments, s):
        L = load_library()
        nbytes = blobs.nbytes if isinstance(blobs, np.ndarray) else len(blobs)
        if nbytes % s.bytes_per_blob:
            raise InvalidBytesLength("blobs buffer is not a multiple of %d bytes" % s.bytes_per_blob)
        n = nbytes // s.bytes_per_blob
        a = _np_u8(blobs, nbytes)
        c = _np_u8(commitments, 48 * n)
        out = np.zeros((n, 48), dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        rc = L.kzg_b200_compute_blob_kzg_proof_batch(s._h, a.ctypes.data, c.ctypes.data, n, out.ctypes.data,
                                                     status.ctypes.data)
        if rc:
            _raise(rc, "compute_blob_kzg_proof_batch")
        return out, status

    @staticmethod
    def compute_kzg_proof_batch(blobs, zs, s):
        L = load_library()
        nbytes = blobs.nbytes if isinstance(blobs, np.ndarray) else len(blobs)
        if nbytes % s.bytes_per_blob:
            raise InvalidBytesLength("blobs buffer is not a multiple of %d bytes" % s.bytes_per_blob)
        n = nbytes // s.bytes_per_blob
        a = _np_u8(blobs, nbytes)
        z = _np_u8(zs, 32 * n)
        proofs = np.zeros((n, 48), dtype=np.uint8)
        ys = np.zeros((n, 32), dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        rc = L.kzg_b200_compute_kzg_proof_batch(s._h, a.ctypes.data, z.ctypes.data, n, proofs.ctypes.data,
                                                ys.ctypes.data, status.ctypes.data)
        if rc:
            _raise(rc, "compute_kzg_proof_batch")
        return proofs, ys, status

    @staticmethod
    def verify_blob_kzg_proof_batch_raw(blobs, commitments, proofs, n, s):
        """Contiguous-buffer form of verify_blob_kzg_proof_batch."""
        L = load_library()
        a = _np_u8(blobs, n * s.bytes_per_blob)
        c = _np_u8(commitments, 48 * n)
        p = _np_u8(proofs, 48 * n)
        ok = ctypes.c_int(0)
        rc = L.kzg_b200_verify_blob_kzg_proof_batch(s._h, a.ctypes.data, c.ctypes.data, p.ctypes.data, n,
                                                    ctypes.byref(ok))
        if rc:
            _raise(rc, "verify_blob_kzg_proof_batch")
        return bool(ok.value)

    # ---- the reference's single-blob API (src/kzg.rs:1013-1078)
    @staticmethod
    def _one(buf, size):
        """A single fixed-size record (the reference's typed `Blob` / `Bytes32` / `Bytes48` make any other length
        impossible; raw buffers get the same check here instead of being read as a batch)."""
        b = _raw(buf)
        if len(b) != size:
            raise InvalidBytesLength("Invalid byte length. Expected %d got %d" % (size, len(b)))
        return b

    @staticmethod
    def blob_to_kzg_commitment(blob, s):
        out, status = Kzg.blob_to_kzg_commitment_batch(Kzg._one(blob, s.bytes_per_blob), s)
        if status[0]:
            _raise(int(status[0]), "blob_to_kzg_commitment")
        return KzgCommitment(out[0].tobytes())

    @staticmethod
    def compute_kzg_proof(blob, z_bytes, s):
        proofs, ys, status = Kzg.compute_kzg_proof_batch(Kzg._one(blob, s.bytes_per_blob), Kzg._one(z_bytes, 32), s)
        if status[0]:
            _raise(int(status[0]), "compute_kzg_proof")
        return KzgProof(proofs[0].tobytes()), Bytes32(ys[0].tobytes())

    @staticmethod
    def compute_blob_kzg_proof(blob, commitment_bytes, s):
        out, status = Kzg.compute_blob_kzg_proof_batch(Kzg._one(blob, s.bytes_per_blob), Kzg._one(commitment_bytes, 48), s)
        if status[0]:
            _raise(int(status[0]), "compute_blob_kzg_proof")
        return KzgProof(out[0].tobytes())

    @staticmethod
    def verify_kzg_proof(commitment_bytes, z_bytes, y_bytes, proof_bytes, s):
        """reference src/kzg.rs:1039-1047."""
        L = load_library()
        c, z, y, p = _raw(commitment_bytes), _raw(z_bytes), _raw(y_bytes), _raw(proof_bytes)
        if len(c) != 48 or len(p) != 48:
            raise InvalidBytesLength("Invalid byte length. Expected 48 got %d" % (len(c) if len(c) != 48 else len(p)))
        if len(z) != 32 or len(y) != 32:
            raise InvalidBytesLength("Invalid byte length. Expected 32 got %d" % (len(z) if len(z) != 32 else len(y)))
        ok = ctypes.c_int(0)
        rc = L.kzg_b200_verify_kzg_proof(s._h, bytes(c), bytes(z), bytes(y), bytes(p), ctypes.byref(ok))
        if rc:
            _raise(rc, "verify_kzg_proof")
        return bool(ok.value)

    @staticmethod
    def verify_blob_kzg_proof(blob, commitment_bytes, proof_bytes, s):
        return Kzg.verify_blob_kzg_proof_batch_raw(Kzg._one(blob, s.bytes_per_blob), Kzg._one(commitment_bytes, 48),
                                                   Kzg._one(proof_bytes, 48), 1, s)

    @staticmethod
    def verify_blob_kzg_proof_batch(blobs, commitments, proofs, s):
        """reference src/kzg.rs:637-693: lists of Blob / KzgCommitment / KzgProof."""
        if not (len(blobs) == len(commitments) == len(proofs)):
            raise BadArgs("Inconsistent lengths, blobs: %d, commitments: %d, proofs: %d"
                          % (len(blobs), len(commitments), len(proofs)))
        n = len(blobs)
        if n == 0:
            return True
        return Kzg.verify_blob_kzg_proof_batch_raw(b"".join(_raw(b) for b in blobs),
                                                   b"".join(_raw(c) for c in commitments),
                                                   b"".join(_raw(p) for p in proofs), n, s)

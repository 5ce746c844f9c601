"""ctypes binding of oracle/libkzg_oracle.so (see kzg_oracle.c).  TEST INFRASTRUCTURE ONLY.

The wrappers mirror the reference's `Kzg` associated functions (reference
src/kzg.rs:983-1079): they return the value on success and raise OracleError
where the reference returns `Err(_)`.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkzg_oracle.so")


class OracleError(Exception):
    def __init__(self, code):
        super().__init__("oracle error code %d" % code)
        self.code = code


def build(force=False):
    src = os.path.join(_HERE, "kzg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, cp, sz, ci = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int
        pi = ctypes.POINTER(ctypes.c_int)
        L.ko_load_trusted_setup.argtypes = [cp, sz, cp, sz, ctypes.POINTER(vp)]
        L.ko_load_trusted_setup_file.argtypes = [cp, ctypes.POINTER(vp)]
        L.ko_free.argtypes = [vp]
        L.ko_free.restype = None
        L.ko_field_elements_per_blob.argtypes = [vp]
        L.ko_field_elements_per_blob.restype = sz
        L.ko_blob_to_kzg_commitment.argtypes = [vp, cp, cp]
        L.ko_compute_kzg_proof.argtypes = [vp, cp, cp, cp, cp]
        L.ko_compute_blob_kzg_proof.argtypes = [vp, cp, cp, cp]
        L.ko_verify_kzg_proof.argtypes = [vp, cp, cp, cp, cp, pi]
        L.ko_verify_blob_kzg_proof.argtypes = [vp, cp, cp, cp, pi]
        L.ko_verify_blob_kzg_proof_batch.argtypes = [vp, cp, cp, cp, sz, pi]
        L.ko_blob_to_kzg_commitment_many.argtypes = [vp, vp, sz, vp, vp, ci]
        L.ko_compute_blob_kzg_proof_many.argtypes = [vp, vp, vp, sz, vp, vp, ci]
        L.ko_blob_to_kzg_commitment_many_fast.argtypes = [vp, vp, sz, vp, vp, ci]
        L.ko_compute_challenge.argtypes = [vp, cp, cp, cp]
        L.ko_evaluate_polynomial.argtypes = [vp, cp, cp, cp]
        L.ko_validate_kzg_g1.argtypes = [cp]
        L.ko_g1_lincomb.argtypes = [cp, cp, sz, cp]
        L.ko_sha256.argtypes = [cp, sz, cp]
        L.ko_sha256.restype = None
        L.ko_pairings_verify.argtypes = [cp, cp, cp, cp, pi]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise OracleError(rc)


class OracleSettings:
    """Stand-in for `KzgSettings` (reference src/kzg.rs:27-40)."""

    def __init__(self, handle):
        self._h = handle
        self.n = lib().ko_field_elements_per_blob(handle)
        self.bytes_per_blob = 32 * self.n

    @classmethod
    def load_trusted_setup(cls, g1_bytes, g2_bytes):
        h = ctypes.c_void_p()
        _check(lib().ko_load_trusted_setup(g1_bytes, len(g1_bytes) // 48, g2_bytes, len(g2_bytes) // 96,
                                           ctypes.byref(h)))
        return cls(h)

    @classmethod
    def load_trusted_setup_file(cls, path):
        h = ctypes.c_void_p()
        _check(lib().ko_load_trusted_setup_file(os.fsencode(path), ctypes.byref(h)))
        return cls(h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ko_free(self._h)
            self._h = None

    # ---- the Kzg API (reference src/kzg.rs:1013-1078)
    def blob_to_kzg_commitment(self, blob):
        assert len(blob) == self.bytes_per_blob
        out = ctypes.create_string_buffer(48)
        _check(lib().ko_blob_to_kzg_commitment(self._h, blob, out))
        return out.raw

    def compute_kzg_proof(self, blob, z):
        assert len(blob) == self.bytes_per_blob and len(z) == 32
        proof, y = ctypes.create_string_buffer(48), ctypes.create_string_buffer(32)
        _check(lib().ko_compute_kzg_proof(self._h, blob, z, proof, y))
        return proof.raw, y.raw

    def compute_blob_kzg_proof(self, blob, commitment):
        assert len(blob) == self.bytes_per_blob and len(commitment) == 48
        out = ctypes.create_string_buffer(48)
        _check(lib().ko_compute_blob_kzg_proof(self._h, blob, commitment, out))
        return out.raw

    def verify_kzg_proof(self, commitment, z, y, proof):
        ok = ctypes.c_int(0)
        _check(lib().ko_verify_kzg_proof(self._h, commitment, z, y, proof, ctypes.byref(ok)))
        return bool(ok.value)

    def verify_blob_kzg_proof(self, blob, commitment, proof):
        ok = ctypes.c_int(0)
        _check(lib().ko_verify_blob_kzg_proof(self._h, blob, commitment, proof, ctypes.byref(ok)))
        return bool(ok.value)

    def verify_blob_kzg_proof_batch(self, blobs, commitments, proofs):
        if not (len(blobs) == len(commitments) == len(proofs)):
            raise OracleError(1)  # reference src/kzg.rs:644-651
        ok = ctypes.c_int(0)
        _check(lib().ko_verify_blob_kzg_proof_batch(self._h, b"".join(blobs), b"".join(commitments),
                                                    b"".join(proofs), len(blobs), ctypes.byref(ok)))
        return bool(ok.value)

    # ---- many-blob drivers over contiguous buffers (numpy uint8 arrays); used for timing
    def blob_to_kzg_commitment_many(self, blobs_np, nthreads=1):
        import numpy as np
        n = blobs_np.size // self.bytes_per_blob
        out = np.zeros((n, 48), dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        lib().ko_blob_to_kzg_commitment_many(self._h, blobs_np.ctypes.data, n, out.ctypes.data,
                                             status.ctypes.data, nthreads)
        return out, status

    def blob_to_kzg_commitment_many_fast(self, blobs_np, nthreads=1):
        """The timed CPU baseline (bench.py): the same bytes as blob_to_kzg_commitment_many from the fast
        restatement of the bucket method (batch-affine buckets, mulx Montgomery) -- see kzg_oracle.c."""
        import numpy as np
        blobs_np = np.ascontiguousarray(blobs_np)
        n = blobs_np.size // self.bytes_per_blob
        out = np.zeros((n, 48), dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        lib().ko_blob_to_kzg_commitment_many_fast(self._h, blobs_np.ctypes.data, n, out.ctypes.data,
                                                  status.ctypes.data, nthreads)
        return out, status

    def compute_blob_kzg_proof_many(self, blobs_np, commitments_np, nthreads=1):
        import numpy as np
        n = blobs_np.size // self.bytes_per_blob
        out = np.zeros((n, 48), dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        lib().ko_compute_blob_kzg_proof_many(self._h, blobs_np.ctypes.data, commitments_np.ctypes.data, n,
                                             out.ctypes.data, status.ctypes.data, nthreads)
        return out, status

    # ---- primitives for kernel-level checks
    def compute_challenge(self, blob, commitment):
        out = ctypes.create_string_buffer(32)
        _check(lib().ko_compute_challenge(self._h, blob, commitment, out))
        return out.raw

    def evaluate_polynomial(self, blob, z):
        out = ctypes.create_string_buffer(32)
        _check(lib().ko_evaluate_polynomial(self._h, blob, z, out))
        return out.raw


def validate_kzg_g1(b):
    return lib().ko_validate_kzg_g1(b) == 0


def g1_lincomb(points, scalars):
    out = ctypes.create_string_buffer(48)
    _check(lib().ko_g1_lincomb(b"".join(points), b"".join(scalars), len(points), out))
    return out.raw


def sha256(data):
    out = ctypes.create_string_buffer(32)
    lib().ko_sha256(data, len(data), out)
    return out.raw


def pairings_verify(a1, a2, b1, b2):
    ok = ctypes.c_int(0)
    _check(lib().ko_pairings_verify(a1, a2, b1, b2, ctypes.byref(ok)))
    return bool(ok.value)

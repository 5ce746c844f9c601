"""-m gpu: the C ABI from a plain C program (tests/cabi_smoke.c, built with gcc against include/kzg_b200.h and
linked to libkzg_b200.so) -- the boundary a Rust / Go / C caller binds, with no ctypes in between."""
import os
import subprocess

import pytest

from golden_util import GOLDEN, golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_commits_proves_and_verifies_a_reference_vector(tmp_path):
    import kzg_rust_b200  # noqa: F401  (fails loudly when the library has not been built)
    libdir = os.path.join(ROOT, "kzg_rust_b200")
    exe = str(tmp_path / "cabi_smoke")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "cabi_smoke.c"),
                           "-L", libdir, "-lkzg_b200", "-Wl,-rpath," + libdir])
    g = golden()
    case = next(c for c in g.by_fn("blob_to_kzg_commitment") if c["output"] is not None)
    (tmp_path / "blob.bin").write_bytes(g.get_bytes(case["input"]["blob"]))
    (tmp_path / "expected.bin").write_bytes(bytes.fromhex(case["output"][2:]))
    out = subprocess.run([exe, os.path.join(GOLDEN, "trusted_setup.bin"), str(tmp_path / "blob.bin"), str(tmp_path / "expected.bin")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "cabi_smoke ok" in out.stdout

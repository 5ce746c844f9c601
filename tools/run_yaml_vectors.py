#!/usr/bin/env python3
"""Run a tree of c-kzg-4844 style YAML vectors -- <root>/<function>/<suite>/<case>/data.yaml, the layout the
reference's own test runner walks (src/lib.rs:14-204, schemas in src/test_formats/*.rs) -- through this repo's
mirror of `impl Kzg`, i.e. through the C ABI on the GPU.

    python tools/run_yaml_vectors.py /path/to/tests [trusted_setup.txt] [--oracle]

`output: null` means the call must fail (any Error), as in the reference's runner.  --oracle runs the CPU
oracle instead of the GPU path (test infrastructure; works without a GPU).  Exit code 1 on any mismatch."""
import glob
import os
import sys

import yaml

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def unhex(x):
    s = x[2:] if x.startswith("0x") else x
    return bytes.fromhex(s)


class GpuApi:
    def __init__(self, setup_path):
        import kzg_rust_b200 as k
        self.k = k
        if setup_path:
            self.s = k.Kzg.load_trusted_setup_file(setup_path, 0, 8)
        else:
            from golden_util import golden
            g = golden()
            self.s = k.KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, 8)
        self.Error = (k.Error, ValueError)

    def blob(self, b): return self.k.Blob.from_bytes(b)
    def b48(self, b): return self.k.Bytes48.from_bytes(b)
    def b32(self, b): return self.k.Bytes32.from_bytes(b)
    def blob_to_kzg_commitment(self, blob): return bytes(self.k.Kzg.blob_to_kzg_commitment(blob, self.s))
    def compute_kzg_proof(self, blob, z):
        p, y = self.k.Kzg.compute_kzg_proof(blob, z, self.s)
        return bytes(p), bytes(y)
    def compute_blob_kzg_proof(self, blob, c): return bytes(self.k.Kzg.compute_blob_kzg_proof(blob, c, self.s))
    def verify_kzg_proof(self, c, z, y, p): return self.k.Kzg.verify_kzg_proof(c, z, y, p, self.s)
    def verify_blob_kzg_proof(self, blob, c, p): return self.k.Kzg.verify_blob_kzg_proof(blob, c, p, self.s)
    def verify_blob_kzg_proof_batch(self, blobs, cs, ps): return self.k.Kzg.verify_blob_kzg_proof_batch(blobs, cs, ps, self.s)


class OracleApi:
    def __init__(self, setup_path):
        from oracle.binding import OracleError, OracleSettings
        if setup_path:
            self.s = OracleSettings.load_trusted_setup_file(setup_path)
        else:
            from golden_util import golden
            g = golden()
            self.s = OracleSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes)
        self.Error = (OracleError, ValueError)

    @staticmethod
    def _sized(b, n):
        if len(b) != n:
            raise ValueError("length")
        return b
    def blob(self, b): return self._sized(b, 131072)
    def b48(self, b): return self._sized(b, 48)
    def b32(self, b): return self._sized(b, 32)
    def blob_to_kzg_commitment(self, blob): return self.s.blob_to_kzg_commitment(blob)
    def compute_kzg_proof(self, blob, z): return self.s.compute_kzg_proof(blob, z)
    def compute_blob_kzg_proof(self, blob, c): return self.s.compute_blob_kzg_proof(blob, c)
    def verify_kzg_proof(self, c, z, y, p): return self.s.verify_kzg_proof(c, z, y, p)
    def verify_blob_kzg_proof(self, blob, c, p): return self.s.verify_blob_kzg_proof(blob, c, p)
    def verify_blob_kzg_proof_batch(self, blobs, cs, ps): return self.s.verify_blob_kzg_proof_batch(blobs, cs, ps)


def run_case(api, fn, data):
    i, want = data["input"], data["output"]
    try:
        if fn == "blob_to_kzg_commitment":
            got = "0x" + api.blob_to_kzg_commitment(api.blob(unhex(i["blob"]))).hex()
        elif fn == "compute_kzg_proof":
            p, y = api.compute_kzg_proof(api.blob(unhex(i["blob"])), api.b32(unhex(i["z"])))
            got = ["0x" + p.hex(), "0x" + y.hex()]
        elif fn == "compute_blob_kzg_proof":
            got = "0x" + api.compute_blob_kzg_proof(api.blob(unhex(i["blob"])), api.b48(unhex(i["commitment"]))).hex()
        elif fn == "verify_kzg_proof":
            got = api.verify_kzg_proof(api.b48(unhex(i["commitment"])), api.b32(unhex(i["z"])), api.b32(unhex(i["y"])), api.b48(unhex(i["proof"])))
        elif fn == "verify_blob_kzg_proof":
            got = api.verify_blob_kzg_proof(api.blob(unhex(i["blob"])), api.b48(unhex(i["commitment"])), api.b48(unhex(i["proof"])))
        elif fn == "verify_blob_kzg_proof_batch":
            got = api.verify_blob_kzg_proof_batch([api.blob(unhex(b)) for b in i["blobs"]], [api.b48(unhex(c)) for c in i["commitments"]],
                                                  [api.b48(unhex(p)) for p in i["proofs"]])
        else:
            return None
    except api.Error:
        got = None
    if isinstance(want, (list, tuple)):
        want = list(want)
    return got == want


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if not args:
        raise SystemExit(__doc__)
    root = args[0]
    setup = args[1] if len(args) > 1 else None
    api = OracleApi(setup) if "--oracle" in sys.argv else GpuApi(setup)
    total = bad = 0
    for fn_dir in sorted(glob.glob(os.path.join(root, "*"))):
        fn = os.path.basename(fn_dir)
        n = ok = 0
        for path in sorted(glob.glob(os.path.join(fn_dir, "*", "*", "data.yaml"))):
            with open(path) as fh:
                data = yaml.safe_load(fh)
            r = run_case(api, fn, data)
            if r is None:
                continue
            n += 1
            ok += bool(r)
            if not r:
                print("MISMATCH", path)
        if n:
            print("%-32s %d/%d" % (fn, ok, n))
        total += n
        bad += n - ok
    print("total %d, mismatches %d" % (total, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

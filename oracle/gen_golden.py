#!/usr/bin/env python3
"""Convert the reference's known-answer vectors into compact fixtures.

Run in the build container only (it reads /root/reference, which does not
exist on the GPU box):

    python oracle/gen_golden.py

Reads   /root/reference/tests/<fn>/small/<case>/data.yaml   (208 cases; the
        format is described by reference src/test_formats/*.rs and consumed by
        the runner at reference src/lib.rs:14-204)
        /root/reference/trusted_setup.txt   (format: reference src/kzg.rs:906-979)
Writes  tests/golden/vectors.json      one record per case, long byte strings
                                       replaced by {"ref": k} into blobs.bin
        tests/golden/blobs.bin         the distinct long byte strings, concatenated
        tests/golden/trusted_setup.bin raw 4096x48 B G1 (file order) + 65x96 B G2

The 53 MB YAML tree holds only a handful of distinct blobs, so the fixtures are
about 1.6 MB.  No reference *source* is copied, only test data.
"""
import glob
import hashlib
import json
import os
import sys

import yaml

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

FUNCS = [
    "blob_to_kzg_commitment",
    "compute_kzg_proof",
    "compute_blob_kzg_proof",
    "verify_kzg_proof",
    "verify_blob_kzg_proof",
    "verify_blob_kzg_proof_batch",
]


def main():
    os.makedirs(OUT, exist_ok=True)
    pool = bytearray()
    index = {}  # sha -> k
    table = []  # k -> [offset, length]

    def intern(hexstr):
        """Long hex strings go to blobs.bin; short ones stay inline."""
        if not isinstance(hexstr, str):
            return hexstr
        if len(hexstr) < 1000:
            return hexstr
        body = hexstr[2:] if hexstr.startswith("0x") else hexstr
        try:
            raw = bytes.fromhex(body)
        except ValueError:
            return hexstr  # malformed hex stays inline so the test sees it as-is
        key = hashlib.sha256(raw).hexdigest()
        if key not in index:
            index[key] = len(table)
            table.append([len(pool), len(raw)])
            pool.extend(raw)
        return {"ref": index[key]}

    def walk(v):
        if isinstance(v, dict):
            return {k: walk(x) for k, x in v.items()}
        if isinstance(v, list):
            return [walk(x) for x in v]
        return intern(v)

    records = []
    for fn in FUNCS:
        paths = sorted(glob.glob(os.path.join(REF, "tests", fn, "*", "*", "data.yaml")))
        assert paths, fn
        for path in paths:
            with open(path) as fh:
                doc = yaml.safe_load(fh)
            name = os.path.basename(os.path.dirname(path))
            records.append({"fn": fn, "name": name, "input": walk(doc["input"]),
                            "output": doc["output"]})
    with open(os.path.join(OUT, "vectors.json"), "w") as fh:
        json.dump({"blobs": table, "cases": records}, fh, indent=0, sort_keys=True)
    with open(os.path.join(OUT, "blobs.bin"), "wb") as fh:
        fh.write(pool)

    toks = open(os.path.join(REF, "trusted_setup.txt")).read().split()
    n1, n2 = int(toks[0]), int(toks[1])
    assert (n1, n2) == (4096, 65) and len(toks) == 2 + n1 + n2
    g1 = b"".join(bytes.fromhex(t) for t in toks[2:2 + n1])
    g2 = b"".join(bytes.fromhex(t) for t in toks[2 + n1:])
    assert len(g1) == n1 * 48 and len(g2) == n2 * 96
    with open(os.path.join(OUT, "trusted_setup.bin"), "wb") as fh:
        fh.write(g1 + g2)
    print("cases:", len(records), "distinct long strings:", len(table),
          "pool bytes:", len(pool), file=sys.stderr)


if __name__ == "__main__":
    main()

"""Loader for the fixtures written by oracle/gen_golden.py (the reference's own vectors)."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self):
        with open(os.path.join(GOLDEN, "vectors.json")) as fh:
            doc = json.load(fh)
        with open(os.path.join(GOLDEN, "blobs.bin"), "rb") as fh:
            pool = fh.read()
        self._long = [pool[o:o + n] for o, n in doc["blobs"]]
        self.cases = doc["cases"]
        with open(os.path.join(GOLDEN, "trusted_setup.bin"), "rb") as fh:
            raw = fh.read()
        self.g1_bytes = raw[:4096 * 48]
        self.g2_bytes = raw[4096 * 48:]
        assert len(self.g2_bytes) == 65 * 96

    def by_fn(self, fn):
        return [c for c in self.cases if c["fn"] == fn]

    def get_bytes(self, v):
        """hex string or {"ref": k} -> bytes; raises ValueError on malformed hex
        (the reference treats that like any other Err: src/lib.rs:42-50)."""
        if isinstance(v, dict):
            return self._long[v["ref"]]
        s = v[2:] if v.startswith("0x") else v
        return bytes.fromhex(s)

    def write_setup_text(self, path):
        """Re-create the reference's trusted_setup.txt layout (src/kzg.rs:906-979)."""
        with open(path, "w") as fh:
            fh.write("4096\n65\n")
            for i in range(4096):
                fh.write(self.g1_bytes[48 * i:48 * i + 48].hex() + "\n")
            for i in range(65):
                fh.write(self.g2_bytes[96 * i:96 * i + 96].hex() + "\n")


_G = None


def golden():
    global _G
    if _G is None:
        _G = Golden()
    return _G

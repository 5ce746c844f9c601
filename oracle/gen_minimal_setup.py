#!/usr/bin/env python3
"""Write tests/golden/trusted_setup_4.bin: a minimal-preset (FIELD_ELEMENTS_PER_BLOB = 4)
testing setup with the same secret as the reference's bundled mainnet testing setup
(tau = 1337): 4 Lagrange-basis G1 points [L_k(tau)]G1 in natural
order followed by the same 65 monomial G2 points.  The reference snapshot ships no minimal
setup (README.md:8-9 promises the preset; only src/trusted_setup.rs:143-151 hints at it),
so minimal-preset parity is oracle-vs-CUDA only ("parity unpinned")."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.pymodel import G1, R, g1_compress, g1_mul  # noqa: E402

TAU = 1337
N = 4
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

w = pow(7, (R - 1) // N, R)
roots = [pow(w, k, R) for k in range(N)]
pts = b""
for k in range(N):
    num, den = 1, 1
    for m in range(N):
        if m != k:
            num = num * (TAU - roots[m]) % R
            den = den * (roots[k] - roots[m]) % R
    pts += g1_compress(g1_mul(G1, num * pow(den, -1, R) % R))
with open(os.path.join(OUT, "trusted_setup.bin"), "rb") as fh:
    g2 = fh.read()[4096 * 48:]
with open(os.path.join(OUT, "trusted_setup_4.bin"), "wb") as fh:
    fh.write(pts + g2)
print("wrote", len(pts) + len(g2), "bytes")

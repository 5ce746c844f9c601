"""Shared helpers for the -m gpu tests: contexts over the CUDA library and oracle settings."""
import functools
import os

import numpy as np

from golden_util import GOLDEN, golden


def synthetic_blobs(count, n=4096, seed=0xB200):
    """Deterministic stand-in for the reference's bench generator (benches/kzg_benches.rs:14-23):
    uniform random bytes with the top byte of every field element zeroed."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    a = rng.integers(0, 256, size=(count, n, 32), dtype=np.uint8)
    a[:, :, 0] = 0
    return a.reshape(count, n * 32)


def minimal_setup_bytes():
    with open(os.path.join(GOLDEN, "trusted_setup_4.bin"), "rb") as fh:
        raw = fh.read()
    return raw[:4 * 48], raw[4 * 48:]


# The reference's vectors run twice: on a small comb (6 MB of table) and on the configuration bench.py measures
# (comb_width 0 = automatic: 23 on an empty B200, 72 GB of table, 4096-blob chunks).
VECTOR_WIDTHS = [8, 0]
VECTOR_WIDTH_IDS = ["g8", "auto"]


@functools.lru_cache(maxsize=None)
def gpu_settings(preset="mainnet", comb_width=8):
    """One context per (preset, comb width), kept for the whole test session.  Explicit widths get a small
    workspace (128-blob chunks) so that several contexts fit beside the automatic one."""
    from kzg_rust_b200 import KzgSettings
    old = os.environ.get("KZG_B200_CHUNK")
    if comb_width != 0:
        os.environ["KZG_B200_CHUNK"] = "128"
    try:
        if preset == "mainnet":
            g = golden()
            return KzgSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes, 0, comb_width)
        g1, g2 = minimal_setup_bytes()
        return KzgSettings.load_trusted_setup(g1, g2, 0, comb_width)
    finally:
        if comb_width != 0:
            if old is None:
                del os.environ["KZG_B200_CHUNK"]
            else:
                os.environ["KZG_B200_CHUNK"] = old


@functools.lru_cache(maxsize=None)
def oracle_settings(preset="mainnet"):
    from oracle.binding import OracleSettings
    if preset == "mainnet":
        g = golden()
        return OracleSettings.load_trusted_setup(g.g1_bytes, g.g2_bytes)
    g1, g2 = minimal_setup_bytes()
    return OracleSettings.load_trusted_setup(g1, g2)

"""-m gpu: the device field arithmetic itself (through kzg_b200_debug_field_op) against Python integers:
the canonical Montgomery operations and the lazy [0, 2p) forms the MSM levels use -- all bit-exact."""
import ctypes

import numpy as np
import pytest

from gpu_util import gpu_settings

pytestmark = pytest.mark.gpu
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _pack(vals, n):
    a = np.zeros((len(vals), n), dtype=np.uint32)
    for r, v in enumerate(vals):
        for i in range(n):
            a[r, i] = (v >> (32 * i)) & 0xffffffff
    return a


def _unpack(a):
    return [sum(int(x) << (32 * i) for i, x in enumerate(row)) for row in a]


def _run(op, av, bv, words):
    import kzg_rust_b200 as k
    L = k.load_library()
    s = gpu_settings("mainnet", 8)
    L.kzg_b200_debug_field_op.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    a, b = _pack(av, words), _pack(bv, words)
    out = np.zeros_like(a)
    rc = L.kzg_b200_debug_field_op(s._h, op, a.ctypes.data, b.ctypes.data, out.ctypes.data, len(av))
    assert rc == 0
    return _unpack(out)


def _values(mod, bound, count, seed):
    rng = np.random.default_rng(seed)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, 2 ** 48 - 1, 2 ** 48, 2 ** 96 - 1, 0xffffffff, 2 ** 336 % mod]
    if bound > mod:
        edge += [mod, mod + 1, bound - 1, bound - 2]
    edge = [e for e in edge if e < bound]
    rand = [int.from_bytes(rng.bytes(64), "big") % bound for _ in range(count)]
    av = rand + [e for e in edge for _ in edge]
    bv = rand[::-1] + [e for _ in edge for e in edge]
    return av, bv


def test_canonical_fp_and_fr_ops():
    Rp_inv = pow(pow(2, 384, P), -1, P)
    av, bv = _values(P, P, 3000, 1)
    assert _run(0, av, bv, 12) == [x * y * Rp_inv % P for x, y in zip(av, bv)]
    assert _run(3, av, bv, 12) == [(x + y) % P for x, y in zip(av, bv)]
    assert _run(4, av, bv, 12) == [(x - y) % P for x, y in zip(av, bv)]
    nz = [x for x in av if x][:400]
    Rp = pow(2, 384, P)
    assert _run(1, nz, nz, 12) == [pow(x * Rp_inv % P, -1, P) * Rp % P for x in nz]  # Montgomery in, Montgomery out
    Rr_inv = pow(pow(2, 256, R), -1, R)
    av, bv = _values(R, R, 3000, 2)
    assert _run(2, av, bv, 8) == [x * y * Rr_inv % R for x, y in zip(av, bv)]


def test_lazy_fp_ops_stay_below_2p():
    Rp_inv = pow(pow(2, 384, P), -1, P)
    av, bv = _values(P, 2 * P, 3000, 3)
    got = _run(7, av, bv, 12)
    assert all(g < 2 * P for g in got)
    assert [g % P for g in got] == [x * y * Rp_inv % P for x, y in zip(av, bv)]
    got = _run(8, av, bv, 12)
    assert all(g < 2 * P for g in got)
    assert [g % P for g in got] == [(x - y) % P for x, y in zip(av, bv)]

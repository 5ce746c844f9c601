"""Pure-Python big-int model of BLS12-381 + the EIP-4844 blob path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
file; it is used by tests/ (small cases), by tools/gen_consts.py (to derive
Montgomery constants) and to cross-check oracle/kzg_oracle.c.

It restates, with Python integers:
  * Fp / Fr arithmetic, G1 affine group law, ZCash point (de)compression as
    done by blst for the calls in reference src/utils.rs:221-227, 282-315;
  * Fp2 / Fp12 and an optimal-ate pairing (slow, generic) standing in for
    `pairings_verify` (reference src/utils.rs:189-214);
  * the KZG blob functions of reference src/kzg.rs:282-693.

The arithmetic itself lives in blst 0.3.11 (Cargo.lock:44-47), which is not
vendored under /root/reference; the restatement follows the published curve
and serialisation definitions and is pinned by the reference's vectors.
"""
import hashlib

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
BLS_X = 0xd201000000010000  # |x|; the curve parameter is -BLS_X
G1X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
G2X = (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
       0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e)
G2Y = (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
       0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)
G1 = (G1X, G1Y)
G2 = (G2X, G2Y)

FIELD_ELEMENTS_PER_BLOB = 4096
BYTES_PER_BLOB = 32 * FIELD_ELEMENTS_PER_BLOB
FIAT_SHAMIR_PROTOCOL_DOMAIN = b"FSBLOBVERIFY_V1_"
RANDOM_CHALLENGE_KZG_BATCH_DOMAIN = b"RCKZGBATCH___V1_"


class KzgError(Exception):
    pass


# ----------------------------------------------------------------- G1 (affine; None = infinity)
def g1_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - 4) % P == 0


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_mul(pt, k):
    k %= R
    res = None
    add = pt
    while k:
        if k & 1:
            res = g1_add(res, add)
        add = g1_add(add, add)
        k >>= 1
    return res


def g1_mul_raw(pt, k):
    """Multiply by an arbitrary non-negative integer (no reduction mod r)."""
    res = None
    add = pt
    while k:
        if k & 1:
            res = g1_add(res, add)
        add = g1_add(add, add)
        k >>= 1
    return res


def g1_in_subgroup(pt):
    return g1_mul_raw(pt, R) is None


def g1_compress(pt):
    """blst_p1_compress semantics (reference src/utils.rs:221-227)."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def fp_sqrt(a):
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


def g1_uncompress(b):
    """blst_p1_uncompress semantics; raises KzgError on any malformed input."""
    if len(b) != 48:
        raise KzgError("length")
    if not b[0] & 0x80:
        raise KzgError("uncompressed flag")  # blst: BLST_BAD_ENCODING
    if b[0] & 0x40:
        if b[0] & 0x3F or any(b[1:]):
            raise KzgError("bad infinity encoding")
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= P:
        raise KzgError("x >= p")
    y = fp_sqrt((x * x * x + 4) % P)
    if y is None:
        raise KzgError("not on curve")
    if (y > (P - 1) // 2) != bool(b[0] & 0x20):
        y = P - y
    return (x, y)


def validate_kzg_g1(b):
    """reference src/utils.rs:282-315."""
    pt = g1_uncompress(b)
    if pt is None:
        return None
    if not g1_in_subgroup(pt):
        raise KzgError("not in G1")
    return pt


# ----------------------------------------------------------------- Fp2 / G2
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)


def f2_pow(a, e):
    res = (1, 0)
    while e:
        if e & 1:
            res = f2_mul(res, a)
        a = f2_mul(a, a)
        e >>= 1
    return res


def f2_sqrt(a):
    """Square root in Fp2 = Fp[u]/(u^2+1), p = 3 mod 4 (complex method)."""
    if a == (0, 0):
        return (0, 0)
    a1 = f2_pow(a, (P - 3) // 4)
    alpha = f2_mul(f2_mul(a1, a1), a)
    x0 = f2_mul(a1, a)
    if alpha == (P - 1, 0):
        cand = f2_mul((0, 1), x0)
    else:
        b = f2_pow(f2_add((1, 0), alpha), (P - 1) // 2)
        cand = f2_mul(b, x0)
    return cand if f2_mul(cand, cand) == (a[0] % P, a[1] % P) else None


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_add(y1, y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(pt, k):
    res = None
    while k:
        if k & 1:
            res = g2_add(res, pt)
        pt = g2_add(pt, pt)
        k >>= 1
    return res


def g2_neg(pt):
    return None if pt is None else (pt[0], ((-pt[1][0]) % P, (-pt[1][1]) % P))


def g2_uncompress(b):
    """ZCash G2 format: 96 bytes, x.c1 first then x.c0; flags in byte 0."""
    if len(b) != 96:
        raise KzgError("length")
    if not b[0] & 0x80:
        raise KzgError("uncompressed flag")
    if b[0] & 0x40:
        if b[0] & 0x3F or any(b[1:]):
            raise KzgError("bad infinity encoding")
        return None
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    if x0 >= P or x1 >= P:
        raise KzgError("x >= p")
    x = (x0, x1)
    rhs = f2_add(f2_mul(f2_mul(x, x), x), (4, 4))
    y = f2_sqrt(rhs)
    if y is None:
        raise KzgError("not on curve")
    # sign: lexicographic on (c1, c0)
    big = (y[1] > (P - 1) // 2) if y[1] != 0 else (y[0] > (P - 1) // 2)
    if big != bool(b[0] & 0x20):
        y = ((-y[0]) % P, (-y[1]) % P)
    return (x, y)


# ----------------------------------------------------------------- Fp12 as Fp[w]/(w^12 - 2 w^6 + 2)
# (u = w^6 - 1 satisfies u^2 = -1; xi = u + 1 = w^6.)
def f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for k in range(22, 11, -1):
        c = t[k]
        if c:
            t[k - 6] += 2 * c
            t[k - 12] -= 2 * c
    return [v % P for v in t[:12]]


F12_ONE = [1] + [0] * 11


def f12_pow(a, e):
    res = F12_ONE
    while e:
        if e & 1:
            res = f12_mul(res, a)
        a = f12_mul(a, a)
        e >>= 1
    return res


def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def f12_inv(a):
    """Extended Euclid on polynomials over Fp."""
    mod = [2, 0, 0, 0, 0, 0, P - 2, 0, 0, 0, 0, 0, 1]
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], mod
    while _poly_deg(low):
        # r = high / low
        dl, dh = _poly_deg(low), _poly_deg(high)
        q = [0] * 13
        temp = list(high)
        inv_lead = pow(low[dl], -1, P)
        for i in range(dh - dl, -1, -1):
            q[i] = temp[dl + i] * inv_lead % P
            for c in range(dl + 1):
                temp[c + i] = (temp[c + i] - q[i] * low[c]) % P
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * q[j]) % P
                new[i + j] = (new[i + j] - low[i] * q[j]) % P
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, P)
    return [c * inv0 % P for c in lm[:12]]


def _f2_to_f12(c):
    """a + b u  ->  (a - b) + b w^6."""
    out = [0] * 12
    out[0] = (c[0] - c[1]) % P
    out[6] = c[1] % P
    return out


_W2 = [0, 0, 1] + [0] * 9
_W3 = [0, 0, 0, 1] + [0] * 8
_W2_INV = None
_W3_INV = None


def _twist(q):
    """Map a G2 point on E'(Fp2): y^2 = x^3 + 4(u+1) to E(Fp12): y^2 = x^3 + 4."""
    global _W2_INV, _W3_INV
    if _W2_INV is None:
        _W2_INV = f12_inv(_W2)
        _W3_INV = f12_inv(_W3)
    x, y = q
    return (f12_mul(_f2_to_f12(x), _W2_INV), f12_mul(_f2_to_f12(y), _W3_INV))


def _f12_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def _f12_sub(a, b):
    return [(x - y) % P for x, y in zip(a, b)]


def _f12_scalar(a, k):
    return [x * k % P for x in a]


def _e12_double(pt):
    x, y = pt
    lam = f12_mul(_f12_scalar(f12_mul(x, x), 3), f12_inv(_f12_scalar(y, 2)))
    nx = _f12_sub(f12_mul(lam, lam), _f12_scalar(x, 2))
    ny = _f12_sub(f12_mul(lam, _f12_sub(x, nx)), y)
    return (nx, ny), lam


def _e12_add(a, b):
    x1, y1 = a
    x2, y2 = b
    lam = f12_mul(_f12_sub(y2, y1), f12_inv(_f12_sub(x2, x1)))
    nx = _f12_sub(_f12_sub(f12_mul(lam, lam), x1), x2)
    ny = _f12_sub(f12_mul(lam, _f12_sub(x1, nx)), y1)
    return (nx, ny), lam


def miller_loop(q, p):
    """f_{|x|,Q}(P) over Fp12 (generic affine formulas; test-only speed)."""
    if q is None or p is None:
        return list(F12_ONE)
    Q = _twist(q)
    px = [p[0]] + [0] * 11
    py = [p[1]] + [0] * 11
    T = Q
    f = list(F12_ONE)
    for bit in bin(BLS_X)[3:]:
        T2, lam = _e12_double(T)
        line = _f12_sub(f12_mul(lam, _f12_sub(px, T[0])), _f12_sub(py, T[1]))
        f = f12_mul(f12_mul(f, f), line)
        T = T2
        if bit == "1":
            T2, lam = _e12_add(T, Q)
            line = _f12_sub(f12_mul(lam, _f12_sub(px, T[0])), _f12_sub(py, T[1]))
            f = f12_mul(f, line)
            T = T2
    return f


def final_exp(f):
    return f12_pow(f, (P ** 12 - 1) // R)


def pairings_verify(a1, a2, b1, b2):
    """e(a1, a2) == e(b1, b2)  (reference src/utils.rs:189-214)."""
    f = f12_mul(miller_loop(a2, g1_neg(a1)), miller_loop(b2, b1))
    return final_exp(f) == F12_ONE


# ----------------------------------------------------------------- setup
def reverse_bits(n, order):
    bits = order.bit_length() - 1
    return int(format(n, "0%db" % bits)[::-1], 2) if bits else 0


def bit_reversal_permutation(vals):
    n = len(vals)
    return [vals[reverse_bits(i, n)] for i in range(n)]


def compute_roots_of_unity(n=FIELD_ELEMENTS_PER_BLOB):
    """reference src/kzg.rs:764-799: w = 7^((r-1)/n), powers, bit-reversed."""
    w = pow(7, (R - 1) // n, R)
    roots = [pow(w, i, R) for i in range(n)]
    return bit_reversal_permutation(roots)


class Settings:
    """Python stand-in for KzgSettings (reference src/kzg.rs:27-40)."""

    def __init__(self, g1_bytes, g2_bytes, n=FIELD_ELEMENTS_PER_BLOB, decode_g1=True):
        self.n = n
        self.roots = compute_roots_of_unity(n)
        self.g1_bytes = bit_reversal_permutation(list(g1_bytes))
        self.g1 = [g1_uncompress(b) for b in self.g1_bytes] if decode_g1 else None
        self.g2 = [g2_uncompress(b) for b in g2_bytes[:2]]


# ----------------------------------------------------------------- KZG blob path
def bytes_to_bls_field(b):
    v = int.from_bytes(b, "big")
    if v >= R:
        raise KzgError("non-canonical field element")
    return v


def blob_to_polynomial(blob, n=FIELD_ELEMENTS_PER_BLOB):
    if len(blob) != 32 * n:
        raise KzgError("blob length")
    return [bytes_to_bls_field(blob[32 * i:32 * i + 32]) for i in range(n)]


def compute_challenge(blob, commitment_bytes, n=FIELD_ELEMENTS_PER_BLOB):
    validate_kzg_g1(commitment_bytes)
    data = (FIAT_SHAMIR_PROTOCOL_DOMAIN + (0).to_bytes(8, "big") + n.to_bytes(8, "big")
            + blob + commitment_bytes)
    return int.from_bytes(hashlib.sha256(data).digest(), "big") % R


def evaluate_polynomial_in_evaluation_form(poly, z, roots):
    n = len(poly)
    for i in range(n):
        if z == roots[i]:
            return poly[i]
    acc = 0
    for i in range(n):
        acc += poly[i] * roots[i] % R * pow(z - roots[i], -1, R)
    return acc % R * pow(n, -1, R) % R * (pow(z, n, R) - 1) % R


def quotient_evals(poly, z, y, roots):
    """reference src/kzg.rs:470-523."""
    n = len(poly)
    q = [0] * n
    m = None
    for i in range(n):
        if z == roots[i]:
            m = i
            continue
        q[i] = (poly[i] - y) * pow(roots[i] - z, -1, R) % R
    if m is not None:
        acc = 0
        for i in range(n):
            if i == m:
                continue
            acc += (poly[i] - y) * roots[i] % R * pow(z * (z - roots[i]) % R, -1, R)
        q[m] = acc % R
    return q


def g1_lincomb(points, scalars):
    acc = None
    for pt, s in zip(points, scalars):
        if s and pt is not None:
            acc = g1_add(acc, g1_mul(pt, s))
    return acc

// sha256.cuh -- FIPS 180-4 SHA-256 for the two Fiat-Shamir hashes of the blob path:
// `compute_challenge` (reference src/kzg.rs:298-339, device: one hash stream per blob) and
// `compute_r_powers` (reference src/utils.rs:426-474, host: one short sequential hash per
// batch).  The reference calls blst_sha256 for both.  Words are big-endian message words
// already converted to native integers.
#pragma once
#include "bigint.cuh"

namespace kzg {

#define KZG_SHA256_K                                                                                    \
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, \
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, \
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, \
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, \
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, \
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, \
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, \
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u

KZG_HD uint32_t sha_rotr(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

KZG_HD void sha256_init(uint32_t h[8]) {
    h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
    h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
}

// x + y as x * one + y with a run-time one = 1: on the device that is an IMAD, which issues on the FMA pipe, where a plain
// addition would queue on the 16-lane integer pipe behind the rotations and boolean functions that bound a hash stream.
KZG_HD uint32_t sha_fadd(uint32_t x, uint32_t y, uint32_t one) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(one), "r"(y));
    return r;
#else
    (void)one;
    return x + y;
#endif
}

// One compression.  w[16] is clobbered (it is the rolling message schedule).  Fully
// unrolled so that w[] and the round constants stay in registers / immediates.
// (With one stream per thread the plain additions are the faster form: written as IMADs the compression of 16,384 streams
// took 4.18 ms instead of 3.80 ms.  The grouped kernel, where a scheduler carries one warp, gains from them: frpath.cuh.)
KZG_HD void sha256_compress(uint32_t h[8], uint32_t w[16]) {
    constexpr uint32_t K[64] = {KZG_SHA256_K};
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i];
        } else {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
            w[i & 15] = wi;
        }
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K[i] + wi;
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// Incremental hasher over a byte stream (host side: compute_r_powers; device side: tails).
struct Sha256 {
    uint32_t h[8];
    uint32_t w[16];
    uint64_t len;   // bytes absorbed
    KZG_HD void init() {
        sha256_init(h);
        len = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = 0;
    }
    KZG_HD void put_byte(uint8_t v) {
        uint32_t pos = (uint32_t)(len & 63);
        uint32_t wi = pos >> 2, sh = 24 - 8 * (pos & 3);
        if ((pos & 3) == 0) w[wi] = 0;
        w[wi] |= (uint32_t)v << sh;
        len++;
        if (pos == 63) {
            uint32_t t[16];
            for (int i = 0; i < 16; i++) t[i] = w[i];
            sha256_compress(h, t);
        }
    }
    KZG_HD void update(const uint8_t *p, size_t n) {
        size_t i = 0;
        while (i < n && (len & 63) != 0) put_byte(p[i++]);
        for (; i + 64 <= n; i += 64) {  // whole blocks straight from the input
            uint32_t t[16];
            for (int k = 0; k < 16; k++) {
                const uint8_t *q = p + i + 4 * k;
                t[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
            }
            sha256_compress(h, t);
            len += 64;
        }
        for (; i < n; i++) put_byte(p[i]);
    }
    KZG_HD void finish(uint8_t out[32]) {
        uint64_t bits = len * 8;
        put_byte(0x80);
        while ((len & 63) != 56) put_byte(0);
        for (int i = 7; i >= 0; i--) put_byte((uint8_t)(bits >> (8 * i)));
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
            out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
        }
    }
};

}  // namespace kzg

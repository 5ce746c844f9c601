// host_sha256.cpp -- see host_sha256.h
#include "host_sha256.h"

#include <string.h>

#include "sha256.cuh"

#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#define KZG_HAVE_SHANI 1
#endif

static void compress_portable(uint32_t h[8], const uint8_t *p, size_t blocks) {
    for (size_t b = 0; b < blocks; b++, p += 64) {
        uint32_t w[16];
        for (int k = 0; k < 16; k++)
            w[k] = ((uint32_t)p[4 * k] << 24) | ((uint32_t)p[4 * k + 1] << 16) | ((uint32_t)p[4 * k + 2] << 8) | p[4 * k + 3];
        kzg::sha256_compress(h, w);
    }
}

#ifdef KZG_HAVE_SHANI
// Two rounds per SHA256RNDS2; the state lives in two registers as (A B E F) and (C D G H).  Four rounds per step:
// add the round constants to four schedule words, run 2 x 2 rounds, and advance the message schedule with
// SHA256MSG1 / SHA256MSG2.
__attribute__((target("sha,sse4.1,ssse3"))) static void compress_shani(uint32_t h[8], const uint8_t *p, size_t blocks) {
    static const uint32_t K[64] = {KZG_SHA256_K};
    const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i tmp = _mm_loadu_si128((const __m128i *)&h[0]);     // D C B A (high .. low = h[3] .. h[0])
    __m128i state1 = _mm_loadu_si128((const __m128i *)&h[4]);  // H G F E
    tmp = _mm_shuffle_epi32(tmp, 0xB1);                        // C D A B
    state1 = _mm_shuffle_epi32(state1, 0x1B);                  // E F G H
    __m128i state0 = _mm_alignr_epi8(tmp, state1, 8);          // A B E F
    state1 = _mm_blend_epi16(state1, tmp, 0xF0);               // C D G H
    for (size_t b = 0; b < blocks; b++, p += 64) {
        const __m128i save0 = state0, save1 = state1;
        __m128i m[4];
        for (int i = 0; i < 4; i++) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 16 * i)), bswap);
        for (int i = 0; i < 16; i++) {
            // m[i & 3] holds W[4i .. 4i+3]
            __m128i msg = _mm_add_epi32(m[i & 3], _mm_loadu_si128((const __m128i *)&K[4 * i]));
            state1 = _mm_sha256rnds2_epu32(state1, state0, msg);
            msg = _mm_shuffle_epi32(msg, 0x0E);
            state0 = _mm_sha256rnds2_epu32(state0, state1, msg);
            if (i < 12) {
                // W[4(i+4) .. ] = msg2(msg1(W[4i..], W[4i+4..]) + (W[4i+9 .. 4i+12]), W[4i+12 ..])
                __m128i t = _mm_sha256msg1_epu32(m[i & 3], m[(i + 1) & 3]);
                t = _mm_add_epi32(t, _mm_alignr_epi8(m[(i + 3) & 3], m[(i + 2) & 3], 4));
                m[i & 3] = _mm_sha256msg2_epu32(t, m[(i + 3) & 3]);
            }
        }
        state0 = _mm_add_epi32(state0, save0);
        state1 = _mm_add_epi32(state1, save1);
    }
    tmp = _mm_shuffle_epi32(state0, 0x1B);       // F E B A
    state1 = _mm_shuffle_epi32(state1, 0xB1);    // D C H G
    state0 = _mm_blend_epi16(tmp, state1, 0xF0); // D C B A
    state1 = _mm_alignr_epi8(state1, tmp, 8);    // H G F E
    _mm_storeu_si128((__m128i *)&h[0], state0);
    _mm_storeu_si128((__m128i *)&h[4], state1);
}
static bool cpu_has_shani() {
    unsigned a, b, c, d;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
    const bool sha = (b >> 29) & 1;
    if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
    const bool sse41 = (c >> 19) & 1, ssse3 = (c >> 9) & 1;
    return sha && sse41 && ssse3;
}
#endif

typedef void (*compress_fn)(uint32_t *, const uint8_t *, size_t);
static compress_fn pick() {
#ifdef KZG_HAVE_SHANI
    if (cpu_has_shani()) return compress_shani;
#endif
    return compress_portable;
}
static const compress_fn g_compress = pick();

bool host_sha256_accelerated() { return g_compress != compress_portable; }

void HostSha256::init() {
    kzg::sha256_init(h);
    len = 0;
}
void HostSha256::update(const uint8_t *p, size_t n) {
    size_t fill = (size_t)(len & 63);
    len += n;
    if (fill) {
        size_t take = 64 - fill < n ? 64 - fill : n;
        memcpy(buf + fill, p, take);
        p += take;
        n -= take;
        if (fill + take < 64) return;
        g_compress(h, buf, 1);
    }
    if (n >= 64) {
        g_compress(h, p, n / 64);
        p += n & ~(size_t)63;
        n &= 63;
    }
    if (n) memcpy(buf, p, n);
}
void HostSha256::finish(uint8_t out[32]) {
    const uint64_t bits = len * 8;
    size_t fill = (size_t)(len & 63);
    buf[fill++] = 0x80;
    if (fill > 56) {
        memset(buf + fill, 0, 64 - fill);
        g_compress(h, buf, 1);
        fill = 0;
    }
    memset(buf + fill, 0, 56 - fill);
    for (int i = 0; i < 8; i++) buf[56 + i] = (uint8_t)(bits >> (8 * (7 - i)));
    g_compress(h, buf, 1);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
    }
}

// test hook (tests/test_host_logic.py): one-shot hash with the dispatched or the portable compression function
extern "C" void kzg_b200_host_sha256(const uint8_t *p, size_t n, int portable, uint8_t out[32]) {
    if (!portable) {
        HostSha256 s;
        s.init();
        // odd-sized pieces exercise the buffering
        size_t off = 0, step = 1;
        while (off < n) {
            size_t take = step < n - off ? step : n - off;
            s.update(p + off, take);
            off += take;
            step = step * 3 + 1;
        }
        s.finish(out);
        return;
    }
    kzg::Sha256 s;
    s.init();
    s.update(p, n);
    s.finish(out);
}

// frpath.cuh -- scalar-field (Fr) side of the blob path: roots of unity, the Fiat-Shamir
// challenge, barycentric evaluation, the quotient polynomial and the r-power linear
// combinations of batch verification.
//
//   k_challenge        reference compute_challenge              src/kzg.rs:298-339
//   k_eval_quotient    evaluate_polynomial_in_evaluation_form   src/kzg.rs:346-389
//                      + the quotient of compute_kzg_proof_impl src/kzg.rs:461-523
//                      (both share one batch inversion: 1/(w_i - z) = -1/(z - w_i))
//   k_load_scalars     bytes_to_bls_field                       src/utils.rs:262-275
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "blobpath.cuh"
#include "sha256.cuh"

namespace kzg {

// reference compute_roots_of_unity (src/kzg.rs:764-799) + expand_root_of_unity (:734-761):
// w = 7^((r-1)/n); [w^0 .. w^(n-1)] in bit-reversed order, Montgomery form, on the device.
inline int fr_setup_roots(int n, fr_t **d_roots, cudaStream_t stream) {
    fr_t w;
    {
        const uint32_t w4096[8] = {FR_ROOT_OF_UNITY_4096_LIMBS}, w4[8] = {FR_ROOT_OF_UNITY_4_LIMBS};
        for (int i = 0; i < 8; i++) w.l[i] = n == 4096 ? w4096[i] : w4[i];
    }
    fe_to_mont(w, w);
    std::vector<fr_t> nat(n), rev(n);
    nat[0] = fe_one<FrParams>();
    for (int i = 1; i < n; i++) fe_mul(nat[i], nat[i - 1], w);
    fr_t chk;
    fe_mul(chk, nat[n - 1], w);
    if (!fe_eq(chk, nat[0])) return KZG_INTERNAL;  // w^n must be 1
    for (int i = 0; i < n; i++) {
        uint32_t r = 0, v = (uint32_t)i;
        for (int o = n; o > 1; o >>= 1) { r = (r << 1) | (v & 1); v >>= 1; }
        rev[i] = nat[r];
    }
    if (cudaMalloc(d_roots, (size_t)n * sizeof(fr_t)) != cudaSuccess) return KZG_CUDA;
    if (cudaMemcpyAsync(*d_roots, rev.data(), (size_t)n * sizeof(fr_t), cudaMemcpyHostToDevice, stream) != cudaSuccess) return KZG_CUDA;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return KZG_CUDA;
    return KZG_OK;
}

#if defined(__CUDACC__)

KZG_D fr_t fr_inv_n_mont(int n) {
    fr_t r;
    constexpr uint32_t a[8] = {FR_INV_N_MONT_4096_LIMBS}, b[8] = {FR_INV_N_MONT_4_LIMBS};
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = n == 4096 ? a[i] : b[i];
    return r;
}
KZG_D void ld_fr(fr_t &r, const fr_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
}
KZG_D void st_fr(fr_t *p, const fr_t &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.l[0], r.l[1], r.l[2], r.l[3]);
    q[1] = make_uint4(r.l[4], r.l[5], r.l[6], r.l[7]);
}
// 32 bytes -> 8 big-endian message words
KZG_D void ld_be_words8(uint32_t *w, const uint8_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    w[0] = __byte_perm(a.x, 0, 0x0123); w[1] = __byte_perm(a.y, 0, 0x0123);
    w[2] = __byte_perm(a.z, 0, 0x0123); w[3] = __byte_perm(a.w, 0, 0x0123);
    w[4] = __byte_perm(b.x, 0, 0x0123); w[5] = __byte_perm(b.y, 0, 0x0123);
    w[6] = __byte_perm(b.z, 0, 0x0123); w[7] = __byte_perm(b.w, 0, 0x0123);
}
KZG_D void st_scalar_be32(uint8_t *p, const fr_t &canon) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(__byte_perm(canon.l[7], 0, 0x0123), __byte_perm(canon.l[6], 0, 0x0123),
                      __byte_perm(canon.l[5], 0, 0x0123), __byte_perm(canon.l[4], 0, 0x0123));
    q[1] = make_uint4(__byte_perm(canon.l[3], 0, 0x0123), __byte_perm(canon.l[2], 0, 0x0123),
                      __byte_perm(canon.l[1], 0, 0x0123), __byte_perm(canon.l[0], 0, 0x0123));
}

// ------------------------------------------------------------------ Fiat-Shamir challenge
// One hash stream per blob: SHA-256("FSBLOBVERIFY_V1_" || u64be(0) || u64be(n) || blob ||
// commitment) reduced mod r.  The message is 32 + 32 n + 48 bytes: block 0 carries the
// 32-byte header and element 0, every further block two elements, the one after the last
// element the first 32 commitment bytes, and the final block the last 16 commitment bytes
// plus the padding.  The next block is loaded while the current one is compressed.
// z_out: canonical little-endian limbs.
__global__ void __launch_bounds__(64) k_challenge(const uint8_t *__restrict__ blobs, const uint8_t *__restrict__ commitments,
                                                  uint32_t count, int n, fr_t *__restrict__ z_out) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    const uint8_t *blob = blobs + (size_t)b * n * 32;
    const uint8_t *cm = commitments + (size_t)b * 48;
    uint32_t h[8], w[16], nx[16];
    sha256_init(h);
    w[0] = 0x4653424cu; w[1] = 0x4f425645u; w[2] = 0x52494659u; w[3] = 0x5f56315fu;  // "FSBLOBVERIFY_V1_"
    w[4] = 0; w[5] = 0; w[6] = 0; w[7] = (uint32_t)n;
    ld_be_words8(w + 8, blob);
    const int pairs = n / 2;  // blocks 1 .. pairs-1 hold elements (2k-1, 2k)
#pragma unroll 1
    for (int k = 1; k < pairs; k++) {
        ld_be_words8(nx, blob + 64 * (size_t)k - 32);
        ld_be_words8(nx + 8, blob + 64 * (size_t)k);
        sha256_compress(h, w);
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = nx[i];
    }
    ld_be_words8(nx, blob + 32 * (size_t)(n - 1));
    ld_be_words8(nx + 8, cm);
    sha256_compress(h, w);
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = nx[i];
    sha256_compress(h, w);
    {
        const uint4 t = *reinterpret_cast<const uint4 *>(cm + 32);
        w[0] = __byte_perm(t.x, 0, 0x0123); w[1] = __byte_perm(t.y, 0, 0x0123);
        w[2] = __byte_perm(t.z, 0, 0x0123); w[3] = __byte_perm(t.w, 0, 0x0123);
        w[4] = 0x80000000u;
#pragma unroll
        for (int i = 5; i < 14; i++) w[i] = 0;
        uint64_t bits = (uint64_t)(32 + 32 * (uint64_t)n + 48) * 8;
        w[14] = (uint32_t)(bits >> 32);
        w[15] = (uint32_t)bits;
    }
    sha256_compress(h, w);
    fr_t z;
#pragma unroll
    for (int i = 0; i < 8; i++) z.l[i] = h[7 - i];
    scalar_reduce(z);
    st_fr(z_out + b, z);
}

// The same hash with G LANES per blob (G = 32, 16, 8, 4, 2), for batches that cannot fill the GPU with one stream per
// thread.  A hash stream is sequential, and a warp -- whether it carries 32 streams or one -- issues one instruction
// every 2 cycles (the 16-lane integer pipe of its scheduler): ~1000 instructions per compression for the 64 rounds and
// ~600 for the message schedule, 3.1 - 4 ms per blob for any batch up to one warp per scheduler (592 warps = 18,944
// streams).  The schedule does not depend on the running state: here the G lanes of a blob expand the schedules of G
// consecutive blocks at once (W_t + K_t for the 64 rounds, through shared memory), then every lane runs only the rounds
// of those G blocks in turn (the G lanes of a blob redundantly): 1024 + 600 / G instructions per compression, i.e.
// 1.9 ms per blob at G = 32 (one warp per blob, up to 592 blobs) ... 2.4 ms at G = 4 (8 blobs per warp, up to 4,736).
// fr_launch_challenge picks the largest G that keeps the launch at one warp per scheduler.
#define KZG_CHALLENGE_GROUP_WARPS 592
// big-endian message word j of the 64-byte block blk of  "FSBLOBVERIFY_V1_" || u64be(0) || u64be(n) || blob || commitment || padding
KZG_D uint32_t challenge_msg_word(const uint8_t *__restrict__ blob, const uint8_t *__restrict__ cm, int n, uint32_t blk, int j) {
    const uint32_t m = 64u * blk + 4u * (uint32_t)j, blob_end = 32u + 32u * (uint32_t)n, msg_end = blob_end + 48u;
    if (m < 32u) {
        const uint32_t hdr[8] = {0x4653424cu, 0x4f425645u, 0x52494659u, 0x5f56315fu, 0u, 0u, 0u, (uint32_t)n};
        return hdr[m >> 2];
    }
    if (m < blob_end) return __byte_perm(*reinterpret_cast<const uint32_t *>(blob + (m - 32u)), 0, 0x0123);
    if (m < msg_end) return __byte_perm(*reinterpret_cast<const uint32_t *>(cm + (m - blob_end)), 0, 0x0123);
    if (m == msg_end) return 0x80000000u;
    const uint32_t nblocks = (msg_end + 9u + 63u) / 64u;
    if (blk == nblocks - 1 && j == 15) return msg_end * 8u;  // the bit length fits 32 bits
    return 0u;
}
// one block of four warps; kw: 4 x 64 x 32 words of shared memory
// blk_begin / blk_end: the 64-byte blocks of the message this launch hashes (multiples of 32 blocks, or the end of the message).
// A launch that does not start at block 0 takes the running state of blob b from state[b], one that does not reach the end
// leaves it there: the hash of a chunk whose blobs are still being uploaded runs slice by slice behind the copies
// (verify_chunk_a, proof_verify.inl).
template <int G>
KZG_D void challenge_group_block(uint32_t (*kw)[64 * 32], uint32_t block, const uint8_t *__restrict__ blobs,
                                 const uint8_t *__restrict__ commitments, uint32_t count, int n, fr_t *__restrict__ z_out, uint32_t one,
                                 uint32_t blk_begin, uint32_t blk_end, uint4 *__restrict__ state) {
    constexpr uint32_t K[64] = {KZG_SHA256_K};
    auto fadd = [one](uint32_t x, uint32_t y) -> uint32_t { return sha_fadd(x, y, one); };  // an IMAD (FMA pipe), see sha256.cuh
    constexpr uint32_t PER_WARP = 32 / G;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, slot = lane / G, sub = lane % G;
    if ((block * 4 + warp) * PER_WARP >= count) return;  // the whole warp
    // a warp's last slots may be past the end: they hash the last blob again and write nothing
    const uint32_t b = min((block * 4 + warp) * PER_WARP + slot, count - 1);
    const bool live = (block * 4 + warp) * PER_WARP + slot < count;
    const uint8_t *blob = blobs + (size_t)b * n * 32;
    const uint8_t *cm = commitments + (size_t)b * 48;
    const uint32_t msg_end = 32u + 32u * (uint32_t)n + 48u, nblocks = (msg_end + 9u + 63u) / 64u;
    uint32_t h[8];
    blk_end = min(blk_end, nblocks);
    if (blk_begin == 0) {
        sha256_init(h);
    } else {
        const uint4 lo = state[2 * b], hi = state[2 * b + 1];
        h[0] = lo.x; h[1] = lo.y; h[2] = lo.z; h[3] = lo.w; h[4] = hi.x; h[5] = hi.y; h[6] = hi.z; h[7] = hi.w;
    }
    uint32_t *mine = kw[warp];
#pragma unroll 1
    for (uint32_t base = blk_begin; base < blk_end; base += G) {
        const uint32_t blk = base + sub;
        if (blk < blk_end) {
            uint32_t w[16];
            if (blk >= 1 && 64u * blk + 64u <= 32u + 32u * (uint32_t)n) {  // both halves inside the blob: 128-bit loads
                ld_be_words8(w, blob + 64 * (size_t)blk - 32);
                ld_be_words8(w + 8, blob + 64 * (size_t)blk);
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++) w[j] = challenge_msg_word(blob, cm, n, blk, j);
            }
#pragma unroll
            for (int t = 0; t < 64; t++) {
                uint32_t wt;
                if (t < 16) {
                    wt = w[t];
                } else {
                    uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
                    uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                    uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
                    wt = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
                    w[t & 15] = wt;
                }
                mine[t * 32 + lane] = wt + K[t];
            }
        }
        __syncwarp();
        const uint32_t nb = min((uint32_t)G, blk_end - base);
#pragma unroll 1
        for (uint32_t i = 0; i < nb; i++) {
            const uint32_t col = slot * G + i;
            uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
            for (int t = 0; t < 64; t++) {
                // A scheduler that carries one warp runs one dependent chain, so the depth of a round counts too:
                // h + K + W and d + h + K + W do not depend on this round's e, which leaves rotate -> xor -> ONE
                // three-input add on the path from e to the next e (and from a to the next a).
                // The rotations and the boolean functions (10 instructions per round) go to the 16-lane integer pipe, 2 cycles
                // each.  The two sums on the dependent path (e -> next e, a -> next a) are one three-input add each; the
                // others are written as multiply-adds by a run-time 1 so that they issue on the FMA pipe beside them.
                const uint32_t hkw = fadd(hh, mine[t * 32 + col]), dhkw = fadd(d, hkw);
                const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
                const uint32_t ch = (e & f) ^ (~e & g);
                const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
                const uint32_t mj = (a & bb) ^ (a & c) ^ (bb & c);
                const uint32_t t1 = fadd(hkw, fadd(S1, ch));
                hh = g; g = f; f = e; e = dhkw + S1 + ch; d = c; c = bb; bb = a; a = t1 + S0 + mj;
            }
            h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
        }
        __syncwarp();
    }
    if (sub == 0 && live) {
        if (blk_end < nblocks) {
            state[2 * b] = make_uint4(h[0], h[1], h[2], h[3]);
            state[2 * b + 1] = make_uint4(h[4], h[5], h[6], h[7]);
            return;
        }
        fr_t z;
#pragma unroll
        for (int i = 0; i < 8; i++) z.l[i] = h[7 - i];
        scalar_reduce(z);
        st_fr(z_out + b, z);
    }
}
template <int G>
__global__ void __launch_bounds__(128) k_challenge_group(const uint8_t *__restrict__ blobs, const uint8_t *__restrict__ commitments,
                                                         uint32_t count, int n, fr_t *__restrict__ z_out, uint32_t one,
                                                         uint32_t blk_begin, uint32_t blk_end, uint4 *__restrict__ state) {
    __shared__ uint32_t kw[4][64 * 32];  // [round t][lane]: conflict-free writes; reads are broadcasts inside a blob's lanes
    challenge_group_block<G>(kw, blockIdx.x, blobs, commitments, count, n, z_out, one, blk_begin, blk_end, state);
}
// caller-supplied evaluation points (compute_kzg_proof): 32 big-endian bytes each, must be
// canonical (reference src/kzg.rs:452 -> bytes_to_bls_field)
__global__ void k_load_scalars(const uint8_t *__restrict__ in, uint32_t count, fr_t *__restrict__ out, int32_t *status) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fr_t s;
    scalar_from_be32(s, in + 32ull * i);
    if (!fr_is_canonical(s)) { atomicMax(status + i, (int)KZG_BADARGS); fe_set_zero(s); }
    st_fr(out + i, s);
}

// ------------------------------------------------------------------ evaluation + quotient
// One CTA per blob.  With d_i = z - w_i:
//   y   = (z^n - 1)/n * sum_i p_i w_i / d_i                    (or p_m when z == w_m)
//   q_i = (p_i - y)/(w_i - z) = (y - p_i)/d_i                  (i != m)
//   q_m = 1/z * sum_{i != m} (p_i - y) w_i / d_i               (only when z == w_m)
// The n inverses come from one Montgomery-trick inversion per CTA: thread-local prefix
// products, a block-wide prefix and suffix scan of the per-thread totals, one field
// inversion, and the unwinding.  inv[] (n per blob, global, L2-resident) holds the prefix
// products and then the inverses.  QUOT = false stops after y (verification).
// Outputs: zy_out (optional) = z || y as 32-byte big-endian records; with QUOT the canonical q_i
// replace the inverses in inv[], ready for the recoding of the MSM (msm_digits_from_scalars).
#define KZG_EVAL_THREADS 256
template <bool QUOT>
__global__ void __launch_bounds__(KZG_EVAL_THREADS) k_eval_quotient(
    const uint8_t *__restrict__ blobs, const fr_t *__restrict__ z_canon, const fr_t *__restrict__ roots, int n,
    fr_t *__restrict__ inv, fr_t *__restrict__ poly, uint8_t *__restrict__ zy_out, int32_t *status) {
    constexpr int T = KZG_EVAL_THREADS;
    __shared__ fr_t sh_pre[2][T];
    __shared__ fr_t sh_suf[2][T];
    __shared__ fr_t sh_val[2];
    __shared__ int sh_m;
    const int tid = threadIdx.x;
    const uint32_t b = blockIdx.x;
    const uint8_t *blob = blobs + (size_t)b * n * 32;
    fr_t *binv = inv + (size_t)b * n, *bpoly = poly + (size_t)b * n;
    const fr_t one = fe_one<FrParams>();
    fr_t z;
    ld_fr(z, z_canon + b);
    if (zy_out && tid == 0) st_scalar_be32(zy_out + 64ull * b, z);
    fe_to_mont(z, z);
    if (tid == 0) sh_m = -1;
    __syncthreads();
    // pass 1: p_i, d_i, thread-local prefix products
    fr_t run = one;
#pragma unroll 1
    for (int i = tid; i < n; i += T) {
        fr_t s, p, d, w;
        scalar_from_be32(s, blob + 32ull * i);
        if (!fr_is_canonical(s)) { atomicMax(status + b, (int)KZG_BADARGS); fe_set_zero(s); }
        // p_i stays CANONICAL: read as a Montgomery residue it is p_i / R, the barycentric sum and the
        // quotient are linear in p, and every other factor (w_i, 1/d_i, (z^n - 1)/n, 1/z) is a true
        // Montgomery residue -- so y and q_i come out as Montgomery(y / R) = the canonical y, q_i.
        // Two conversions per element (8 -> 6 products) are never done.
        p = s;
        st_fr(bpoly + i, p);
        ld_fr(w, roots + i);
        fe_sub(d, z, w);
        if (fe_is_zero(d)) { sh_m = i; d = one; }
        st_fr(binv + i, run);
        fe_mul(run, run, d);
    }
    sh_pre[0][tid] = run;
    sh_suf[0][tid] = run;
    __syncthreads();
    // inclusive prefix and suffix products over the T per-thread totals
    int cur = 0;
#pragma unroll 1
    for (int off = 1; off < T; off <<= 1) {
        fr_t a = sh_pre[cur][tid], s = sh_suf[cur][tid];
        if (tid >= off) fe_mul(a, sh_pre[cur][tid - off], a);
        if (tid + off < T) fe_mul(s, s, sh_suf[cur][tid + off]);
        sh_pre[cur ^ 1][tid] = a;
        sh_suf[cur ^ 1][tid] = s;
        cur ^= 1;
        __syncthreads();
    }
    if (tid == 0) {
        fr_t t;
        fr_inv(t, sh_pre[cur][T - 1]);
        sh_val[0] = t;
    }
    __syncthreads();
    const int m = sh_m;
    fr_t inv_run = sh_val[0];
    if (tid > 0) fe_mul(inv_run, inv_run, sh_pre[cur][tid - 1]);
    if (tid + 1 < T) fe_mul(inv_run, inv_run, sh_suf[cur][tid + 1]);
    // pass 2 (backwards): inverses and the barycentric sum
    fr_t acc;
    fe_set_zero(acc);
    int last = tid + ((n - 1 - tid) / T) * T;  // largest i = tid (mod T) below n (negative range when tid >= n)
#pragma unroll 1
    for (int i = (tid < n ? last : -1); i >= 0; i -= T) {
        fr_t w, d, pre, iv;
        ld_fr(w, roots + i);
        fe_sub(d, z, w);
        if (i == m) d = one;
        ld_fr(pre, binv + i);
        fe_mul(iv, inv_run, pre);
        fe_mul(inv_run, inv_run, d);
        st_fr(binv + i, iv);
        if (m < 0) {
            fr_t p, t;
            ld_fr(p, bpoly + i);
            fe_mul(t, iv, w);
            fe_mul(t, t, p);
            fe_add(acc, acc, t);
        }
    }
    __syncthreads();  // sh_pre is reused for the reduction
    fr_t *red = sh_pre[0];
    red[tid] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = T / 2; s > 0; s >>= 1) {
        if (tid < s) { fr_t t; fe_add(t, red[tid], red[tid + s]); red[tid] = t; }
        __syncthreads();
    }
    if (tid == 0) {
        fr_t y;
        if (m >= 0) {
            ld_fr(y, bpoly + m);
        } else {
            fr_t zn = z;
            for (int k = 1; k < n; k <<= 1) fe_sqr(zn, zn);  // z^n, n a power of two
            fe_sub(zn, zn, one);
            fe_mul(y, red[0], zn);
            fe_mul(y, y, fr_inv_n_mont(n));
        }
        sh_val[1] = y;
        if (zy_out) {
            st_scalar_be32(zy_out + 64ull * b + 32, y);  // canonical already (see pass 1)
        }
    }
    __syncthreads();
    if (!QUOT) return;
    const fr_t y = sh_val[1];
    fe_set_zero(acc);
#pragma unroll 1
    for (int i = tid; i < n; i += T) {
        if (i == m) continue;
        fr_t p, iv, q;
        ld_fr(p, bpoly + i);
        ld_fr(iv, binv + i);
        fe_sub(q, y, p);
        fe_mul(q, q, iv);
        if (m >= 0) {
            fr_t w, t;
            ld_fr(w, roots + i);
            fe_mul(t, q, w);      // = -(p_i - y) w_i / d_i
            fe_sub(acc, acc, t);
        }
        st_fr(binv + i, q);  // q is the canonical q_i
    }
    if (m >= 0) {  // uniform per CTA
        __syncthreads();
        red[tid] = acc;
        __syncthreads();
#pragma unroll 1
        for (int s = T / 2; s > 0; s >>= 1) {
            if (tid < s) { fr_t t; fe_add(t, red[tid], red[tid + s]); red[tid] = t; }
            __syncthreads();
        }
        if (tid == 0) {
            fr_t zi, q;
            fr_inv(zi, z);  // z = w_m != 0
            fe_mul(q, red[0], zi);
            st_fr(binv + m, q);
        }
    }
}

// ------------------------------------------------------------------ batch verification, phase B
// Per blob i (global index first + i): r_i = r^(first+i) and the three products
//   V_i = [r_i] proof_i,   U1_i = [r_i] C_i,   U2_i = [r_i z_i] proof_i,   s_i = r_i y_i
// (reference verify_kzg_proof_batch src/kzg.rs:579-627 and compute_r_powers
// src/utils.rs:426-474; sum_i r_i [y_i]G is folded into one scalar).
// Three threads per blob, one 255-bit double-and-add each (the batch is small and this kernel
// is latency-bound, so the scalar multiplications run side by side).
// out[i] = V_i, out[count + 2i] = U1_i, out[count + 2i + 1] = U2_i (Jacobian), sy[i] = s_i.
__global__ void __launch_bounds__(96) k_verify_terms(const g1_affine_t *__restrict__ cpts, const g1_affine_t *__restrict__ ppts,
                                                     const uint8_t *__restrict__ zy, fr_t r_canon, uint64_t first,
                                                     uint32_t count, g1_jac_t *__restrict__ out_pts, fr_t *__restrict__ sy) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i = t / 3, role = t - 3 * i;
    if (i >= count) return;
    fr_t r, ri = fe_one<FrParams>();
    fe_to_mont(r, r_canon);
    {
        uint64_t e = first + i;
        fr_t base = r;
#pragma unroll 1
        while (e) {
            if (e & 1) fe_mul(ri, ri, base);
            fe_sqr(base, base);
            e >>= 1;
        }
    }
    fr_t k = ri;
    if (role == 0) {
        fr_t y, s;
        scalar_from_be32(y, zy + 64ull * i + 32);
        fe_to_mont(y, y);
        fe_mul(s, ri, y);
        st_fr(sy + i, s);
    } else if (role == 2) {
        fr_t z;
        scalar_from_be32(z, zy + 64ull * i);
        fe_to_mont(z, z);
        fe_mul(k, ri, z);
    }
    fe_from_mont(k, k);
    const g1_affine_t base_pt = role == 1 ? cpts[i] : ppts[i];
    g1_jac_t acc;
    g1j_mul_glv(acc, base_pt, k.l);
    out_pts[role == 0 ? i : count + 2 * i + (role - 1)] = acc;
}
// Sums of Jacobian points in two stages, so that a 16,384-blob batch (49,152 terms) is spread over the whole GPU:
//   k_jac_sum_slices: segment 0 = pts[0, count0), segment 1 = pts[count0, count0 + count1); block b sums slice
//                     b % per of segment b / per (strided partial sums per thread, then a shared-memory tree)
//                     -> partials[b], Jacobian
//   k_jac_sum:        block s sums pts[s ? count0 : 0, ...) of its segment -> out[s], affine
#define KZG_JSUM_THREADS 128
KZG_D void jac_block_sum(g1_jac_t &total, const g1_jac_t *__restrict__ base, uint32_t lo, uint32_t hi, g1_jac_t *red) {
    g1_jac_t acc;
    g1j_set_inf(acc);
#pragma unroll 1
    for (uint32_t i = lo + threadIdx.x; i < hi; i += KZG_JSUM_THREADS) {
        g1_jac_t t = base[i];
        g1j_add(acc, acc, t);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = KZG_JSUM_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            g1_jac_t a = red[threadIdx.x], b = red[threadIdx.x + s];
            g1j_add(a, a, b);
            red[threadIdx.x] = a;
        }
        __syncthreads();
    }
    total = red[0];
}
__global__ void __launch_bounds__(KZG_JSUM_THREADS) k_jac_sum_slices(const g1_jac_t *__restrict__ pts, uint32_t count0, uint32_t count1,
                                                                     uint32_t per, g1_jac_t *__restrict__ partials) {
    __shared__ g1_jac_t red[KZG_JSUM_THREADS];
    const uint32_t seg = blockIdx.x / per, j = blockIdx.x - seg * per;
    const uint32_t cnt = seg ? count1 : count0;
    const uint32_t lo = (uint32_t)((uint64_t)cnt * j / per), hi = (uint32_t)((uint64_t)cnt * (j + 1) / per);
    g1_jac_t total;
    jac_block_sum(total, pts + (seg ? count0 : 0), lo, hi, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}
__global__ void __launch_bounds__(KZG_JSUM_THREADS) k_jac_sum(const g1_jac_t *__restrict__ pts, uint32_t count0, uint32_t count1,
                                                              g1_affine_t *__restrict__ out) {
    __shared__ g1_jac_t red[KZG_JSUM_THREADS];
    g1_jac_t total;
    jac_block_sum(total, pts + (blockIdx.x ? count0 : 0), 0, blockIdx.x ? count1 : count0, red);
    if (threadIdx.x == 0) {
        g1_affine_t a;
        g1j_to_affine(a, total);
        out[blockIdx.x] = a;
    }
}
// sum of `count` Montgomery scalars -> out[0] (one CTA)
__global__ void __launch_bounds__(256) k_fr_sum(const fr_t *__restrict__ in, uint32_t count, fr_t *__restrict__ out) {
    __shared__ fr_t red[256];
    fr_t acc;
    fe_set_zero(acc);
    for (uint32_t i = threadIdx.x; i < count; i += 256) { fr_t t; ld_fr(t, in + i); fe_add(acc, acc, t); }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { fr_t t; fe_add(t, red[threadIdx.x], red[threadIdx.x + s]); red[threadIdx.x] = t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}
// partial record of one shard: A (96 B) || B (96 B) || s (32 B); points as uncompressed
// big-endian affine coordinates, bit 6 of byte 0 set for infinity
__global__ void k_write_partial(const g1_affine_t *__restrict__ ab, const fr_t *__restrict__ s, uint8_t *__restrict__ out) {
    if (threadIdx.x < 2) {
        g1_affine_t p = ab[threadIdx.x];
        uint8_t *o = out + 96 * threadIdx.x;
        if (g1a_is_inf(p)) {
            for (int i = 0; i < 96; i++) o[i] = 0;
            o[0] = 0x40;
        } else {
            fp_t c;
            fe_from_mont(c, p.x); fp_to_be48(o, c);
            fe_from_mont(c, p.y); fp_to_be48(o + 48, c);
        }
    } else if (threadIdx.x == 2) {
        fr_t c;
        fe_from_mont(c, s[0]);
        scalar_to_be32(out + 192, c);
    }
}

#endif  // __CUDACC__

}  // namespace kzg

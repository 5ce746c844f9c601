// host_sha256.h -- SHA-256 on the host for the one sequential hash of batch verification (`compute_r_powers`,
// reference src/utils.rs:426-474; the reference calls blst_sha256).  Uses the x86 SHA extensions when the CPU has
// them (a 16,384-blob batch hashes 2.6 MB), the portable compression function of sha256.cuh otherwise.
#pragma once
#include <stddef.h>
#include <stdint.h>

struct HostSha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len;
    void init();
    void update(const uint8_t *p, size_t n);
    void finish(uint8_t out[32]);
};
// true when the SHA-NI path is in use (reported by tools, not needed for correctness)
bool host_sha256_accelerated();

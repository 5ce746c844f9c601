// host_pairing.h -- the host-side final check (stand-in for blst's pairing, which the
// reference calls at src/utils.rs:189-214) and the G2 part of trusted-setup loading.
#pragma once
#include <stddef.h>
#include <stdint.h>

// decode the 65 G2 points (reference src/kzg.rs:874-887) and reject a setup given in monomial
// form (is_trusted_setup_in_lagrange_form, src/kzg.rs:802-830).  g1 in file order.
// Returns 0 or 1 (= bad arguments).
int host_check_setup(const uint8_t *g1_bytes, const uint8_t *g2_bytes, size_t n2);

// A G2 point with its Miller-loop lines precomputed (done once per context for [tau]G2).
struct host_g2_prepared;
host_g2_prepared *host_g2_prepare(const uint8_t g2_compressed[96]);  // nullptr on a bad encoding
void host_g2_prepared_free(host_g2_prepared *h);

// e(a1, a2) == e(b1, b2) with G1 points given as 96-byte uncompressed affine records
// (x || y big-endian, bit 6 of byte 0 = infinity); b2 == nullptr means the G2 generator.
int host_pairing_check_uncompressed(const uint8_t a1[96], const host_g2_prepared *a2, const uint8_t b1[96],
                                    const host_g2_prepared *b2_or_null_for_generator, int *ok);

// e(a1, a2) == e(b1, b2) on compressed inputs; returns 0 / 1 (bad encoding), result in *ok
int host_pairings_verify(const uint8_t a1[48], const uint8_t a2[96], const uint8_t b1[48], const uint8_t b2[96], int *ok);

// Last step of batch verification (reference src/kzg.rs:618-625): partials = k records of
// A (96 B) || B (96 B) || s (32 B big-endian scalar); accepts iff
// e(sum A, [tau]G2) == e(sum B - [sum s]G1, G2).  Returns 0 / 1 (malformed record).
int host_verify_finish(const uint8_t *partials, size_t k, const host_g2_prepared *tau, int *ok);

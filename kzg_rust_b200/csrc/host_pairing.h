// host_pairing.h -- the host-side final check (stand-in for blst's pairing, which the
// reference calls at src/utils.rs:189-214) and the G2 part of trusted-setup loading.
#pragma once
#include <stddef.h>
#include <stdint.h>

// decode the 65 G2 points (reference src/kzg.rs:874-887) and reject a setup given in monomial
// form (is_trusted_setup_in_lagrange_form, src/kzg.rs:802-830).  g1 in file order.
int host_check_setup(const uint8_t *g1_bytes, const uint8_t *g2_bytes, size_t n2);
// e(a1, a2) == e(b1, b2) on compressed inputs; returns 0 / error code, result in *ok
int host_pairings_verify(const uint8_t a1[48], const uint8_t a2[96], const uint8_t b1[48], const uint8_t b2[96], int *ok);

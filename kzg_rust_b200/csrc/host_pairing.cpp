#include "host_pairing.h"
// TEMPORARY first-slice stub; replaced by the real host pairing in the next milestone.
int host_check_setup(const uint8_t *, const uint8_t *, size_t) { return 0; }
int host_pairings_verify(const uint8_t *, const uint8_t *, const uint8_t *, const uint8_t *, int *ok) { *ok = 0; return 2; }

/* cabi_smoke.c -- the C ABI of include/kzg_b200.h exercised from plain C (no ctypes, no Python in the process):
 * load the setup, commit one blob, prove it, verify the pair, compare the commitment with the expected 48 bytes.
 *
 *   gcc -O2 -I include -o cabi_smoke tests/cabi_smoke.c -L kzg_rust_b200 -lkzg_b200 -Wl,-rpath,$PWD/kzg_rust_b200
 *   ./cabi_smoke setup.bin blob.bin expected_commitment.bin      (setup.bin = 4096 x 48 B g1 || 65 x 96 B g2)
 *
 * tests/test_gpu_cabi.py builds and runs it on the GPU box with a vector of the reference. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kzg_b200.h"

static unsigned char *read_file(const char *path, size_t want) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    unsigned char *buf = (unsigned char *)malloc(want);
    if (fread(buf, 1, want, f) != want) { fprintf(stderr, "%s is shorter than %zu bytes\n", path, want); exit(2); }
    fclose(f);
    return buf;
}

int main(int argc, char **argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s setup.bin blob.bin expected48.bin\n", argv[0]); return 2; }
    const size_t n = 4096, blob_bytes = n * KZG_B200_BYTES_PER_FIELD_ELEMENT;
    unsigned char *setup = read_file(argv[1], n * KZG_B200_BYTES_PER_G1 + KZG_B200_NUM_G2_POINTS * KZG_B200_BYTES_PER_G2);
    unsigned char *blob = read_file(argv[2], blob_bytes);
    unsigned char *expected = read_file(argv[3], KZG_B200_BYTES_PER_COMMITMENT);
    kzg_b200_ctx *ctx = NULL;
    int rc = kzg_b200_ctx_create(setup, n, setup + n * KZG_B200_BYTES_PER_G1, KZG_B200_NUM_G2_POINTS, 0, 7, &ctx);
    if (rc != KZG_B200_OK) { fprintf(stderr, "ctx_create failed: %d\n", rc); return 1; }
    if (kzg_b200_field_elements_per_blob(ctx) != n || kzg_b200_comb_width(ctx) != 7) { fprintf(stderr, "bad context\n"); return 1; }
    unsigned char commitment[48], proof[48];
    int32_t status = -1;
    rc = kzg_b200_blob_to_kzg_commitment_batch(ctx, blob, 1, commitment, &status);
    if (rc != KZG_B200_OK || status != KZG_B200_OK) { fprintf(stderr, "commit failed: %d / %d\n", rc, (int)status); return 1; }
    if (memcmp(commitment, expected, 48) != 0) { fprintf(stderr, "commitment differs from the reference vector\n"); return 1; }
    rc = kzg_b200_compute_blob_kzg_proof_batch(ctx, blob, commitment, 1, proof, &status);
    if (rc != KZG_B200_OK || status != KZG_B200_OK) { fprintf(stderr, "proof failed: %d / %d\n", rc, (int)status); return 1; }
    int ok = 0;
    rc = kzg_b200_verify_blob_kzg_proof_batch(ctx, blob, commitment, proof, 1, &ok);
    if (rc != KZG_B200_OK || !ok) { fprintf(stderr, "verify failed: %d / %d\n", rc, ok); return 1; }
    proof[47] ^= 1;  /* not a valid encoding any more, or a different point: either way not accepted */
    ok = 1;
    rc = kzg_b200_verify_blob_kzg_proof_batch(ctx, blob, commitment, proof, 1, &ok);
    if (rc == KZG_B200_OK && ok) { fprintf(stderr, "tampered proof accepted\n"); return 1; }
    blob[0] = 0xff;  /* first element >= r: per-blob status, not a call failure */
    rc = kzg_b200_blob_to_kzg_commitment_batch(ctx, blob, 1, commitment, &status);
    if (rc != KZG_B200_OK || status != KZG_B200_BAD_ARGS) { fprintf(stderr, "non-canonical blob not rejected: %d / %d\n", rc, (int)status); return 1; }
    kzg_b200_ctx_destroy(ctx);
    printf("cabi_smoke ok\n");
    free(setup); free(blob); free(expected);
    return 0;
}

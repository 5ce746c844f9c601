// proof / verify entry points (first slice: not yet wired)
extern "C" int kzg_b200_compute_blob_kzg_proof_batch(kzg_b200_ctx *, const uint8_t *, const uint8_t *, size_t, uint8_t *, int32_t *) { return KZG_B200_INTERNAL_ERROR; }
extern "C" int kzg_b200_compute_kzg_proof_batch(kzg_b200_ctx *, const uint8_t *, const uint8_t *, size_t, uint8_t *, uint8_t *, int32_t *) { return KZG_B200_INTERNAL_ERROR; }
extern "C" int kzg_b200_verify_blob_kzg_proof_batch(kzg_b200_ctx *, const uint8_t *, const uint8_t *, const uint8_t *, size_t, int *) { return KZG_B200_INTERNAL_ERROR; }
extern "C" int kzg_b200_verify_phase_a(kzg_b200_ctx *, const uint8_t *, const uint8_t *, const uint8_t *, size_t, uint8_t *) { return KZG_B200_INTERNAL_ERROR; }
extern "C" int kzg_b200_compute_r(const kzg_b200_ctx *, const uint8_t *, const uint8_t *, const uint8_t *, size_t, uint8_t *) { return KZG_B200_INTERNAL_ERROR; }
extern "C" int kzg_b200_verify_phase_b(kzg_b200_ctx *, const uint8_t *, const uint8_t *, const uint8_t *, size_t, const uint8_t *, uint64_t, uint8_t *) { return KZG_B200_INTERNAL_ERROR; }
extern "C" int kzg_b200_verify_finish(const kzg_b200_ctx *, const uint8_t *, size_t, int *) { return KZG_B200_INTERNAL_ERROR; }
extern "C" int kzg_b200_compute_blob_kzg_proof_device(kzg_b200_ctx *, const uint8_t *, const uint8_t *, size_t, uint8_t *, int32_t *) { return KZG_B200_INTERNAL_ERROR; }
extern "C" int kzg_b200_pairings_verify(const uint8_t a1[48], const uint8_t a2[96], const uint8_t b1[48], const uint8_t b2[96], int *ok) { return host_pairings_verify(a1, a2, b1, b2, ok); }

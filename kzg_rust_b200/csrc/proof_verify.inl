// proof_verify.inl -- host sequencing of the proof and verification paths (included by
// kzg_b200.cu).  Kernels live in frpath.cuh / msm.cuh; the only host arithmetic is the
// SHA-256 of compute_r and the final pairing check (host_pairing.cpp).

static int decode_points(kzg_b200_ctx *ctx, const uint8_t *d_bytes, g1_affine_t *d_out, int32_t *d_status, size_t count,
                         int check_subgroup, size_t status_mod) {
    stage_begin(ctx, KZG_B200_STAGE_VALIDATE);
    int rc = g1_launch_decode(ctx->cur->stream, d_bytes, d_out, d_status, count, check_subgroup, status_mod);
    stage_end(ctx, 1);
    ctx->launches++;
    return rc;
}
// The same on the context's side stream, forked from and joined to the current lane's stream by the caller:
// small verification calls are chains of latency-bound kernels, and decompression + subgroup checks
// do not depend on the challenge hash that runs meanwhile.
static int decode_points_side(kzg_b200_ctx *ctx, const uint8_t *d_bytes, g1_affine_t *d_out, int32_t *d_status, size_t count,
                              int check_subgroup, size_t status_mod) {
    CU(cudaEventRecord(ctx->ev_side_fork, ctx->cur->stream));
    CU(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side_fork, 0));
    RC(g1_launch_decode(ctx->side_stream, d_bytes, d_out, d_status, count, check_subgroup, status_mod));
    CU(cudaEventRecord(ctx->ev_side_join, ctx->side_stream));
    ctx->launches++;
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ compute_blob_kzg_proof / compute_kzg_proof
// One chunk on the device.  Exactly one of d_commitments (blob proof: z from the Fiat-Shamir
// hash, reference src/kzg.rs:533-544) and d_zbytes (proof at a caller-supplied point,
// src/kzg.rs:446-457) is non-null.  d_zy (optional) receives z || y.
// d_z_ready (optional): challenges of this chunk already computed (device-resident calls hash all blobs of
// the call in one launch: a 4096-blob chunk alone is 28 SHA-256 streams per SM, which is latency-bound).
static int proof_chunk(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments, const uint8_t *d_zbytes,
                       size_t count, uint8_t *d_proofs, uint8_t *d_zy, int32_t *d_status, const fr_t *d_z_ready = nullptr,
                       size_t off = 0, const DeferredCompress *dc = nullptr) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    cudaStream_t st = ln->stream;
    bool side = false;
    if (d_z_ready) {
        // the caller validated the commitments (status already holds the verdicts) and hashed the challenges
        CU(cudaMemcpyAsync(ln->d_z, d_z_ready, count * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
    } else if (d_commitments) {
        CU(cudaMemsetAsync(d_status, 0, count * sizeof(int32_t), st));
        // small calls: the commitment check runs on the side stream beside the challenge hash and the evaluation
        side = count <= 256 && !ctx->profile;
        if (side) RC(decode_points_side(ctx, d_commitments, ln->d_pts, d_status, count, 1, count));
        else RC(decode_points(ctx, d_commitments, ln->d_pts, d_status, count, 1, count));
        stage_begin(ctx, KZG_B200_STAGE_CHALLENGE);
        int rc = fr_launch_challenge(st, d_blobs, d_commitments, count, ctx->n, ln->d_z);
        stage_end(ctx, 1);
        RC(rc);
    } else {
        CU(cudaMemsetAsync(d_status, 0, count * sizeof(int32_t), st));
        RC(fr_launch_load_scalars(st, d_zbytes, count, ln->d_z, d_status));
    }
    ctx->launches++;
    stage_begin(ctx, KZG_B200_STAGE_EVAL);
    int rc = fr_launch_eval(st, 1, d_blobs, ln->d_z, ctx->d_roots, ctx->n, ln->d_inv, ln->d_poly, d_zy, d_status, count);
    stage_end(ctx, 1);
    ctx->launches++;
    RC(rc);
    RC(msm_digits_from_scalars(ctx, ln->d_inv, count));  // the quotient, canonical, where the inverses were
    const g1_affine_t *res = nullptr;
    RC(msm_run(ctx, count, &res));
    if (side) CU(cudaStreamWaitEvent(st, ctx->ev_side_join, 0));  // the compression reads the status the check wrote
    return compress_or_park(ctx, res, off, count, d_status, d_proofs, dc);
}

extern "C" int kzg_b200_compute_blob_kzg_proof_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments,
                                                      size_t n, uint8_t *d_proofs_out, int32_t *d_status) {
    if (!ctx || (n && (!d_blobs || !d_commitments || !d_proofs_out || !d_status))) return KZG_B200_BAD_ARGS;
    if (!aligned16(d_blobs) || !aligned16(d_commitments) || !aligned16(d_proofs_out)) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t bpb = (size_t)ctx->n * 32;
    // all commitment validations and Fiat-Shamir challenges of the call in one launch each on the
    // caller-visible stream, before the lanes fork (per 4096-blob chunk both are latency-bound)
    const fr_t *d_z_all = nullptr;
    if (n > ctx->chunk) {
        if (n > ctx->z_all_elems) {
            if (ctx->d_z_all) CU(cudaFree(ctx->d_z_all));
            ctx->d_z_all = nullptr;
            ctx->z_all_elems = 0;
            CU(cudaMalloc(&ctx->d_z_all, n * (sizeof(fr_t) + sizeof(g1_affine_t))));
            ctx->z_all_elems = n;
        }
        ctx->cur = &ctx->lanes[0];
        CU(cudaMemsetAsync(d_status, 0, n * sizeof(int32_t), ctx->stream));
        RC(decode_points(ctx, d_commitments, reinterpret_cast<g1_affine_t *>(ctx->d_z_all + n), d_status, n, 1, n));
        stage_begin(ctx, KZG_B200_STAGE_CHALLENGE);
        int rc = fr_launch_challenge(ctx->stream, d_blobs, d_commitments, n, ctx->n, ctx->d_z_all);
        stage_end(ctx, 1);
        ctx->launches++;
        RC(rc);
        d_z_all = ctx->d_z_all;
    }
    DeferredCompress dc;
    RC(deferred_begin(ctx, n, &dc));
    RC(lanes_begin(ctx));
    size_t i = 0;
    for (size_t off = 0; off < n; off += ctx->chunk, i++) {
        size_t cnt = std::min(ctx->chunk, n - off);
        lane_select(ctx, i);
        RC(proof_chunk(ctx, d_blobs + off * bpb, d_commitments + off * 48, nullptr, cnt, d_proofs_out + off * 48, nullptr,
                       d_status + off, d_z_all ? d_z_all + off : nullptr, off, &dc));
    }
    RC(lanes_end(ctx));
    return deferred_finish(ctx, &dc, d_status, d_proofs_out);
}

extern "C" int kzg_b200_compute_blob_kzg_proof_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments,
                                                     size_t n, uint8_t *proofs_out, int32_t *status) {
    if (!ctx || (n && (!blobs || !commitments || !proofs_out || !status))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t bpb = (size_t)ctx->n * 32, ch = ctx->chunk;
    return staged_chunks(
        ctx, n,
        [&](int slot, size_t off, size_t cnt) -> int {
            CU(cudaMemcpyAsync(ctx->d_stage_in + slot * ch * bpb, blobs + off * bpb, cnt * bpb, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaMemcpyAsync(ctx->d_stage_aux + slot * ch * 96, commitments + off * 48, cnt * 48, cudaMemcpyHostToDevice, ctx->copy_stream));
            return KZG_B200_OK;
        },
        [&](int slot, size_t off, size_t cnt) -> int {
            uint8_t *d_out = ctx->d_stage_out + slot * ch * 96;
            int32_t *d_st = ctx->d_status + slot * ch;
            RC(proof_chunk(ctx, ctx->d_stage_in + slot * ch * bpb, ctx->d_stage_aux + slot * ch * 96, nullptr, cnt, d_out, nullptr, d_st));
            CU(cudaMemcpyAsync(proofs_out + off * 48, d_out, cnt * 48, cudaMemcpyDeviceToHost, ctx->cur->stream));
            CU(cudaMemcpyAsync(status + off, d_st, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->cur->stream));
            return KZG_B200_OK;
        });
}

extern "C" int kzg_b200_compute_kzg_proof_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *z, size_t n,
                                                uint8_t *proofs_out, uint8_t *y_out, int32_t *status) {
    if (!ctx || (n && (!blobs || !z || !proofs_out || !y_out || !status))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t bpb = (size_t)ctx->n * 32, ch = ctx->chunk;
    std::vector<uint8_t> zy(n * 64);
    RC(staged_chunks(
        ctx, n,
        [&](int slot, size_t off, size_t cnt) -> int {
            CU(cudaMemcpyAsync(ctx->d_stage_in + slot * ch * bpb, blobs + off * bpb, cnt * bpb, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaMemcpyAsync(ctx->d_stage_aux + slot * ch * 96, z + off * 32, cnt * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
            return KZG_B200_OK;
        },
        [&](int slot, size_t off, size_t cnt) -> int {
            uint8_t *d_out = ctx->d_stage_out + slot * ch * 96;
            int32_t *d_st = ctx->d_status + slot * ch;
            RC(proof_chunk(ctx, ctx->d_stage_in + slot * ch * bpb, nullptr, ctx->d_stage_aux + slot * ch * 96, cnt, d_out, ctx->cur->d_zy, d_st));
            CU(cudaMemcpyAsync(proofs_out + off * 48, d_out, cnt * 48, cudaMemcpyDeviceToHost, ctx->cur->stream));
            CU(cudaMemcpyAsync(zy.data() + off * 64, ctx->cur->d_zy, cnt * 64, cudaMemcpyDeviceToHost, ctx->cur->stream));
            CU(cudaMemcpyAsync(status + off, d_st, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->cur->stream));
            return KZG_B200_OK;
        }));
    for (size_t i = 0; i < n; i++) {
        if (status[i] == KZG_B200_OK) memcpy(y_out + i * 32, zy.data() + 64 * i + 32, 32);
        else memset(y_out + i * 32, 0, 32);
    }
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ verify_blob_kzg_proof_batch
// Phase A (reference src/kzg.rs:671-683, per blob): validate C_i and proof_i, z_i, y_i.
static int verify_phase_a_locked(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments, const uint8_t *proofs,
                                 size_t n, uint8_t *zy_out) {
    const size_t bpb = (size_t)ctx->n * 32, ch = ctx->chunk;
    const size_t nchunks_total = (n + ch - 1) / ch;
    std::vector<int32_t> st(n);
    RC(staged_chunks(
        ctx, n,
        [&](int slot, size_t off, size_t cnt) -> int {
            uint8_t *aux = ctx->d_stage_aux + slot * ch * 96;
            CU(cudaMemcpyAsync(ctx->d_stage_in + slot * ch * bpb, blobs + off * bpb, cnt * bpb, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaMemcpyAsync(aux, commitments + off * 48, cnt * 48, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaMemcpyAsync(aux + cnt * 48, proofs + off * 48, cnt * 48, cudaMemcpyHostToDevice, ctx->copy_stream));
            return KZG_B200_OK;
        },
        [&](int slot, size_t off, size_t cnt) -> int {
            const uint8_t *d_blobs = ctx->d_stage_in + slot * ch * bpb, *aux = ctx->d_stage_aux + slot * ch * 96;
            int32_t *d_st = ctx->d_status + slot * ch;
            kzg_b200_ctx::Lane *ln = ctx->cur;
            cudaStream_t sm = ln->stream;
            CU(cudaMemsetAsync(d_st, 0, cnt * sizeof(int32_t), sm));
            // single-chunk calls validate the points beside the hash (profiling keeps the stages apart)
            const bool side = nchunks_total == 1 && !ctx->profile;
            if (side) RC(decode_points_side(ctx, aux, ln->d_pts, d_st, 2 * cnt, 1, cnt));
            else RC(decode_points(ctx, aux, ln->d_pts, d_st, 2 * cnt, 1, cnt));
            stage_begin(ctx, KZG_B200_STAGE_CHALLENGE);
            int rc = fr_launch_challenge(sm, d_blobs, aux, cnt, ctx->n, ln->d_z);
            stage_end(ctx, 1);
            RC(rc);
            stage_begin(ctx, KZG_B200_STAGE_EVAL);
            rc = fr_launch_eval(sm, 0, d_blobs, ln->d_z, ctx->d_roots, ctx->n, ln->d_inv, ln->d_poly, ln->d_zy, d_st, cnt);
            stage_end(ctx, 1);
            ctx->launches += 2;
            RC(rc);
            if (side) CU(cudaStreamWaitEvent(sm, ctx->ev_side_join, 0));
            CU(cudaMemcpyAsync(zy_out + off * 64, ln->d_zy, cnt * 64, cudaMemcpyDeviceToHost, sm));
            CU(cudaMemcpyAsync(st.data() + off, d_st, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, sm));
            return KZG_B200_OK;
        }));
    for (size_t i = 0; i < n; i++)
        if (st[i] != KZG_B200_OK) return KZG_B200_BAD_ARGS;
    return KZG_B200_OK;
}
extern "C" int kzg_b200_verify_phase_a(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments,
                                       const uint8_t *proofs, size_t n, uint8_t *zy_out) {
    if (!ctx || (n && (!blobs || !commitments || !proofs || !zy_out))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    return verify_phase_a_locked(ctx, blobs, commitments, proofs, n, zy_out);
}

// reference compute_r_powers (src/utils.rs:426-474), the hash only: the points are hashed in
// their compressed form, which for a validated point is the caller's own 48 bytes.
extern "C" int kzg_b200_compute_r(const kzg_b200_ctx *ctx, const uint8_t *commitments, const uint8_t *zy,
                                  const uint8_t *proofs, size_t n_total, uint8_t r_out[32]) {
    if (!ctx || !r_out || (n_total && (!commitments || !zy || !proofs))) return KZG_B200_BAD_ARGS;
    Sha256 h;
    h.init();
    h.update((const uint8_t *)"RCKZGBATCH___V1_", 16);
    uint8_t u[8];
    for (int i = 0; i < 8; i++) u[i] = (uint8_t)((uint64_t)ctx->n >> (56 - 8 * i));
    h.update(u, 8);
    for (int i = 0; i < 8; i++) u[i] = (uint8_t)((uint64_t)n_total >> (56 - 8 * i));
    h.update(u, 8);
    for (size_t i = 0; i < n_total; i++) {
        h.update(commitments + 48 * i, 48);
        h.update(zy + 64 * i, 64);
        h.update(proofs + 48 * i, 48);
    }
    uint8_t d[32];
    h.finish(d);
    fr_t r;
    scalar_from_be32(r, d);
    scalar_reduce(r);
    scalar_to_be32(r_out, r);
    return KZG_B200_OK;
}

// Phase B: the r-power linear combinations of one shard (reference src/kzg.rs:601-622).
// d_pts_ready (optional): the 2n points phase A decoded on this context (commitments, then proofs)
static int verify_phase_b_locked(kzg_b200_ctx *ctx, const uint8_t *commitments, const uint8_t *zy, const uint8_t *proofs,
                                 size_t n, const uint8_t r[32], uint64_t first_index, uint8_t partial_out[224],
                                 const g1_affine_t *d_pts_ready = nullptr, int check_subgroup = 0) {
    if (n == 0) {
        memset(partial_out, 0, 224);
        partial_out[0] = 0x40;
        partial_out[96] = 0x40;
        return KZG_B200_OK;
    }
    fr_t rc;
    scalar_from_be32(rc, r);
    if (!fr_is_canonical(rc)) return KZG_B200_BAD_ARGS;
    ctx->cur = &ctx->lanes[0];
    // one buffer: bytes (48 + 48 + 64) n | status 2n | points 2n | terms 3n (Jacobian) | sums 2 | sy n + 1 | partial
    const size_t o_c = 0, o_p = o_c + 48 * n, o_zy = o_p + 48 * n;
    size_t o_st = (o_zy + 64 * n + 15) / 16 * 16;
    size_t o_pts = (o_st + 2 * n * sizeof(int32_t) + 15) / 16 * 16;
    size_t o_terms = o_pts + 2 * n * sizeof(g1_affine_t);
    size_t o_sums = o_terms + 3 * n * sizeof(g1_jac_t);
    size_t o_sy = o_sums + 2 * sizeof(g1_affine_t);
    size_t o_part = o_sy + (n + 1) * sizeof(fr_t);
    size_t total = o_part + 256;
    if (total > ctx->vb_bytes) {  // grow-only buffer kept by the context
        if (ctx->d_vb) CU(cudaFree(ctx->d_vb));
        ctx->d_vb = nullptr;
        ctx->vb_bytes = 0;
        CU(cudaMalloc(&ctx->d_vb, total + total / 2));
        ctx->vb_bytes = total + total / 2;
    }
    uint8_t *d = ctx->d_vb;
    CU(cudaMemcpyAsync(d + o_c, commitments, 48 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d + o_p, proofs, 48 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d + o_zy, zy, 64 * n, cudaMemcpyHostToDevice, ctx->stream));
    int32_t *d_st = (int32_t *)(d + o_st);
    g1_affine_t *pts = (g1_affine_t *)(d + o_pts), *sums = (g1_affine_t *)(d + o_sums);
    g1_jac_t *terms = (g1_jac_t *)(d + o_terms);
    fr_t *sy = (fr_t *)(d + o_sy);
    CU(cudaMemsetAsync(d_st, 0, 2 * n * sizeof(int32_t), ctx->stream));
    // c and p byte arrays are contiguous: decode both with one launch (phase A did the subgroup checks)
    if (d_pts_ready) CU(cudaMemcpyAsync(pts, d_pts_ready, 2 * n * sizeof(g1_affine_t), cudaMemcpyDeviceToDevice, ctx->stream));
    else RC(decode_points(ctx, d + o_c, pts, d_st, 2 * n, check_subgroup, 2 * n));
    stage_begin(ctx, KZG_B200_STAGE_VERIFY_TERMS);
    int lrc = fr_launch_verify_terms(ctx->stream, pts, pts + n, d + o_zy, rc, first_index, n, terms, sy);
    if (lrc == KZG_B200_OK) lrc = fr_launch_verify_sums(ctx->stream, terms, sy, n, sums, sy + n, d + o_part);
    stage_end(ctx, 4);
    ctx->launches += 4;
    RC(lrc);
    std::vector<int32_t> st(2 * n);
    CU(cudaMemcpyAsync(st.data(), d_st, 2 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(partial_out, d + o_part, 224, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
    for (size_t i = 0; i < 2 * n; i++)
        if (st[i] != KZG_B200_OK) return KZG_B200_BAD_ARGS;
    return KZG_B200_OK;
}
extern "C" int kzg_b200_verify_phase_b(kzg_b200_ctx *ctx, const uint8_t *commitments, const uint8_t *zy, const uint8_t *proofs,
                                       size_t n, const uint8_t r[32], uint64_t first_index, uint8_t partial_out[224]) {
    if (!ctx || !r || !partial_out || (n && (!commitments || !zy || !proofs))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    return verify_phase_b_locked(ctx, commitments, zy, proofs, n, r, first_index, partial_out);
}

extern "C" int kzg_b200_verify_finish(const kzg_b200_ctx *ctx, const uint8_t *partials, size_t n_partials, int *ok) {
    if (!ctx || !ok || (n_partials && !partials)) return KZG_B200_BAD_ARGS;
    *ok = 0;
    return host_verify_finish(partials, n_partials, ctx->tau_prepared, ok) == 0 ? KZG_B200_OK : KZG_B200_BAD_ARGS;
}

extern "C" int kzg_b200_verify_blob_kzg_proof_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments,
                                                    const uint8_t *proofs, size_t n, int *ok) {
    if (!ctx || !ok || (n && (!blobs || !commitments || !proofs))) return KZG_B200_BAD_ARGS;
    *ok = 0;
    if (n == 0) { *ok = 1; return KZG_B200_OK; }  // reference src/kzg.rs:653-655
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    // n == 1: the reference takes the single-proof equation e(C - [y]G, G2) == e(proof, [tau - z]G2)
    // (src/kzg.rs:658-660 -> :409-426); for points of G1 it holds exactly when the batch equation
    // with r^0 = 1 does, so the same path serves both.
    std::vector<uint8_t> zy(n * 64);
    RC(verify_phase_a_locked(ctx, blobs, commitments, proofs, n, zy.data()));
    uint8_t r[32], partial[224];
    RC(kzg_b200_compute_r(ctx, commitments, zy.data(), proofs, n, r));
    // a single-chunk call still has phase A's decoded points in lane 0's workspace
    RC(verify_phase_b_locked(ctx, commitments, zy.data(), proofs, n, r, 0, partial, n <= ctx->chunk ? ctx->lanes[0].d_pts : nullptr));
    return host_verify_finish(partial, 1, ctx->tau_prepared, ok) == 0 ? KZG_B200_OK : KZG_B200_BAD_ARGS;
}

// reference verify_kzg_proof (src/kzg.rs:429-445 -> verify_kzg_proof_impl :409-426): no blob, the caller gives
// z and the claimed y.  e(C - [y]G, G2) == e(proof, [tau - z]G2) is, for points of G1, the batch equation with
// one term and r^0 = 1 -- e(proof, [tau]G2) == e(C - [y]G + [z]proof, G2) -- so phase B and the final check are
// reused; here the points get their subgroup check in phase B because there is no phase A.
extern "C" int kzg_b200_verify_kzg_proof(kzg_b200_ctx *ctx, const uint8_t commitment[48], const uint8_t z[32],
                                         const uint8_t y[32], const uint8_t proof[48], int *ok) {
    if (!ctx || !commitment || !z || !y || !proof || !ok) return KZG_B200_BAD_ARGS;
    *ok = 0;
    fr_t t;
    scalar_from_be32(t, z);
    if (!fr_is_canonical(t)) return KZG_B200_BAD_ARGS;  // bytes_to_bls_field, src/utils.rs:262-275
    scalar_from_be32(t, y);
    if (!fr_is_canonical(t)) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    uint8_t zy[64], one[32] = {0}, partial[224];
    memcpy(zy, z, 32);
    memcpy(zy + 32, y, 32);
    one[31] = 1;
    RC(verify_phase_b_locked(ctx, commitment, zy, proof, 1, one, 0, partial, nullptr, 1));
    return host_verify_finish(partial, 1, ctx->tau_prepared, ok) == 0 ? KZG_B200_OK : KZG_B200_BAD_ARGS;
}

extern "C" int kzg_b200_pairings_verify(const uint8_t a1[48], const uint8_t a2[96], const uint8_t b1[48], const uint8_t b2[96],
                                        int *ok) {
    if (!a1 || !a2 || !b1 || !b2 || !ok) return KZG_B200_BAD_ARGS;
    return host_pairings_verify(a1, a2, b1, b2, ok) == 0 ? KZG_B200_OK : KZG_B200_BAD_ARGS;
}

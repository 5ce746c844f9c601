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
    kzg_b200_ctx::Lane *ln = ctx->cur;
    CU(cudaEventRecord(ln->ev_side_fork, ln->stream));
    CU(cudaStreamWaitEvent(ln->side_stream, ln->ev_side_fork, 0));
    RC(g1_launch_decode(ln->side_stream, d_bytes, d_out, d_status, count, check_subgroup, status_mod));
    CU(cudaEventRecord(ln->ev_side_join, ln->side_stream));
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
        // small calls: the commitment check runs on the side stream beside the challenge hash and the evaluation.  When the
        // commitments were uploaded ahead of the blobs (host calls), it starts as soon as they are there.
        side = count <= 256 && !ctx->profile;
        if (side && ctx->aux_ready) {
            CU(cudaStreamWaitEvent(ln->side_stream, ctx->aux_ready, 0));
            CU(cudaMemsetAsync(d_status, 0, count * sizeof(int32_t), ln->side_stream));
            CU(cudaEventRecord(ln->ev_side_fork, ln->side_stream));
            CU(cudaStreamWaitEvent(st, ln->ev_side_fork, 0));  // the status is zeroed before the evaluation may mark it
            RC(g1_launch_decode(ln->side_stream, d_commitments, ln->d_pts, d_status, count, 1, count));
            CU(cudaEventRecord(ln->ev_side_join, ln->side_stream));
            ctx->launches++;
        } else {
            CU(cudaMemsetAsync(d_status, 0, count * sizeof(int32_t), st));
            if (side) RC(decode_points_side(ctx, d_commitments, ln->d_pts, d_status, count, 1, count));
            else RC(decode_points(ctx, d_commitments, ln->d_pts, d_status, count, 1, count));
        }
        stage_begin(ctx, KZG_B200_STAGE_CHALLENGE);
        int rc = fr_launch_challenge(st, d_blobs, d_commitments, count, ctx->n, ln->d_z, ctx->call_blobs, ctx->sms);
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
    RC(msm_digits_from_scalars(ctx, ln->d_inv, count, msm_form(ctx, count, dc) == 2));  // the quotient, canonical, where the inverses were
    // the compression reads the status the check wrote: it waits for the side stream
    return msm_and_compress(ctx, off, count, d_status, d_proofs, dc, side ? ln->ev_side_join : nullptr);
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
        int rc = fr_launch_challenge(ctx->stream, d_blobs, d_commitments, n, ctx->n, ctx->d_z_all, n, ctx->sms);
        stage_end(ctx, 1);
        ctx->launches++;
        RC(rc);
        d_z_all = ctx->d_z_all;
    }
    ctx->call_blobs = n;
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
    DeferredCompress dc;
    const PiecePlan plan = msm_piece_plan(ctx, n);
    RC(deferred_begin(ctx, n, &dc, n > std::min(plan.piece, ctx->chunk)));
    return staged_chunks(
        ctx, n, plan.piece,
        [&](int slot, size_t off, size_t cnt) -> int {
            CU(cudaMemcpyAsync(ctx->d_stage_aux + slot * ch * 96, commitments + off * 48, cnt * 48, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaEventRecord(ctx->ev_aux[slot], ctx->copy_stream));
            ctx->aux_recorded[slot] = true;
            CU(cudaMemcpyAsync(ctx->d_stage_in + slot * ch * bpb, blobs + off * bpb, cnt * bpb, cudaMemcpyHostToDevice, ctx->copy_stream));
            return KZG_B200_OK;
        },
        [&](int slot, size_t off, size_t cnt) -> int {
            uint8_t *d_out = dc.sums ? dc.out + off * 48 : ctx->d_stage_out + slot * ch * 96;
            int32_t *d_st = dc.sums ? dc.status + off : ctx->d_status + slot * ch;
            RC(proof_chunk(ctx, ctx->d_stage_in + slot * ch * bpb, ctx->d_stage_aux + slot * ch * 96, nullptr, cnt, d_out, nullptr, d_st,
                           nullptr, off, &dc));
            if (!dc.sums) {
                CU(cudaMemcpyAsync(proofs_out + off * 48, d_out, cnt * 48, cudaMemcpyDeviceToHost, ctx->cur->stream));
                CU(cudaMemcpyAsync(status + off, d_st, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->cur->stream));
            }
            return KZG_B200_OK;
        },
        [&]() -> int {
            if (!dc.sums) return KZG_B200_OK;
            RC(deferred_finish(ctx, &dc, dc.status, dc.out));
            CU(cudaMemcpyAsync(proofs_out, dc.out, n * 48, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(status, dc.status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            return KZG_B200_OK;
        },
        plan.first);
}

extern "C" int kzg_b200_compute_kzg_proof_batch(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *z, size_t n,
                                                uint8_t *proofs_out, uint8_t *y_out, int32_t *status) {
    if (!ctx || (n && (!blobs || !z || !proofs_out || !y_out || !status))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t bpb = (size_t)ctx->n * 32, ch = ctx->chunk;
    std::vector<uint8_t> zy(n * 64);
    const PiecePlan plan = msm_piece_plan(ctx, n);
    RC(staged_chunks(
        ctx, n, plan.piece,
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
        },
        plan.first));
    for (size_t i = 0; i < n; i++) {
        if (status[i] == KZG_B200_OK) memcpy(y_out + i * 32, zy.data() + 64 * i + 32, 32);
        else memset(y_out + i * 32, 0, 32);
    }
    return KZG_B200_OK;
}

// ------------------------------------------------------------------ verify_blob_kzg_proof_batch
// Device buffers of one verification call (grow-only, kept by the context): the decoded points and the (z, y)
// records of the WHOLE call, so that phase B reads what phase A left on the device, plus phase B's own workspace.
struct VerifyBufs {
    g1_affine_t *pts;      // 2n: commitments, then proofs
    uint8_t *zy;           // n x 64
    uint8_t *in_bytes;     // 96 n: commitment and proof bytes when they have to be uploaded for phase B
    int32_t *status;       // 2n
    g1_jac_t *terms;       // 3n
    g1_jac_t *partials;    // 2 x KZG_VERIFY_SUM_BLOCKS
    g1_affine_t *sums;     // 2
    fr_t *sy;              // n + 1
    uint8_t *partial;      // 256
    uint8_t *pip;          // workspace of the bucket method (fr_pip_workspace_bytes)
};
static int verify_bufs(kzg_b200_ctx *ctx, size_t n, VerifyBufs *vb, bool keep_phase_a = false) {
    if (!keep_phase_a) ctx->va_valid = false;  // the buffers are about to be rewritten
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_pts = 0, o_zy = o_pts + up(2 * n * sizeof(g1_affine_t)), o_in = o_zy + up(64 * n), o_st = o_in + up(96 * n),
                 o_terms = o_st + up(2 * n * sizeof(int32_t)), o_partials = o_terms + up(3 * n * sizeof(g1_jac_t)),
                 o_sums = o_partials + up(2 * KZG_VERIFY_SUM_BLOCKS * sizeof(g1_jac_t)), o_sy = o_sums + up(2 * sizeof(g1_affine_t)),
                 o_part = o_sy + up((n + 1) * sizeof(fr_t)), o_pip = o_part + 256, total = o_pip + fr_pip_workspace_bytes(n);
    if (total > ctx->vb_bytes) {
        if (ctx->d_vb) CU(cudaFree(ctx->d_vb));
        ctx->d_vb = nullptr;
        ctx->vb_bytes = 0;
        CU(cudaMalloc(&ctx->d_vb, total + total / 2));
        ctx->vb_bytes = total + total / 2;
    }
    uint8_t *d = ctx->d_vb;
    vb->pts = (g1_affine_t *)(d + o_pts); vb->zy = d + o_zy; vb->in_bytes = d + o_in; vb->status = (int32_t *)(d + o_st);
    vb->terms = (g1_jac_t *)(d + o_terms); vb->partials = (g1_jac_t *)(d + o_partials); vb->sums = (g1_affine_t *)(d + o_sums);
    vb->sy = (fr_t *)(d + o_sy); vb->partial = d + o_part; vb->pip = d + o_pip;
    return KZG_B200_OK;
}

// Phase A of one chunk on the current lane (reference src/kzg.rs:671-683, per blob): validate C_i and proof_i
// (beside the hash, on the lane's side stream), z_i, y_i.  d_cp: cnt commitments then cnt proofs, 48 B each.
// The decoded points go to pts[off + i] (commitments) and pts[n_total + off + i] (proofs), (z, y) to zy[off + i].
static int verify_chunk_a(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments, const uint8_t *d_proofs, size_t cnt,
                          size_t off, size_t n_total, const VerifyBufs &vb, int32_t *d_st) {
    kzg_b200_ctx::Lane *ln = ctx->cur;
    cudaStream_t sm = ln->stream;
    const bool side = !ctx->profile;  // profiling keeps the stages apart
    cudaStream_t sd = side ? ln->side_stream : sm;
    if (side && ctx->aux_ready) {
        // host calls: the commitments and proofs of the slot were uploaded ahead of its blobs.  The validation starts as
        // soon as they are there -- under the upload of the blobs -- and therefore never at the same instant as the hash:
        // two latency-bound launches released by the same event were placed on the same SMs every other call (device
        // timestamps of a six-blob call: hash 2.03 / validation 1.71 ms apart, 2.43 / 3.12 ms together).
        CU(cudaStreamWaitEvent(sd, ctx->aux_ready, 0));
        CU(cudaMemsetAsync(d_st, 0, cnt * sizeof(int32_t), sd));
        CU(cudaEventRecord(ln->ev_side_fork, sd));
        CU(cudaStreamWaitEvent(sm, ln->ev_side_fork, 0));  // the status is zeroed before the evaluation may mark it
    } else {
        CU(cudaMemsetAsync(d_st, 0, cnt * sizeof(int32_t), sm));
        if (side) {
            CU(cudaEventRecord(ln->ev_side_fork, sm));
            CU(cudaStreamWaitEvent(sd, ln->ev_side_fork, 0));
        }
    }
    stage_begin(ctx, KZG_B200_STAGE_VALIDATE);
    int rc = g1_launch_decode2(sd, d_commitments, d_proofs, vb.pts + off, vb.pts + n_total + off, d_st, cnt, 1);
    stage_end(ctx, 1);
    RC(rc);
    if (side) CU(cudaEventRecord(ln->ev_side_join, sd));
    stage_begin(ctx, KZG_B200_STAGE_CHALLENGE);
    if (ctx->cur_slices > 1) {
        // the blobs arrive in column slices: slice k holds message blocks [k B, (k+1) B) of every blob (the 32-byte header shifts
        // the blocks by half a block, so the block that straddles two slices goes with the later one), the last slice runs to the
        // end of the message (commitment + padding).  Each launch waits for its slice only.
        const int S = ctx->cur_slices, g = fr_challenge_form(cnt, ctx->call_blobs, ctx->sms);
        const uint32_t per = (uint32_t)((size_t)ctx->n * 32 / S / 64);
        for (int k = 0; k < S && rc == KZG_B200_OK; k++) {
            CU(cudaStreamWaitEvent(sm, ctx->cur_slice_ev[k], 0));
            rc = fr_launch_challenge_range(sm, d_blobs, d_commitments, cnt, ctx->n, ln->d_z, g, k * per, k == S - 1 ? 0xffffffffu : (k + 1) * per,
                                           ln->d_sha_state);
        }
        stage_end(ctx, S);
        ctx->launches += S - 1;
    } else {
        rc = fr_launch_challenge(sm, d_blobs, d_commitments, cnt, ctx->n, ln->d_z, ctx->call_blobs, ctx->sms);
        stage_end(ctx, 1);
    }
    RC(rc);
    stage_begin(ctx, KZG_B200_STAGE_EVAL);
    rc = fr_launch_eval(sm, 0, d_blobs, ln->d_z, ctx->d_roots, ctx->n, ln->d_inv, ln->d_poly, vb.zy + 64 * off, d_st, cnt);
    stage_end(ctx, 1);
    ctx->launches += 3;
    RC(rc);
    if (side) CU(cudaStreamWaitEvent(sm, ln->ev_side_join, 0));
    return KZG_B200_OK;
}

// Phase A over host buffers: chunks staged through the upload slots, two chunks in flight.
// keep != nullptr: the decoded points and (z, y) records stay in the call's device buffers for phase B.
// Verification chunks are smaller than the staging slots (KZG_VERIFY_PIECE blobs): a chunk's hash and evaluation are
// latency-bound -- 2 .. 2.6 ms whether it has 256 blobs or 4,096 -- so after the last upload only a SMALL chunk should be
// left to compute, and the earlier ones run under the uploads behind them (4,096 blobs: 19.2 -> 16.x ms).  The (z, y)
// records and the status words come back to pinned memory, so the host never waits inside the loop and both lanes stay fed.
#define KZG_VERIFY_PIECE 1024
static int verify_phase_a_locked(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments, const uint8_t *proofs,
                                 size_t n, uint8_t *zy_out, VerifyBufs *keep = nullptr) {
    const size_t bpb = (size_t)ctx->n * 32, ch = ctx->chunk;
    VerifyBufs vb;
    RC(verify_bufs(ctx, n, &vb));
    if (keep) *keep = vb;
    RC(ensure_pinned(ctx, n * (64 + sizeof(int32_t))));
    uint8_t *h_zy = ctx->h_pin;
    int32_t *h_st = reinterpret_cast<int32_t *>(ctx->h_pin + 64 * n);
    // KZG_B200_HASH_SLICES (1 = plain copies), KZG_B200_SLICE_MIN_BLOBS: smallest chunk that is uploaded in slices
    const int slices = std::min(KZG_HASH_SLICES, std::max(1, env_int("KZG_B200_HASH_SLICES", KZG_HASH_SLICES)));
    const int slice_min = std::max(1, env_int("KZG_B200_SLICE_MIN_BLOBS", 64));
    RC(staged_chunks(
        ctx, n, (size_t)std::max(1, env_int("KZG_B200_VERIFY_PIECE", KZG_VERIFY_PIECE)),
        [&](int slot, size_t off, size_t cnt) -> int {
            uint8_t *aux = ctx->d_stage_aux + slot * ch * 96;
            CU(cudaMemcpyAsync(aux, commitments + off * 48, cnt * 48, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaMemcpyAsync(aux + cnt * 48, proofs + off * 48, cnt * 48, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaEventRecord(ctx->ev_aux[slot], ctx->copy_stream));
            ctx->aux_recorded[slot] = true;
            uint8_t *d_blobs = ctx->d_stage_in + slot * ch * bpb;
            // Column slices (see internal.h): worth it when the hash has lanes to spare (a grouped form) and a slice still is a
            // long row per blob; the hash of a slice must be a whole number of 32-block groups.
            const int S = slices;
            const bool sliced = S > 1 && !ctx->profile && cnt >= (size_t)slice_min && bpb % ((size_t)S * 64 * 32) == 0 &&
                                fr_challenge_form(cnt, n, ctx->sms) >= 2;
            if (!sliced) {
                CU(cudaMemcpyAsync(d_blobs, blobs + off * bpb, cnt * bpb, cudaMemcpyHostToDevice, ctx->copy_stream));
                return KZG_B200_OK;
            }
            const size_t w = bpb / S;
            for (int k = 0; k < S; k++) {
                CU(cudaMemcpy2DAsync(d_blobs + k * w, bpb, blobs + off * bpb + k * w, bpb, w, cnt, cudaMemcpyHostToDevice, ctx->copy_stream));
                CU(cudaEventRecord(ctx->ev_slice[slot][k], ctx->copy_stream));
            }
            ctx->slices_recorded[slot] = S;
            return KZG_B200_OK;
        },
        [&](int slot, size_t off, size_t cnt) -> int {
            const uint8_t *d_blobs = ctx->d_stage_in + slot * ch * bpb, *aux = ctx->d_stage_aux + slot * ch * 96;
            int32_t *d_st = ctx->d_status + slot * ch;
            RC(verify_chunk_a(ctx, d_blobs, aux, aux + cnt * 48, cnt, off, n, vb, d_st));
            cudaStream_t sm = ctx->cur->stream;
            CU(cudaMemcpyAsync(h_zy + off * 64, vb.zy + 64 * off, cnt * 64, cudaMemcpyDeviceToHost, sm));
            CU(cudaMemcpyAsync(h_st + off, d_st, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, sm));
            return KZG_B200_OK;
        }));
    memcpy(zy_out, h_zy, 64 * n);
    for (size_t i = 0; i < n; i++)
        if (h_st[i] != KZG_B200_OK) return KZG_B200_BAD_ARGS;
    return KZG_B200_OK;
}
static void points_digest(uint8_t out[32], const uint8_t *commitments, const uint8_t *proofs, size_t n) {
    HostSha256 h;
    h.init();
    uint64_t n64 = n;
    h.update(reinterpret_cast<const uint8_t *>(&n64), 8);
    h.update(commitments, 48 * n);
    h.update(proofs, 48 * n);
    h.finish(out);
}
extern "C" int kzg_b200_verify_phase_a(kzg_b200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments,
                                       const uint8_t *proofs, size_t n, uint8_t *zy_out) {
    if (!ctx || (n && (!blobs || !commitments || !proofs || !zy_out))) return KZG_B200_BAD_ARGS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    RC(verify_phase_a_locked(ctx, blobs, commitments, proofs, n, zy_out));
    if (n) {  // every point passed validation and sits decoded in the context's buffers: remember which bytes they were
        points_digest(ctx->va_digest, commitments, proofs, n);
        ctx->va_n = n;
        ctx->va_valid = true;
    }
    return KZG_B200_OK;
}

// reference compute_r_powers (src/utils.rs:426-474), the hash only: the points are hashed in
// their compressed form, which for a validated point is the caller's own 48 bytes.
static void compute_r_begin(HostSha256 &h, uint64_t field_elements, uint64_t n_total) {
    h.init();
    h.update((const uint8_t *)"RCKZGBATCH___V1_", 16);
    uint8_t u[8];
    for (int i = 0; i < 8; i++) u[i] = (uint8_t)(field_elements >> (56 - 8 * i));
    h.update(u, 8);
    for (int i = 0; i < 8; i++) u[i] = (uint8_t)(n_total >> (56 - 8 * i));
    h.update(u, 8);
}
static void compute_r_update(HostSha256 &h, const uint8_t *commitments, const uint8_t *zy, const uint8_t *proofs, size_t count) {
    for (size_t i = 0; i < count; i++) {
        h.update(commitments + 48 * i, 48);
        h.update(zy + 64 * i, 64);
        h.update(proofs + 48 * i, 48);
    }
}
static void compute_r_finish(HostSha256 &h, uint8_t r_out[32]) {
    uint8_t d[32];
    h.finish(d);
    fr_t r;
    scalar_from_be32(r, d);
    scalar_reduce(r);
    scalar_to_be32(r_out, r);
}
extern "C" int kzg_b200_compute_r(const kzg_b200_ctx *ctx, const uint8_t *commitments, const uint8_t *zy,
                                  const uint8_t *proofs, size_t n_total, uint8_t r_out[32]) {
    if (!ctx || !r_out || (n_total && (!commitments || !zy || !proofs))) return KZG_B200_BAD_ARGS;
    HostSha256 h;
    compute_r_begin(h, (uint64_t)ctx->n, n_total);
    compute_r_update(h, commitments, zy, proofs, n_total);
    compute_r_finish(h, r_out);
    return KZG_B200_OK;
}

// Phase B on the device: the r-power linear combinations of one shard (reference src/kzg.rs:601-622) from
// decoded, validated points vb.pts (n commitments at [0, n), n proofs at [n_stride, n_stride + n)) and vb.zy.
static int verify_phase_b_device(kzg_b200_ctx *ctx, const VerifyBufs &vb, size_t n, size_t n_stride, const uint8_t r[32],
                                 uint64_t first_index, uint8_t partial_out[224]) {
    fr_t rc;
    scalar_from_be32(rc, r);
    if (!fr_is_canonical(rc)) return KZG_B200_BAD_ARGS;
    ctx->cur = &ctx->lanes[0];
    stage_begin(ctx, KZG_B200_STAGE_VERIFY_TERMS);
    // KZG_B200_VERIFY_LADDER=1: one GLV ladder per term (k_verify_terms) instead of the bucket method -- the
    // independent path the tests compare the partial records with.  KZG_B200_PIP_C: force the window width.
    int lrc, nl;
    if (env_int("KZG_B200_VERIFY_LADDER", 0)) {
        lrc = fr_launch_verify_terms(ctx->stream, vb.pts, vb.pts + n_stride, vb.zy, rc, first_index, n, vb.terms, vb.sy);
        if (lrc == KZG_B200_OK) lrc = fr_launch_verify_sums(ctx->stream, vb.terms, vb.sy, n, vb.sums, vb.sy + n, vb.partial, vb.partials);
        nl = 5;
    } else {
        lrc = fr_launch_verify_pippenger(ctx->stream, vb.pts, vb.pts + n_stride, vb.zy, rc, first_index, n, env_int("KZG_B200_PIP_C", 0),
                                         vb.pip, vb.sy, vb.sums, vb.sy + n, vb.partial);
        nl = KZG_PIP_LAUNCHES;
    }
    stage_end(ctx, nl);
    ctx->launches += nl;
    RC(lrc);
    CU(cudaMemcpyAsync(partial_out, vb.partial, 224, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
    return KZG_B200_OK;
}
static void empty_partial(uint8_t partial_out[224]) {
    memset(partial_out, 0, 224);
    partial_out[0] = 0x40;
    partial_out[96] = 0x40;
}
// Phase B from host bytes: nothing is assumed about them -- the points are decompressed AND subgroup-checked here,
// z_i and y_i must be canonical (a caller may run phase B on another context, or on re-fetched data, than phase A).
// The one shortcut: when the commitment and proof bytes are exactly the ones this context's last phase A validated
// (same SHA-256), their decoded points are still in the context's buffers and are used as they are.
static int verify_phase_b_locked(kzg_b200_ctx *ctx, const uint8_t *commitments, const uint8_t *zy, const uint8_t *proofs,
                                 size_t n, const uint8_t r[32], uint64_t first_index, uint8_t partial_out[224]) {
    if (n == 0) { empty_partial(partial_out); return KZG_B200_OK; }
    for (size_t i = 0; i < 2 * n; i++) {
        fr_t t;
        scalar_from_be32(t, zy + 32 * i);
        if (!fr_is_canonical(t)) return KZG_B200_BAD_ARGS;  // bytes_to_bls_field, src/utils.rs:262-275
    }
    bool reuse = false;
    if (ctx->va_valid && ctx->va_n == n) {
        uint8_t dg[32];
        points_digest(dg, commitments, proofs, n);
        reuse = memcmp(dg, ctx->va_digest, 32) == 0;
    }
    VerifyBufs vb;
    RC(verify_bufs(ctx, n, &vb, reuse));
    ctx->cur = &ctx->lanes[0];
    CU(cudaMemcpyAsync(vb.zy, zy, 64 * n, cudaMemcpyHostToDevice, ctx->stream));
    if (reuse) return verify_phase_b_device(ctx, vb, n, n, r, first_index, partial_out);
    CU(cudaMemcpyAsync(vb.in_bytes, commitments, 48 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(vb.in_bytes + 48 * n, proofs, 48 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(vb.status, 0, n * sizeof(int32_t), ctx->stream));
    // Small shards: only the decompression runs in front of the sums; the subgroup checks (a 128-step ladder per point, as long
    // as the bucket method itself) run beside them on the side stream -- the sums of a point outside G1 are garbage, and the
    // call is rejected before anyone sees them.  verify_kzg_proof: 4.2 -> 3.1 ms.
    kzg_b200_ctx::Lane *ln = &ctx->lanes[0];
    const bool beside = !ctx->profile && (2 * n + 39) / 40 <= (size_t)ctx->sms / 2;  // ten points per warp, four warps per block
    stage_begin(ctx, KZG_B200_STAGE_VALIDATE);
    int rc = g1_launch_decode2(ctx->stream, vb.in_bytes, vb.in_bytes + 48 * n, vb.pts, vb.pts + n, vb.status, n, beside ? 0 : 1);
    stage_end(ctx, 1);
    ctx->launches++;
    RC(rc);
    if (beside) {
        CU(cudaEventRecord(ln->ev_side_fork, ctx->stream));
        CU(cudaStreamWaitEvent(ln->side_stream, ln->ev_side_fork, 0));
        RC(g1_launch_subgroup2(ln->side_stream, vb.pts, vb.pts + n, vb.status, n));
        CU(cudaEventRecord(ln->ev_side_join, ln->side_stream));
        ctx->launches++;
    }
    RC(verify_phase_b_device(ctx, vb, n, n, r, first_index, partial_out));
    if (beside) CU(cudaStreamWaitEvent(ctx->stream, ln->ev_side_join, 0));
    std::vector<int32_t> st(n);
    CU(cudaMemcpyAsync(st.data(), vb.status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < n; i++)
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
    VerifyBufs vb;
    // KZG_B200_TRACE=1: wall-clock milliseconds of the host-visible steps on stderr (tools/verify_stages.py)
    const bool trace = env_int("KZG_B200_TRACE", 0) != 0;  // 2: also the enqueue / wait phases of the staged chunks
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[kzg_b200 trace] %-12s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    RC(verify_phase_a_locked(ctx, blobs, commitments, proofs, n, zy.data(), &vb));
    lap("phase A");
    uint8_t r[32], partial[224];
    RC(kzg_b200_compute_r(ctx, commitments, zy.data(), proofs, n, r));
    lap("compute_r");
    // phase A left the decoded points and the (z, y) records of the whole call on the device
    RC(verify_phase_b_device(ctx, vb, n, n, r, 0, partial));
    lap("phase B");
    int rc = host_verify_finish(partial, 1, ctx->tau_prepared, ok) == 0 ? KZG_B200_OK : KZG_B200_BAD_ARGS;
    lap("finish");
    return rc;
}

// Phase A over blobs, commitments and proofs that are already in this context's GPU memory (16-byte aligned).
// Every validation and every Fiat-Shamir challenge of the call runs in ONE launch each (per chunk they are latency-bound:
// one SHA-256 stream per blob), the evaluations chunk by chunk.  On return (synchronous) the decoded points and the (z, y)
// records are in vb, and the pinned host buffer holds zy (n x 64), the commitment and proof bytes (n x 48 each) -- what
// compute_r_powers hashes -- and every point and field element has been checked.  hs != nullptr: that hash is computed
// here, chunk by chunk while the GPU evaluates the next chunk.
struct HostRecords { const uint8_t *zy, *commitments, *proofs; };
static int verify_phase_a_device_locked(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments, const uint8_t *d_proofs,
                                        size_t n, VerifyBufs *vb_out, HostSha256 *hs, HostRecords *rec) {
    const size_t bpb = (size_t)ctx->n * 32;
    VerifyBufs vb;
    RC(verify_bufs(ctx, n, &vb));
    *vb_out = vb;
    if (n > ctx->z_all_elems) {
        if (ctx->d_z_all) CU(cudaFree(ctx->d_z_all));
        ctx->d_z_all = nullptr;
        ctx->z_all_elems = 0;
        CU(cudaMalloc(&ctx->d_z_all, n * (sizeof(fr_t) + sizeof(g1_affine_t))));
        ctx->z_all_elems = n;
    }
    kzg_b200_ctx::Lane *ln = ctx->cur = &ctx->lanes[0];
    cudaStream_t sm = ctx->stream, sd = ctx->profile ? sm : ln->side_stream;
    CU(cudaMemsetAsync(vb.status, 0, n * sizeof(int32_t), sm));
    // Where the validation of the 2n points runs: beside the hash while both fit one warp per scheduler (up to ~4,700 blobs);
    // for larger calls beside the evaluation, after the hash -- two latency-bound kernels that share schedulers cost far more
    // than they overlap (profiles/validate_order_r2ab.txt: 8,192 blobs 21.1 ms beside the hash, 14.0 ms beside the evaluation).
    const int order = fr_challenge_form(n, n, ctx->sms) >= 2 ? 0 : 2;
    int rc;
    auto validate = [&](cudaStream_t s_) -> int {
        stage_begin(ctx, KZG_B200_STAGE_VALIDATE);
        int vrc = g1_launch_decode2(s_, d_commitments, d_proofs, vb.pts, vb.pts + n, vb.status, n, 1);
        stage_end(ctx, 1);
        return vrc;
    };
    if (order != 2) {
        CU(cudaEventRecord(ln->ev_side_fork, sm));
        CU(cudaStreamWaitEvent(sd, ln->ev_side_fork, 0));
        RC(validate(sd));
        CU(cudaEventRecord(ln->ev_side_join, sd));
    }
    stage_begin(ctx, KZG_B200_STAGE_CHALLENGE);
    rc = fr_launch_challenge(sm, d_blobs, d_commitments, n, ctx->n, ctx->d_z_all, n, ctx->sms);
    stage_end(ctx, 1);
    RC(rc);
    if (order == 2) {
        CU(cudaEventRecord(ln->ev_side_fork, sm));
        CU(cudaStreamWaitEvent(sd, ln->ev_side_fork, 0));
        RC(validate(sd));
        CU(cudaEventRecord(ln->ev_side_join, sd));
    }
    ctx->launches += 2;
    // The sequential hash of compute_r_powers (160 bytes per blob, host) runs chunk by chunk while the GPU evaluates the
    // next chunk: every chunk's (z, y) records come back to pinned memory behind their own event.
    const size_t nchunks = (n + ctx->chunk - 1) / ctx->chunk;
    RC(ensure_pinned(ctx, n * (64 + 48 + 48 + sizeof(int32_t))));
    while (ctx->ev_chunks.size() < nchunks + 1) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_chunks.push_back(e);
    }
    uint8_t *h_zy = ctx->h_pin, *h_cm = h_zy + 64 * n, *h_pr = h_cm + 48 * n;
    int32_t *h_st = reinterpret_cast<int32_t *>(h_pr + 48 * n);
    CU(cudaStreamWaitEvent(ctx->copy_stream, ln->ev_side_fork, 0));  // after whatever preceded this call on the stream
    CU(cudaMemcpyAsync(h_cm, d_commitments, n * 48, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(cudaMemcpyAsync(h_pr, d_proofs, n * 48, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_chunks[nchunks], ctx->copy_stream));
    stage_begin(ctx, KZG_B200_STAGE_EVAL);
    uint64_t evals = 0;
    for (size_t off = 0; off < n; off += ctx->chunk, evals++) {
        const size_t cnt = std::min(ctx->chunk, n - off);
        RC(fr_launch_eval(sm, 0, d_blobs + off * bpb, ctx->d_z_all + off, ctx->d_roots, ctx->n, ln->d_inv, ln->d_poly, vb.zy + 64 * off,
                          vb.status + off, cnt));
        if (!ctx->profile) {
            CU(cudaMemcpyAsync(h_zy + 64 * off, vb.zy + 64 * off, cnt * 64, cudaMemcpyDeviceToHost, sm));
            CU(cudaEventRecord(ctx->ev_chunks[evals], sm));
        }
    }
    stage_end(ctx, evals);
    ctx->launches += evals;
    if (ctx->profile) {  // the stage timers own the stream's event order: one copy at the end
        CU(cudaMemcpyAsync(h_zy, vb.zy, n * 64, cudaMemcpyDeviceToHost, sm));
        for (size_t c = 0; c < nchunks; c++) CU(cudaEventRecord(ctx->ev_chunks[c], sm));
    }
    CU(cudaStreamWaitEvent(sm, ln->ev_side_join, 0));
    CU(cudaMemcpyAsync(h_st, vb.status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, sm));
    CU(cudaEventSynchronize(ctx->ev_chunks[nchunks]));
    if (hs) {
        for (size_t c = 0; c < nchunks; c++) {
            const size_t off = c * ctx->chunk, cnt = std::min(ctx->chunk, n - off);
            CU(cudaEventSynchronize(ctx->ev_chunks[c]));
            compute_r_update(*hs, h_cm + 48 * off, h_zy + 64 * off, h_pr + 48 * off, cnt);
        }
    }
    CU(cudaStreamSynchronize(sm));
    stage_collect(ctx);
    for (size_t i = 0; i < n; i++)
        if (h_st[i] != KZG_B200_OK) return KZG_B200_BAD_ARGS;
    rec->zy = h_zy;
    rec->commitments = h_cm;
    rec->proofs = h_pr;
    return KZG_B200_OK;
}

// `verify_blob_kzg_proof_batch` over device-resident inputs: 160 bytes per blob come back to the host for the sequential
// hash of compute_r_powers, and the final pairing check runs on the host.  Synchronous.
extern "C" int kzg_b200_verify_blob_kzg_proof_batch_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments,
                                                           const uint8_t *d_proofs, size_t n, int *ok) {
    if (!ctx || !ok || (n && (!d_blobs || !d_commitments || !d_proofs))) return KZG_B200_BAD_ARGS;
    if (!aligned16(d_blobs) || !aligned16(d_commitments) || !aligned16(d_proofs)) return KZG_B200_BAD_ARGS;
    *ok = 0;
    if (n == 0) { *ok = 1; return KZG_B200_OK; }
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    VerifyBufs vb;
    HostRecords rec;
    HostSha256 hs;
    compute_r_begin(hs, (uint64_t)ctx->n, n);
    RC(verify_phase_a_device_locked(ctx, d_blobs, d_commitments, d_proofs, n, &vb, &hs, &rec));
    uint8_t r[32], partial[224];
    compute_r_finish(hs, r);
    RC(verify_phase_b_device(ctx, vb, n, n, r, 0, partial));
    return host_verify_finish(partial, 1, ctx->tau_prepared, ok) == 0 ? KZG_B200_OK : KZG_B200_BAD_ARGS;
}

// Phase A of the two-phase form (kzg_b200_verify_phase_a) for a shard that is already in this context's GPU memory.
// zy_out: n x 64 B on the host; commitments_out / proofs_out (optional, n x 48 B each, host): the shard's compressed points,
// which the caller needs on the host anyway for the exchange and for kzg_b200_compute_r.  The decoded points stay in the
// context: kzg_b200_verify_phase_b on the same bytes uses them as they are.
extern "C" int kzg_b200_verify_phase_a_device(kzg_b200_ctx *ctx, const uint8_t *d_blobs, const uint8_t *d_commitments,
                                              const uint8_t *d_proofs, size_t n, uint8_t *zy_out, uint8_t *commitments_out,
                                              uint8_t *proofs_out) {
    if (!ctx || (n && (!d_blobs || !d_commitments || !d_proofs || !zy_out))) return KZG_B200_BAD_ARGS;
    if (!aligned16(d_blobs) || !aligned16(d_commitments) || !aligned16(d_proofs)) return KZG_B200_BAD_ARGS;
    if (n == 0) return KZG_B200_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    VerifyBufs vb;
    HostRecords rec;
    RC(verify_phase_a_device_locked(ctx, d_blobs, d_commitments, d_proofs, n, &vb, nullptr, &rec));
    memcpy(zy_out, rec.zy, 64 * n);
    if (commitments_out) memcpy(commitments_out, rec.commitments, 48 * n);
    if (proofs_out) memcpy(proofs_out, rec.proofs, 48 * n);
    points_digest(ctx->va_digest, rec.commitments, rec.proofs, n);
    ctx->va_n = n;
    ctx->va_valid = true;
    return KZG_B200_OK;
}

// reference verify_kzg_proof (src/kzg.rs:429-445 -> verify_kzg_proof_impl :409-426): no blob, the caller gives
// z and the claimed y.  e(C - [y]G, G2) == e(proof, [tau - z]G2) is, for points of G1, the batch equation with
// one term and r^0 = 1 -- e(proof, [tau]G2) == e(C - [y]G + [z]proof, G2) -- so phase B (which validates the points
// and the scalars) and the final check are reused.
extern "C" int kzg_b200_verify_kzg_proof(kzg_b200_ctx *ctx, const uint8_t commitment[48], const uint8_t z[32],
                                         const uint8_t y[32], const uint8_t proof[48], int *ok) {
    if (!ctx || !commitment || !z || !y || !proof || !ok) return KZG_B200_BAD_ARGS;
    *ok = 0;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    uint8_t zy[64], one[32] = {0}, partial[224];
    memcpy(zy, z, 32);
    memcpy(zy + 32, y, 32);
    one[31] = 1;
    RC(verify_phase_b_locked(ctx, commitment, zy, proof, 1, one, 0, partial));
    return host_verify_finish(partial, 1, ctx->tau_prepared, ok) == 0 ? KZG_B200_OK : KZG_B200_BAD_ARGS;
}

extern "C" int kzg_b200_pairings_verify(const uint8_t a1[48], const uint8_t a2[96], const uint8_t b1[48], const uint8_t b2[96],
                                        int *ok) {
    if (!a1 || !a2 || !b1 || !b2 || !ok) return KZG_B200_BAD_ARGS;
    return host_pairings_verify(a1, a2, b1, b2, ok) == 0 ? KZG_B200_OK : KZG_B200_BAD_ARGS;
}

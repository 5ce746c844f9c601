#!/usr/bin/env python3
"""bench.py -- blobs/sec of the B200 blob path (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of `blob_to_kzg_commitment` over one batch of synthetic blobs
(BASELINE.json configs[3]: 65,536 blobs; reference generator benches/kzg_benches.rs:14-23).
Blobs are independent, so every rank processes its own batch with its own context and there is
no data-path collective (weak scaling); the only collectives of the headline are the barrier / max-reduce of
the timings.

  value      whole-job blobs/s with the blobs already resident in HBM (device entry point)
  e2e        the same through the host-buffer C ABI call (kzg_b200_blob_to_kzg_commitment_batch):
             pinned host blobs -> H2D -> kernels -> D2H of commitments + status, all timed
  parity     every commitment of the last timed step checked on the device by C == [p(tau)] G1 (tau = 1337, the
             bundled testing setup; an independent ladder, kzg_b200_debug_check_tau_identity) on every rank, plus the
             byte comparison with the CPU oracle on a sample (rank 0, N = 1)
  roofline   the MSM kernel (batch_add_kernel, all instantiations) against the measured
             integer-multiply issue rate; per-stage device times come from CUDA events on
             the context's stream (kzg_b200_profile_*)
  also       compute_blob_kzg_proof (device-resident), the strong-scaling leg (BASELINE.json configs[3] as stated:
             65,536 blobs in total, sharded), verify_blob_kzg_proof_batch at 6 / 1024 / 4096 / 16,384 blobs (configs[2]),
             the sharded 16,384-blob verification when N > 1 (configs[4]), and the table-size trade-off
  cpu_baseline / --impl reference
             the CPU restatement (oracle/, a port: the crate's blst path cannot be built here)
             on the box's host cores
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BYTES_PER_BLOB = 131072
# BASELINE.md section 3 (fixed constants): algorithmic integer-multiply work per blob
IMAD_PER_COMMIT = 3.24e8
IMAD_PER_PROOF = 3.34e8
IMAD_PER_VERIFY = 8.4e6
HBM_BYTES_PER_COMMIT = 131120
STAGES = ["digits", "msm_gather", "msm_tree", "compress", "challenge", "eval", "validate", "verify_terms"]
DTYPE = "u32 limbs (381-bit Fp / 255-bit Fr integers)"


def workload_config(blobs):
    """`config` is the same in both arms (the reference arm times a bounded sample of this workload per step)."""
    return {"workload": "mainnet blob_to_kzg_commitment, %d synthetic blobs per GPU per step (BASELINE.json configs[3]), "
                        "trusted_setup.txt" % blobs,
            "blobs_per_gpu_per_step": blobs,
            "l2": "inputs (%.1f GiB per step) exceed L2" % (blobs * BYTES_PER_BLOB / 2 ** 30)}


def read_setup():
    from golden_util import golden
    g = golden()
    return g.g1_bytes, g.g2_bytes


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_traffic(comb_width):
    """DRAM bytes per average batch_add launch, from the committed ncu capture of this code (profiles/ncu_traffic.json,
    written by tools/summarize_ncu.py from `ncu --set full`); None when no capture matches the configuration."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as fh:
        doc = json.load(fh)
    rec = doc.get(str(comb_width))
    return (rec["dram_bytes_per_avg_launch"], rec.get("source")) if rec else (None, None)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for j, nm in enumerate(names) if any(r[3 + j].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].isdigit() else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


def cpu_sample_size(cores):
    return 64 * cores  # some 10-30 s of CPU work per step


def cpu_commit(o, blobs, cores):
    """The CPU arm's path: the fast restatement when the oracle library has it, else the checker path."""
    fast = getattr(o, "blob_to_kzg_commitment_many_fast", None)
    if fast is not None:
        return fast(blobs, nthreads=cores), "fast"
    return o.blob_to_kzg_commitment_many(blobs, nthreads=cores), "checker"


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference's path on the host cores, on a bounded sample of the
    GPU arm's workload per step (the same sample size as the GPU arm's cpu_baseline leg)."""
    if rank != 0:
        return
    from gpu_util import oracle_settings, synthetic_blobs
    cores = os.cpu_count() or 1
    o = oracle_settings("mainnet")
    per_step = cpu_sample_size(cores)
    blobs = synthetic_blobs(per_step, seed=0xB200)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_commit(o, blobs[:cores], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        (out, st), kind = cpu_commit(o, blobs, cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "%d blobs of the workload per step x %d steps, one blob per thread on %d threads" % (per_step, args.steps, cores)
    print(json.dumps({
        "impl": "reference", "metric": "blob_to_kzg_commitment throughput", "value": value, "unit": "blobs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic",
        "config": workload_config(args.blobs),
        "cpu_baseline": {"value": value, "unit": "blobs/s", "cores": cores, "kind": "port", "sample": sample, "path": kind,
                         "note": "CPU restatement (oracle/), not blst: cargo/rustc and the blst sources are absent here"},
        "e2e": {"value": value, "unit": "blobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blobs", type=int, default=int(os.environ.get("KZG_BENCH_BLOBS", 65536)), help="blobs per GPU per step")
    ap.add_argument("--host-pool", type=int, default=65536, help="pinned host blobs per end-to-end call (the user's batch; 8 GiB pinned at 65536)")
    ap.add_argument("--comb-width", type=int, default=0, help="setup points per table group (0 = automatic)")
    ap.add_argument("--no-proof", action="store_true", help="skip the compute_blob_kzg_proof / verification side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the table-size trade-off (other comb widths)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import kzg_rust_b200 as k
    from kzg_rust_b200 import sharded
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG=VERSION (set on some boxes) makes NCCL print a banner on stdout; rank 0 must print ONE JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE", "ABORT"):
            os.environ["NCCL_DEBUG"] = "NONE"  # VERSION and WARN both print the banner (on stdout)
        dist.init_process_group("nccl", device_id=dev)
    L = k.load_library()
    g1, g2 = read_setup()
    t_create = time.time()
    s = k.KzgSettings.load_trusted_setup(g1, g2, local_rank, args.comb_width)
    t_create = time.time() - t_create
    comb_width, table_gb = s.comb_width, round(s.table_bytes / 1e9, 1)
    stream = torch.cuda.ExternalStream(L.kzg_b200_stream(s._h), device=dev)
    B = args.blobs

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_over_ranks(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(ms):
        return reduce_over_ranks(ms, dist.ReduceOp.MAX)

    # synthetic blobs on the device (reference generator: uniform bytes, top byte of each element 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0xB200 + rank)
    blobs = torch.empty((B, 4096, 32), dtype=torch.uint8, device=dev)
    for lo in range(0, B, 4096):
        hi = min(B, lo + 4096)
        blobs[lo:hi] = torch.randint(0, 256, (hi - lo, 4096, 32), dtype=torch.uint8, device=dev, generator=gen)
    blobs[:, :, 0] = 0
    out = torch.zeros((B, 48), dtype=torch.uint8, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    proofs = torch.zeros((B, 48), dtype=torch.uint8, device=dev)

    def commit_device(count=B, first=0):
        rc = L.kzg_b200_blob_to_kzg_commitment_device(s._h, blobs[first:].data_ptr(), count, out[first:].data_ptr(), status[first:].data_ptr())
        if rc:
            raise SystemExit("kzg_b200_blob_to_kzg_commitment_device failed: %d" % rc)

    def proof_device():
        rc = L.kzg_b200_compute_blob_kzg_proof_device(s._h, blobs.data_ptr(), out.data_ptr(), B, proofs.data_ptr(), status.data_ptr())
        if rc:
            raise SystemExit("kzg_b200_compute_blob_kzg_proof_device failed: %d" % rc)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        L.kzg_b200_synchronize(s._h)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.kzg_b200_launch_count(s._h)
        w0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        L.kzg_b200_synchronize(s._h)
        w1 = time.time()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), L.kzg_b200_launch_count(s._h) - l0, (w0, w1)

    def timed_wall(fn, reps):
        """blocking host-side calls (verification ends with a host pairing): wall clock, max over ranks"""
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps
        barrier()
        return max_over_ranks(dt * 1e3)

    clocks = ClockSampler(local_rank) if rank == 0 and not os.environ.get("KZG_BENCH_NO_CLOCKS") else None

    # ---- device-resident leg (value)
    ms, launches, (w0, w1) = timed(commit_device, args.steps, args.warmup)
    if int(status.any().item()):
        raise SystemExit("synthetic blobs were rejected")
    value = world * B * args.steps / (ms * 1e-3)
    clock_summary = clocks.summary(w0, w1) if clocks else None

    # ---- parity of the timed step's output: every commitment on every rank by the tau identity
    ok = torch.zeros(B, dtype=torch.int32, device=dev)
    tau = (1337).to_bytes(32, "big")
    t0 = time.perf_counter()
    rc = L.kzg_b200_debug_check_tau_identity(s._h, blobs.data_ptr(), out.data_ptr(), B, tau, ok.data_ptr())
    if rc:
        raise SystemExit("kzg_b200_debug_check_tau_identity failed: %d" % rc)
    tau_ok = int(reduce_over_ranks(float(ok.sum().item()), dist.ReduceOp.SUM))
    tau_s = time.perf_counter() - t0
    if tau_ok != world * B:
        raise SystemExit("tau identity failed on %d of %d commitments" % (world * B - tau_ok, world * B))
    out_ref = out.clone()

    # one more step with per-stage CUDA-event timing (profiling runs the chunks one at a time, so
    # the stage times do not overlap; the timed steps above keep two chunks in flight)
    L.kzg_b200_profile_enable(s._h, 1)
    prof_ms, _, _ = timed(commit_device, 1, 0)
    stage_ms = (ctypes.c_double * 8)()
    stage_ln = (ctypes.c_uint64 * 8)()
    L.kzg_b200_profile_read(s._h, stage_ms, stage_ln)
    L.kzg_b200_profile_enable(s._h, 0)

    # ---- strong-scaling leg: BASELINE.json configs[3] as stated -- 65,536 blobs in total, contiguous shards
    lo, hi = sharded.shard_range(B, rank, world)
    strong_ms, _, _ = timed(lambda: commit_device(hi - lo, lo), args.steps, 1)
    strong = {"metric": "blob_to_kzg_commitment throughput, %d blobs in total sharded over %d GPU(s)" % (B, world),
              "value": B * args.steps / (strong_ms * 1e-3), "unit": "blobs/s", "scaling": "strong", "steps": args.steps,
              "ms_per_step": strong_ms / args.steps, "blobs_per_gpu_per_step": hi - lo}
    if not bool((out == out_ref).all().item()):
        raise SystemExit("sharded run changed commitments")

    # ---- end-to-end leg: pinned host blobs through the host-buffer C ABI call
    pool = min(args.host_pool, B)
    while True:
        try:
            h_blobs = torch.empty((pool, BYTES_PER_BLOB), dtype=torch.uint8, pin_memory=True)
            break
        except RuntimeError:  # not enough pinnable host memory for one call over the whole batch: use smaller calls
            if pool <= 4096:
                raise
            pool //= 2
    h_blobs.copy_(blobs[:pool].reshape(pool, BYTES_PER_BLOB))
    h_out = torch.empty((pool, 48), dtype=torch.uint8, pin_memory=True)
    h_status = torch.empty(pool, dtype=torch.int32, pin_memory=True)
    calls = (B + pool - 1) // pool

    def commit_host():
        for _ in range(calls):
            rc = L.kzg_b200_blob_to_kzg_commitment_batch(s._h, h_blobs.data_ptr(), pool, h_out.data_ptr(), h_status.data_ptr())
            if rc:
                raise SystemExit("kzg_b200_blob_to_kzg_commitment_batch failed: %d" % rc)

    e2e_ms, _, _ = timed(commit_host, args.steps, min(args.warmup, 1))
    e2e_value = world * calls * pool * args.steps / (e2e_ms * 1e-3)
    same = bool((h_out.to(dev) == out[:pool]).all().item()) and not bool(h_status.any().item())
    if not same:
        raise SystemExit("host-buffer and device-resident paths disagree")

    # ---- compute_blob_kzg_proof side measurement (the other half of BASELINE.json's metric)
    proof = None
    if not args.no_proof:
        pms, _, _ = timed(proof_device, 1, 1)
        if int(status.any().item()):
            raise SystemExit("proof path rejected synthetic blobs")
        proof = {"metric": "compute_blob_kzg_proof throughput", "value": world * B / (pms * 1e-3), "unit": "blobs/s",
                 "steps": 1, "warmup": 1, "ms_per_step": pms, "imad_roofline_frac": None}

    # ---- verify_blob_kzg_proof_batch (BASELINE.json configs[2] and configs[4])
    verify, verify_sharded = None, None
    if proof is not None:
        nv_max = min(16384, B)
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t).numpy()
        vb = pin(blobs[:nv_max].reshape(nv_max, BYTES_PER_BLOB))
        vc, vp = pin(out[:nv_max]), pin(proofs[:nv_max])
        if rank == 0:
            verify = {}
            for n_v, reps in sorted({(6, 10), (min(1024, B), 3), (min(4096, B), 2), (nv_max, 2)}):
                if not k.Kzg.verify_blob_kzg_proof_batch_raw(vb[:n_v], vc[:n_v], vp[:n_v], n_v, s):
                    raise SystemExit("verify_blob_kzg_proof_batch rejected proofs made by the path itself")
                t0 = time.perf_counter()
                for _ in range(reps):
                    k.Kzg.verify_blob_kzg_proof_batch_raw(vb[:n_v], vc[:n_v], vp[:n_v], n_v, s)
                dt = (time.perf_counter() - t0) / reps
                verify["n=%d" % n_v] = {"ms_per_call": dt * 1e3, "blobs_per_s": n_v / dt, "buffers": "pinned host"}
            okc = ctypes.c_int(0)
            call = lambda: L.kzg_b200_verify_blob_kzg_proof_batch_device(s._h, blobs.data_ptr(), out.data_ptr(), proofs.data_ptr(), nv_max, ctypes.byref(okc))
            if call() or not okc.value:
                raise SystemExit("device-resident verification rejected proofs made by the path itself")
            t0 = time.perf_counter()
            for _ in range(3):
                call()
            dt = (time.perf_counter() - t0) / 3
            verify["n=%d device-resident" % nv_max] = {"ms_per_call": dt * 1e3, "blobs_per_s": nv_max / dt, "buffers": "HBM"}
            bad = vp[:6].copy()
            bad[[0, 5]] = bad[[5, 0]]
            verify["negative_control_rejected"] = not k.Kzg.verify_blob_kzg_proof_batch_raw(vb[:6], vc[:6], bad, 6, s)
            # BASELINE.json configs[0]: one call per sample under the reference's `cargo bench` names (benches/kzg_benches.rs:46-126),
            # host buffers in, host results out, median of 30
            def median_ms(fn, reps=30):
                fn()
                ts = []
                for _ in range(reps):
                    t0 = time.perf_counter()
                    fn()
                    ts.append(time.perf_counter() - t0)
                return sorted(ts)[len(ts) // 2] * 1e3
            z1 = vb[0, 32:64].copy()  # a canonical field element
            y1 = k.Kzg.compute_kzg_proof_batch(vb[:1], z1, s)
            single = {
                "blob_to_kzg_commitment": median_ms(lambda: k.Kzg.blob_to_kzg_commitment_batch(vb[:1], s)),
                "compute_kzg_proof": median_ms(lambda: k.Kzg.compute_kzg_proof_batch(vb[:1], z1, s)),
                "compute_blob_kzg_proof": median_ms(lambda: k.Kzg.compute_blob_kzg_proof_batch(vb[:1], vc[:1], s)),
                "verify_kzg_proof": median_ms(lambda: k.Kzg.verify_kzg_proof(vc[0].tobytes(), z1.tobytes(), y1[1][0].tobytes(), y1[0][0].tobytes(), s)),
                "verify_blob_kzg_proof": median_ms(lambda: k.Kzg.verify_blob_kzg_proof_batch_raw(vb[:1], vc[:1], vp[:1], 1, s)),
                "unit": "ms per call (median of 30, host buffers)",
            }
            verify["single_call_ms"] = single
        # configs[4]: one verdict for 16,384 blobs sharded over the ranks (two small all_gathers + one host pairing)
        if world > 1:
            n_tot = nv_max
            be = sharded.CudaBackend(s)
            vlo, vhi = sharded.shard_range(n_tot, rank, world)
            # rank r holds blobs [vlo, vhi) of the batch (its own synthetic blobs, commitments and proofs)
            sb, sc, sp = vb[vlo:vhi], vc[vlo:vhi], vp[vlo:vhi]
            good = sharded.verify_blob_kzg_proof_batch_sharded(be, sb, sc, sp, n_tot, device=dev)
            vms = timed_wall(lambda: sharded.verify_blob_kzg_proof_batch_sharded(be, sb, sc, sp, n_tot, device=dev), 2)
            sp_bad = sp.copy()
            if rank == world - 1:
                sp_bad[[0, 1]] = sp_bad[[1, 0]]
            tampered = sharded.verify_blob_kzg_proof_batch_sharded(be, sb, sc, sp_bad, n_tot, device=dev)
            # the same with every rank's shard resident in its GPU's memory (no upload in phase A)
            db, dc, dp = blobs[vlo:vhi], out[vlo:vhi], proofs[vlo:vhi]
            good_d = sharded.verify_blob_kzg_proof_batch_sharded_device(be, db, dc, dp, n_tot, dev)
            vms_d = timed_wall(lambda: sharded.verify_blob_kzg_proof_batch_sharded_device(be, db, dc, dp, n_tot, dev), 2)
            if not good_d:
                raise SystemExit("sharded verification (device-resident shards) gave the wrong verdict")
            verify_sharded = {"blobs": n_tot, "gpus": world, "ms_per_verdict": vms, "blobs_per_s": n_tot / (vms * 1e-3),
                              "device_resident_shards": {"ms_per_verdict": vms_d, "blobs_per_s": n_tot / (vms_d * 1e-3)},
                              "accepted": bool(good), "tampered_rejected": not tampered,
                              "exchange": "2 all_gathers per verdict: 160 B per blob (C, z, y, proof) + status, then 225 B per rank",
                              "timing": "wall clock around the blocking call (it ends with a host pairing), max over ranks"}
            if not good or tampered:
                raise SystemExit("sharded verification gave the wrong verdict")
    if clocks:
        clocks.stop()

    if rank != 0:
        s.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (batch_add_kernel: gather level + tree levels)
    imad, imadw, fpmul = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    L.kzg_b200_measure_peaks(s._h, ctypes.byref(imad), ctypes.byref(imadw), ctypes.byref(fpmul))
    peaks, peaks_kind = measured_peaks()
    msm_ms = stage_ms[1] + stage_ms[2]
    msm_launches = int(stage_ln[1] + stage_ln[2])
    blobs_timed = B  # the profiled step
    achieved = IMAD_PER_COMMIT * blobs_timed / (msm_ms * 1e-3) if msm_ms > 0 else 0.0
    groups = -(-4096 // comb_width)
    # what this design executes per blob: 255 bit positions x (groups - 1) affine additions x 6 products, plus the
    # Horner pass (254 doublings at 7 products + 254 mixed additions at 11); 600 IMAD issue slots per product
    additions = 255 * (groups - 1)
    executed_imad = (additions * 6 + 254 * 18) * 600.0
    traffic, traffic_src = ncu_traffic(comb_width)
    roofline = {
        "bound": "imad", "kernel": "batch_add_kernel (GatherPolicy + PairPolicy launches)",
        "achieved": achieved / 1e12, "peak": imad.value / 1e12, "unit": "TIMAD/s", "frac": achieved / imad.value if imad.value else None,
        "traffic": traffic, "traffic_unit": "DRAM bytes per average batch_add launch (ncu --set full, 4096-blob chunk)", "traffic_source": traffic_src,
        "launches": msm_launches, "avg_launch_ms": msm_ms / max(1, msm_launches),
        "algorithmic_imad_per_blob": IMAD_PER_COMMIT,
        "peak_source": "mad.lo.u32 issue-rate micro-benchmark run in this process (kzg_b200_measure_peaks); "
                       "MEASURED_PEAKS.json has no integer figure",
        "wide_mac_per_s": imadw.value, "fp_mul_per_s": fpmul.value,
        "whole_step_frac": IMAD_PER_COMMIT * value / world / imad.value if imad.value else None,
        # against the work this design actually executes
        "additions_per_blob": additions, "executed_imad_per_blob": executed_imad,
        "executed_frac": executed_imad * blobs_timed / (msm_ms * 1e-3) / imad.value if imad.value and msm_ms > 0 else None,
        "executed_whole_step_frac": executed_imad * value / world / imad.value if imad.value else None,
        "additions_per_s": additions * blobs_timed / (msm_ms * 1e-3) if msm_ms > 0 else None,
        "hbm": {"achieved_gbs": HBM_BYTES_PER_COMMIT * value / world / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                "frac": HBM_BYTES_PER_COMMIT * value / world / 1e9 / peaks.get("hbm_gbs", 1.0), "peak_kind": peaks_kind},
        "stage_ms_per_step": {STAGES[i]: stage_ms[i] for i in range(8) if stage_ms[i] > 0},
        "profiled_step_ms": prof_ms,
    }
    if proof is not None and imad.value:
        proof["imad_roofline_frac"] = IMAD_PER_PROOF * proof["value"] / world / imad.value
    if verify is not None and imad.value:
        for rec in verify.values():
            if isinstance(rec, dict) and "blobs_per_s" in rec:
                rec["imad_roofline_frac"] = IMAD_PER_VERIFY * rec["blobs_per_s"] / imad.value

    # ---- CPU baseline (bounded sample, N = 1 only) + byte comparison of the GPU output with the checker
    cpu = None
    parity = "%d/%d commitments by the tau identity on the device (%.1f s)" % (tau_ok, world * B, tau_s)
    if not args.no_cpu and world == 1:
        from gpu_util import oracle_settings
        cores = os.cpu_count() or 1
        o = oracle_settings("mainnet")
        nsamp = min(B, cpu_sample_size(cores))
        sample = blobs[:nsamp].reshape(nsamp, BYTES_PER_BLOB).cpu().numpy()
        cpu_commit(o, sample[:cores], cores)
        t0 = time.perf_counter()
        (exp, est), kind = cpu_commit(o, sample, cores)
        dt = time.perf_counter() - t0
        got = out[:nsamp].cpu().numpy()
        if est.any() or not np.array_equal(got, exp):
            raise SystemExit("GPU commitments differ from the CPU oracle on the sampled blobs")
        cpu = {"value": nsamp / dt, "unit": "blobs/s", "cores": cores, "kind": "port", "path": kind,
               "sample": "first %d blobs of the batch, one blob per thread on %d threads (%.1f s)" % (nsamp, cores, dt),
               "parity": "GPU commitments byte-equal on the sample",
               "note": "CPU restatement (oracle/), not blst"}
        parity += " + %d byte-equal with the CPU oracle" % nsamp

    # ---- table size against throughput: the other comb widths on a 16,384-blob batch (N = 1 only)
    tradeoff = None
    if not args.no_sweep and world == 1 and args.comb_width == 0:
        tradeoff = [{"comb_width": comb_width, "table_gb": table_gb, "additions_per_blob": additions, "blobs_per_s": value,
                     "note": "the timed configuration (default rule: the widest comb whose table takes at most half of the free memory)"}]
        s.close()
        s = None
        nb = min(B, 16384)
        for g in (22, 21, 20):
            sg = k.KzgSettings.load_trusted_setup(g1, g2, local_rank, g)
            stg = torch.cuda.ExternalStream(L.kzg_b200_stream(sg._h), device=dev)
            run = lambda: L.kzg_b200_blob_to_kzg_commitment_device(sg._h, blobs.data_ptr(), nb, out.data_ptr(), status.data_ptr())
            run()
            L.kzg_b200_synchronize(sg._h)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stg)
            run()
            run()
            e1.record(stg)
            e1.synchronize()
            L.kzg_b200_synchronize(sg._h)
            tradeoff.append({"comb_width": g, "table_gb": round(sg.table_bytes / 1e9, 1), "additions_per_blob": 255 * (-(-4096 // g) - 1),
                             "blobs_per_s": 2 * nb / (e0.elapsed_time(e1) * 1e-3),
                             "same_bytes": bool((out[:nb] == out_ref[:nb]).all().item())})
            sg.close()

    line = {
        "metric": "blob_to_kzg_commitment throughput", "value": value, "unit": "blobs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(B),
        "b200": {"comb_width": comb_width, "table_gb": table_gb, "additions_per_blob": additions,
                 "sharding": "independent blobs, one context per GPU, no collective", "ctx_create_s": round(t_create, 2)},
        "e2e": {"value": e2e_value, "unit": "blobs/s", "h2d_bytes_per_step": calls * pool * BYTES_PER_BLOB,
                "d2h_bytes_per_step": calls * pool * 52, "ms_per_step": e2e_ms / args.steps,
                "api": "kzg_b200_blob_to_kzg_commitment_batch on pinned host buffers, %d calls of %d blobs per step" % (calls, pool)},
        "gpu_launches": int(launches),
        "parity": parity,
        "clocks": clock_summary,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "also": {"compute_blob_kzg_proof": proof, "strong_scaling": strong, "verify_blob_kzg_proof_batch": verify,
                 "verify_blob_kzg_proof_batch_sharded": verify_sharded, "comb_width_tradeoff": tradeoff},
    }
    print(json.dumps(line), flush=True)
    if s is not None:
        s.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

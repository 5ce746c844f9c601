#!/bin/bash
# compute-sanitizer over the reference's vectors on the small comb (g = 8): memcheck on commitments, blob proofs and
# batch verification, racecheck (shared-memory hazards: the block-wide inversion, the evaluation kernel's scans, the
# warp-cooperative Horner pass) on commitments and blob proofs.  Usage (on the GPU box): bash tools/sanitize.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_$TAG.txt
: > $OUT
run() {
  echo "== compute-sanitizer --tool $1 :: $2 -k '$3'" | tee -a $OUT
  timeout 1500 compute-sanitizer --tool $1 --error-exitcode 9 --print-limit 5 python -m pytest -x -q "$2" -k "$3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|Invalid|hazard|error|Barrier|divergent" | tail -8 | tee -a $OUT
}
run memcheck tests/test_gpu_commit.py "g8"
run memcheck tests/test_gpu_proof.py "compute_blob_kzg_proof_vectors and g8"
run memcheck tests/test_gpu_verify.py "verify_blob_kzg_proof_batch_vectors and g8"
run racecheck tests/test_gpu_commit.py "reference_vectors and g8"
run racecheck tests/test_gpu_proof.py "compute_blob_kzg_proof_vectors and g8"
# the bucket method of phase B (shared-memory tree of k_pip_rows, the counting sort) and the grouped challenge hash
run memcheck tests/test_gpu_verify.py "bucket_method and (6 or 64 or 300)"
run racecheck tests/test_gpu_verify.py "verify_blob_kzg_proof_batch_vectors and g8"
run racecheck tests/test_gpu_verify.py "challenge_and_evaluation"
# the small-call forms (latency comb, one warp / block per sum, three lanes per doubling: shuffles under partial masks) and
# the sliced uploads with the state-carrying hash launches
run memcheck tests/test_gpu_commit.py "small_batch_form and mainnet"
run memcheck tests/test_gpu_verify.py "sliced_upload"
run synccheck tests/test_gpu_commit.py "reference_vectors and g8"
run synccheck tests/test_gpu_verify.py "verify_blob_kzg_proof_batch_vectors and g8"
# the three-lane subgroup checks of phase B on new data (verify_kzg_proof; shuffles under a 30-lane mask)
run memcheck tests/test_gpu_verify.py "verify_kzg_proof_vectors and g8"
run synccheck tests/test_gpu_verify.py "verify_kzg_proof_vectors and g8"

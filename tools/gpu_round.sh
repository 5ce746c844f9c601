#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, one ncu --set full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag] [quick]
TAG=${1:-r2}
QUICK=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.txt
echo "== bench"
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 5000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ -n "$QUICK" ]; then exit 0; fi
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 1500 gpurun_out/bench_ref_$TAG.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-proof --blobs 16384 \
    > gpurun_out/launches_$TAG.log 2>&1
tail -2 gpurun_out/launches_$TAG.log | cut -c1-600
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:batch_add_kernel<kzg::(Gather|Pair)Policy' -s 8 -c 2 -f -o gpurun_out/prof_$TAG \
    python tools/profile_commit.py 0 4096 2 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
ls -la gpurun_out | tail -20

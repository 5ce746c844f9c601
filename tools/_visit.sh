mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2am.txt
KZG_CRITERION_COMB_WIDTH=20 timeout 600 python tools/criterion_bench.py 50 2>&1 | tail -14 | tee gpurun_out/criterion_r2am.jsonl
KZG_B200_COMB_WIDTH=20 timeout 600 python tools/verify_stages.py 64 2>&1 | tail -6 | tee gpurun_out/verify_stages_r2am.txt
for m in 0 64; do KZG_B200_MSM_SMALL_MAX=$m KZG_B200_COMB_WIDTH=20 timeout 600 python tools/verify_stages.py 64 2>&1 | grep "commit host" | sed "s/^/small_max=$m /" | tee -a gpurun_out/small_ab_r2am.txt; done

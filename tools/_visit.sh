mkdir -p gpurun_out
timeout 900 python tools/small_sweep.py 2>&1 | tail -12 | tee gpurun_out/small_sweep_r2ar.txt

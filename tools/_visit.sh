mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2aw.json 2> gpurun_out/bench_r2aw.err
tail -c 3000 gpurun_out/bench_r2aw.json; tail -3 gpurun_out/bench_r2aw.err

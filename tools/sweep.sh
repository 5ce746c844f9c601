#!/bin/bash
# quick knob sweep of the commit path (bench.py value / e2e only); usage: bash tools/sweep.sh "ENV1=.. ENV2=.." ...
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-proof --blobs ${SWEEP_BLOBS:-32768} 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print(json.dumps({'value': round(d['value']), 'e2e': round(d['e2e']['value']), 'c': d['config']['window_bits'], 'stage_ms': {k: round(v,1) for k,v in r['stage_ms_per_step'].items()}}))
    else: print(line)
" | tee -a gpurun_out/sweep.log
done

#!/bin/bash
# A/B runs of the commit path on the GPU box: every argument is "NAME:ENV1=.. ENV2=.." (the environment of one
# bench.py run; KZG_B200_LIB=build/variants/<x>.so selects a differently built library, see tools/build_variants.py).
# Prints one line per run: commitments/s (device-resident), e2e, stage times of the profiled step.
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; cfg=${spec#*:}
  [ "$cfg" = "$spec" ] && cfg=""
  env $cfg timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-proof --no-sweep --blobs ${SWEEP_BLOBS:-32768} 2>&1 | python -c "
import sys, json
name = sys.argv[1]
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line); r = d['roofline']
        print(json.dumps({'variant': name, 'value': round(d['value']), 'e2e': round(d['e2e']['value']), 'g': d['b200']['comb_width'],
                          'clocks': d['clocks']['sm_mhz'] if d.get('clocks') else None,
                          'stage_ms': {k: round(v, 1) for k, v in r['stage_ms_per_step'].items()}}))
    elif line: print(name, line)
" "$name" | tee -a gpurun_out/sweep_${SWEEP_TAG:-r2}.log
done

#!/bin/bash
# GPU iteration: parity tests + full bench summary
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 5 --warmup 3 2>gpurun_out/quick2.err | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'lat', d.get('latency'), 'rounds', d['config'].get('polish_rounds_mean'))
        print('mixed', d['mixed_gait']['value'], 'h30', d['horizon30']['value'], 'full', d['full_step']['value'], 'wbc', d['wbc']['batch_65536']['value'], 'launch', d['config']['launch'])
    else: print(line)
"
tail -3 gpurun_out/quick2.err

#!/bin/bash
# quick GPU check: parity tests + short C2/C3 bench (tag = $1)
TAG=${1:-q}
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for WL in C2 C3; do
  python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_$WL.json 2> gpurun_out/${TAG}_$WL.err
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_$WL.json')); print('$WL', round(d['value']/1e9,2), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, d['greedy_len'], 'frac', round(d['roofline']['frac'],3))"
done

#!/bin/bash
# quick GPU check: parity tests + short C2 / C3 bench (tag = $1)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for WL in C2 C3; do
  python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-sustained > gpurun_out/${TAG}_${WL}.json 2> gpurun_out/${TAG}_${WL}.err
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_${WL}.json')); print('$WL', round(d['value']/1e9,2), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e9,2), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, d['greedy_len'], 'frac', round(d['roofline']['frac'],3), 'pipeline frac', round(d['roofline']['cast_pipeline']['frac'],3), 'parity', d['parity'] and d['parity']['ok'], d['cast_stats'])"
done

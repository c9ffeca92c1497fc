#!/bin/bash
# GPU suite, then brick size / brick entry (prv_set_brick_cull) on C2 / C3 / C1, then ncu on the default.
#   gpurun --timeout 900 -- 'bash tools/brick_sweep.sh'
O=gpurun_out/brick
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
for WL in C2 C3 C1; do
  for CFG in "8 0" "8 1" "4 0" "4 1" "16 1"; do
    set -- $CFG
    TAG=b$1e$2
    python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-sustained --brick $1 --brick-entry $2 > $O/${WL}_$TAG.json 2> $O/${WL}_$TAG.err
    python -c "
import json; d=json.load(open('$O/${WL}_$TAG.json')); print('$WL brick=$1 entry=$2', round(d['value']/1e9,2),'Grays/s', round(d['ms_per_step'],4),'ms  e2e', round(d['e2e']['value']/1e9,2), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, 'marched', d['cast_stats']['marched'], 'probes', d['cast_stats']['probes_in'], 'frac', round(d['roofline']['frac'],3))"
  done
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"march_kernel|coarse_kernel|cull_kernel" -s 9 -c 3 -o $O/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/prof.log 2>&1

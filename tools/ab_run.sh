#!/bin/bash
# A/B on one box: tools/ab_run.sh "NAME[:ENV=VAL][:--bench-flag=v]..."   e.g.  tools/ab_run.sh A B C D D:PRV_DEPROJ_TABLE=0 R:--brick-entry=1
mkdir -p gpurun_out/ab2
for spec in "$@"; do
  NAME=${spec%%:*}; ENVS=""; FLAGS=""
  IFS=':' read -ra PARTS <<< "$spec"
  for part in "${PARTS[@]:1}"; do
    if [[ "$part" == --* ]]; then FLAGS="$FLAGS ${part/=/ }"; else ENVS="$ENVS $part"; fi
  done
  for WL in C2 C3; do
    TAG=${spec//[:=]/_}_$WL
    env PRV_B200_LIB=$PWD/ab/$NAME.so $ENVS python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-sustained $FLAGS > gpurun_out/ab2/$TAG.json 2> gpurun_out/ab2/$TAG.err
    python -c "
import json; d=json.load(open('gpurun_out/ab2/$TAG.json')); k=d['kernel_ms_per_step']; print('$spec $WL step', round(d['ms_per_step'],4), 'cull', round(k['cull_ms'],4), 'march', round(k['march_ms'],4), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity'] and d['parity']['ok'], 'probes', d['cast_stats']['probes_in'], 'marched', d['cast_stats']['marched'])" 2>&1 | tail -1
  done
done

#!/bin/bash
# A/B of the staging options (prv_set_staging) on C2 / C3, then ncu on C3.  (The march-kernel shape variants of profiles/r2_staging_ab.md
# -- 2 probes in flight, 128-thread blocks, coarse kernel at 6 blocks/SM -- were compile-time switches that have been removed.)
O=gpurun_out/ab
mkdir -p $O
run() { # tag workload extra-args...
  TAG=$1; WL=$2; shift 2
  python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-sustained "$@" > $O/${WL}_$TAG.json 2> $O/${WL}_$TAG.err
  python -c "
import json; d=json.load(open('$O/${WL}_$TAG.json')); print('$WL $TAG', round(d['value']/1e9,2),'Grays/s', round(d['ms_per_step'],4),'ms', {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if k in ('cull_ms','march_ms','greedy_ms')}, 'frac', round(d['roofline']['frac'],3), 'parity', d['parity'] and d['parity']['ok'])"
}
for WL in C2 C3; do
  run base $WL
  run smem $WL --stage-smem 1
  run l2 $WL --stage-l2 1
  run nosmem $WL --stage-smem 0
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"march_kernel|coarse_kernel|cull_kernel" -s 9 -c 3 -o $O/prof_c3 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-sustained --no-parity > $O/prof_c3.log 2>&1

#!/bin/bash
# one ncu --set full capture of march_kernel (and coarse_kernel) on C3 and C2 -> gpurun_out/ncu/
O=gpurun_out/ncu
mkdir -p $O
for WL in C3 C2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"march_kernel|coarse_kernel" -s 4 -c 2 -o $O/prof_$WL -f python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-sustained --no-parity > $O/prof_$WL.log 2>&1
  ls -la $O/prof_$WL.ncu-rep
done

#!/bin/bash
# round-end evidence on one B200: ncu launch list + full capture (C3 headline and C2), traffic per launch, API breakdown
O=gpurun_out/fin
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -3
for WL in C3 C2; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$WL.csv python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-sustained --no-parity > $O/launches_$WL.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"march_kernel|coarse_kernel|cull_kernel|greedy_cluster_kernel|popcount_rows_kernel" -s 15 -c 5 -o $O/prof_$WL python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-sustained --no-parity > $O/prof_$WL.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"splat_resolve_kernel|splat_points_kernel" -c 2 -o $O/prof_splat python -c "
import sys; sys.path.insert(0,'.')
import bench, load_pkg, argparse
prv=load_pkg.load(); synth=bench._synth()
print(bench.measure_c5_splat(prv, synth, argparse.Namespace(steps=1), 0))" > $O/prof_splat.log 2>&1
python tools/e2e_breakdown.py C2 > $O/api_breakdown_C2.txt 2>&1
python tools/e2e_breakdown.py C3 > $O/api_breakdown_C3.txt 2>&1

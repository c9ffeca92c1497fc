#!/bin/bash
# One gpurun call that settles the opt-in second cull level (prv_set_fine_cull):
# exactness on the device, sanitizers on the new kernels, then default vs --fine-cull {4,2,1} on C1 / C2 / C3.
#   gpurun --timeout 900 -- 'bash tools/fine_cull_run.sh'
O=gpurun_out/fine
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
PRV_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q 2>&1 | tail -5 > $O/pytest_experimental.txt; cat $O/pytest_experimental.txt
PRV_TEST_EXPERIMENTAL=1 timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_experimental.py -m gpu -q -x -k "160" > $O/sanitizer_memcheck.log 2>&1; tail -3 $O/sanitizer_memcheck.log
PRV_TEST_EXPERIMENTAL=1 timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_experimental.py -m gpu -q -x -k "160" > $O/sanitizer_racecheck.log 2>&1; tail -3 $O/sanitizer_racecheck.log
for WL in C2 C3 C1; do
  for CFG in "0" "4" "2" "1" "4 --fine-entry" "2 --fine-entry" "1 --fine-entry"; do
    TAG=$(echo $CFG | tr -d ' -')
    python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --fine-cull $CFG > $O/${WL}_fine$TAG.json 2> $O/${WL}_fine$TAG.err
    python -c "
import json; d=json.load(open('$O/${WL}_fine$TAG.json')); print('$WL fine_cull=$CFG', round(d['value']/1e9,2),'Grays/s', round(d['ms_per_step'],4),'ms  e2e', round(d['e2e']['value']/1e9,2), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, 'marched', d['cast_stats']['marched'], 'probes', d['cast_stats']['probes_in'], 'frac', round(d['roofline']['frac'],3))"
  done
done
# the winner's march kernel under ncu (instruction counts per line: tools/ncu_lines.py)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"march_entry_kernel|coarse_fine_kernel" -s 6 -c 2 -o $O/prof_fine python bench.py --steps 1 --warmup 3 --no-cpu-baseline --fine-cull 2 --fine-entry > $O/prof_fine.log 2>&1

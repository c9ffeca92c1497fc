#!/bin/bash
# round-end measurement set on one B200: tests, every workload, reference arm, ncu launch list + full capture, API breakdown, sanitizers
O=gpurun_out/fin
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
python bench.py > $O/bench_C2.json 2> $O/bench_C2.err
for WL in C1 C3 C4 C5; do python bench.py --workload $WL --steps 20 --no-cpu-baseline > $O/bench_$WL.json 2> $O/bench_$WL.err; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python tools/e2e_breakdown.py C2 > $O/api_breakdown_C2.txt 2>&1
python tools/e2e_breakdown.py C5 > $O/api_breakdown_C5.txt 2>&1
python tools/greedy_probe.py > $O/greedy_probe.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"march_kernel|greedy_cluster_kernel|cull_kernel|coarse_kernel" -s 8 -c 4 -o $O/prof_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/prof_full.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -q -x -k "golden or greedy or small" > $O/sanitizer_memcheck.log 2>&1; tail -3 $O/sanitizer_memcheck.log
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -q -x -k "golden or greedy" > $O/sanitizer_racecheck.log 2>&1; tail -3 $O/sanitizer_racecheck.log
for f in C1 C2 C3 C4 C5; do python -c "
import json; d=json.load(open('$O/bench_$f.json')); print('$f', round(d['value']/1e9,2),'Grays/s', round(d['ms_per_step'],4),'ms e2e', round(d['e2e']['value']/1e9,2), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],3), d.get('clocks'))"; done
cat $O/bench_reference.json | cut -c1-400

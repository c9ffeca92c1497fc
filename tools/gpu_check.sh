#!/bin/bash
# one GPU call: smoke + whole GPU suite + default bench (C3 headline + C2 / C5 / ensemble in the same line) + short reference arm + sanitizers
O=gpurun_out/chk
mkdir -p $O
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/chk/bench.json'))
print('C3', round(d['value']/1e9,2),'Grays/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value']/1e9,2), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],3), 'parity', d['parity'] and d['parity']['ok'], 'sustained', d['sustained'])
c=d['c2']; print('C2', round(c['value']/1e9,2), round(c['ms_per_step'],4), 'e2e', round(c['e2e']['value']/1e9,2), {k:round(v,4) for k,v in c['kernel_ms_per_step'].items()}, 'frac', round(c['roofline']['frac'],3), 'greedy', c['roofline_greedy']['frac'])
print('C5', d['c5_splat']['ms_per_step'], d['c5_splat']['roofline_splat']['frac'], d['c5_splat']['kernel_ms_per_step'])
print('ens', d['ensemble_scoring'])
print('cpu', d['cpu_baseline'], d.get('cpu_reference_structure'))
PY
python bench.py --impl reference --steps 2 --warmup 0 --ref-seconds 3 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_golden.py tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -q -x -k "golden or greedy or small or brick or ingest" > $O/sanitizer_memcheck.log 2>&1; tail -4 $O/sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -q -x -k "golden or greedy or brick" > $O/sanitizer_racecheck.log 2>&1; tail -4 $O/sanitizer_racecheck.log

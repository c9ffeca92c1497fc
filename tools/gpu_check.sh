#!/bin/bash
# one GPU call: whole GPU suite + default bench (C3 headline + C2 / C5 in the same line) + reference arm, short
O=gpurun_out/chk
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/chk/bench.json'))
print('C3', round(d['value']/1e9,2),'Grays/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value']/1e9,2), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],3), 'parity', d['parity'] and d['parity']['ok'], 'sustained', d['sustained'])
c=d['c2']; print('C2', round(c['value']/1e9,2), round(c['ms_per_step'],4), 'e2e', round(c['e2e']['value']/1e9,2), {k:round(v,4) for k,v in c['kernel_ms_per_step'].items()}, 'frac', round(c['roofline']['frac'],3), 'greedy', c['roofline_greedy']['frac'])
print('C5', d['c5_splat'])
print('cpu', d['cpu_baseline'], d.get('cpu_reference_structure'))
PY

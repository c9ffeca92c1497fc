timeout 600 python -m pytest tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -5
for B in 8 4; do python bench.py --workload C3 --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-sustained --brick $B > gpurun_out/b$B.json 2>gpurun_out/b$B.err; python -c "
import json; d=json.load(open('gpurun_out/b$B.json')); print('C3 brick $B', round(d['value']/1e9,2), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()}, d['parity']['ok'])"; done
for B in 8 4; do python bench.py --workload C2 --steps 50 --warmup 3 --no-cpu-baseline --no-extras --no-sustained --brick $B > gpurun_out/c2b$B.json 2>gpurun_out/c2b$B.err; python -c "
import json; d=json.load(open('gpurun_out/c2b$B.json')); print('C2 brick $B', round(d['value']/1e9,2), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3))"; done

#!/bin/bash
# 1/2/4/8-GPU scaling of bench.py on one box: weak (C2, one object per GPU) and strong (C3, 1024 views sharded + NCCL all-gather)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  for WL in C2 C3; do
    if [ $N -eq 1 ]; then
      python bench.py --gpus 1 --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${WL}_n${N}.json 2> gpurun_out/scale_${WL}_n${N}.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${WL}_n${N}.json 2> gpurun_out/scale_${WL}_n${N}.err
    fi
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${WL}_n${N}.json"))
    print("${WL} N=${N}: %.2f Grays/s  %.3f ms/step  scaling=%s  e2e=%.2f  greedy_len=%d seq_head=%s" % (d["value"]/1e9, d["ms_per_step"], d["scaling"], d["e2e"]["value"]/1e9, d["greedy_len"], d["greedy_seq"][:6]))
except Exception as e:
    print("${WL} N=${N}: FAILED", e)
PY
  done
done

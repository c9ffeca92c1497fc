#!/bin/bash
# 1/2/4/8-GPU strong scaling of the default bench (C3, 1024 views sharded, peer-memory all-gather) on one box, + the NCCL transport at 8
mkdir -p gpurun_out/scale
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/scale/n$N.json 2> gpurun_out/scale/n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale/n$N.json 2> gpurun_out/scale/n$N.err
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29699 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --gather nccl > gpurun_out/scale/n8_nccl.json 2> gpurun_out/scale/n8_nccl.err
python - <<'PY'
import json
base=None
for tag in ("n1","n2","n4","n8","n8_nccl"):
    try:
        d=json.load(open("gpurun_out/scale/%s.json"%tag))
        if base is None: base=d["value"]
        pr=d.get("per_rank") or []
        print("%s: %.2f Grays/s  %.3f ms/step  x%.2f  parity=%s  sustained %.3f ms  e2e=%.2f  cast max %.3f  gather %s" % (tag, d["value"]/1e9, d["ms_per_step"], d["value"]/base, d["parity"] and d["parity"]["ok"], d["sustained"]["ms_per_step"], d["e2e"]["value"]/1e9, max([r["cast_ms"] for r in pr] or [0]), [round(r["allgather_ms"],3) for r in pr]))
    except Exception as e:
        print(tag, "FAILED", e)
PY

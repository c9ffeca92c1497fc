"""N > 1 GPU parity gate (skipped with fewer than 2 devices; run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

The multi-GPU CUDA path -- views sharded interleaved over the ranks, NCCL all-gather of the coverage rows on the scoring
stream, replicated cluster greedy -- against the oracle's frozen full-size C3 vectors (tests/golden/golden_c3.json): on EVERY
rank the gathered table re-ordered by view id, the counts and the greedy sequence / gains / covered mask must be the oracle's."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("gather", ["p2p", "nccl"])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_view_sharded_c3_matches_the_oracle_on_every_rank(n, gather):
    """gather = p2p: rows stored into every rank's table over NVLink peer memory by the count kernel (PRV_CAST_PUBLISH), flags
    with release / acquire; gather = nccl: ncclAllGather on the scoring stream."""
    if _gpus() < n:
        pytest.skip("needs %d GPUs" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1", "--master-port",
           str(29700 + n + (10 if gather == "nccl" else 0)), os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "3", "--warmup", "3",
           "--no-cpu-baseline", "--no-sustained", "--gather", gather]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["config"]["views"] == 1024
    p = d["parity"]
    assert p["ok"] and p["ranks_ok"] == [True] * n and all(p["checks"].values()), p
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_c3.json")))["cases"][0]
    assert d["greedy_seq"] == golden["greedy_seq"]
    assert len(d["per_rank"]) == n and all(r_["allgather_ms"] > 0 for r_ in d["per_rank"])
    assert ("peer-memory" in d["config"]["sharding"]) == (gather == "p2p")
    assert sum(r_["local_views"] for r_ in d["per_rank"]) == 1024


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 4])
def test_small_sharded_paths(n):
    """Every transport of the coverage rows (peer-memory stores fused into the count kernel, peer-memory stores from the scoring
    stream, ncclAllGather), uneven view split with padding rows, five pipelined steps at a time: gathered table, counts and greedy
    against the oracle on every rank (tests/multi_worker.py)."""
    if _gpus() < n:
        pytest.skip("needs %d GPUs" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1", "--master-port",
           str(29750 + n), os.path.join(ROOT, "tests", "multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and ("MULTI_WORKER_OK world=%d" % n) in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])

"""bench.py contract that can be checked without a GPU: the reference arm's JSON line, and that the own arm fails loudly
(no silent CPU fallback) when there is no B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "2", "--warmup", "0",
                        "--ref-seconds", "0.5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rays_per_sec" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # the reference arm needs nothing of the product: workload through oracle.HostShim, libprv_b200.so never mapped
    assert d["product_library_loaded"] is False


def test_default_workload_is_the_north_star_c3_at_every_n():
    """`bench.py --gpus N` (what the driver runs, no other flags) must select the 1024-view workload C3 for N = 1 and N > 1, so
    the driver's 1 -> 8 curve is strong scaling on the north-star workload, with the golden-vector parity gate in the line."""
    import re
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert re.search(r'add_argument\("--workload", default="C3"', src)
    assert os.path.exists(os.path.join(ROOT, "tests", "golden", "golden_c3.json"))
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_c3.json")))["cases"][0]
    assert g["n_views"] == 1024 and g["size"] == [1280, 960] and g["rays"] == 1024 * 1280 * 960 and len(g["row_sha16"]) == 1024
    assert len(g["greedy_seq"]) == len(g["greedy_gain"]) >= 2 and g["greedy_seq"][0] == 0


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--workload", "C1"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr

"""Packs a few of the reference's Hemisphere/<N>.txt view sets (input fixtures) into one JSON
fixture so tests and bench.py can run where /root/reference does not exist (the GPU box).

    python tests/golden/make_hemisphere_fixture.py   # needs /root/reference, rewrites hemisphere_sets.json

Tokens are kept as the decimal strings of the files so that parsing them reproduces the doubles the
reference's `ifstream >> double` loader (Share_Data.hpp:517-528) sees.
"""
import json
import os

SRC = "/root/reference/PRV_simulation/Hemisphere"
SETS = [3, 5, 32, 100]
out = {}
for n in SETS:
    toks = open(os.path.join(SRC, "%d.txt" % n)).read().split()
    assert len(toks) == 3 * n, (n, len(toks))
    out[str(n)] = [toks[3 * i:3 * i + 3] for i in range(n)]
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hemisphere_sets.json")
with open(dst, "w") as f:
    json.dump({"source": "psc0628/NeRF-PRV PRV_simulation/Hemisphere/<N>.txt", "sets": out}, f, indent=0)
print("wrote", dst, {k: len(v) for k, v in out.items()})

#!/bin/bash
# tools/shard_ab.sh NAME...: per-rank share of C3 at N = 8 / 4 on one GPU for each A/B build
for NAME in "$@"; do for N in 8 4; do PRV_B200_LIB=$PWD/ab/$NAME.so python tools/shard_probe.py $N 0 2>&1 | tail -1; done; done

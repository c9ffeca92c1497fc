#!/bin/bash
# A/B builds of the device library: tools/ab_build.sh NAME SRC_DIR [extra nvcc flags]  ->  ab/NAME.so  (select with PRV_B200_LIB=ab/NAME.so)
# SRC_DIR holds csrc/ + host/ + ../include like nerf-prv_b200 (e.g. a `git worktree` of another commit)
set -e
NAME=$1; SRC=$2; shift 2
mkdir -p ab
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -shared -o ab/$NAME.so "$@" $SRC/csrc/prv_device.cu $SRC/csrc/prv_host.cpp -ldl
echo ab/$NAME.so

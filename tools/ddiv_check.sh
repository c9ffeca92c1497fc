#!/bin/bash
# GPU self-test of ddiv_pair against __ddiv_rn (tests/cuda/test_ddiv_pair.cu), 2^28 triples per class
nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -I nerf-prv_b200/csrc -I include -o /tmp/test_ddiv_pair tests/cuda/test_ddiv_pair.cu 2>&1 | grep -v warning
/tmp/test_ddiv_pair ${1:-28}

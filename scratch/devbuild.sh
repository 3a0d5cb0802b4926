#!/bin/bash
# Fast developer build: only one FFT length (default 4096).  Run __graft_entry__.build(force=True)
# before tests / commits.
N=${1:-4096}; shift
cd "$(dirname "$0")/.."
time nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
  -DSB_DEV_N=$N "$@" -o ${OUT:-scarplet_b200/libscarplet_b200.so} scarplet_b200/csrc/sb_lib.cu

#!/bin/bash
# Development loop: builds tools/lab_infer.cu (the product kernel source + -D switches) in several variants here,
# runs them all on the GPU box in one call.  usage: tools/lab.sh build | run
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -lcuda"
if [ "$1" == "build" ]; then
  shift
  i=0
  rm -f tools/lab_infer_v*
  while [ $# -gt 0 ]; do
    unset CC CXX
    $NV $1 tools/lab_infer.cu -o tools/lab_infer_v$i 2>&1 | grep -v "^$" | head -5 &
    echo "v$i: $1" ; i=$((i+1)); shift
  done
  wait
  ls tools/lab_infer_v*
else
  mkdir -p gpurun_out
  { nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
    for f in tools/lab_infer_v*; do echo "== $f"; timeout 60 $f; timeout 60 $f | head -1; done; } 2>&1 | tee gpurun_out/lab.txt
fi

#!/bin/bash
# Runs the sm_100a probe on the GPU box; every GEMM case in its own process so one trap does not hide the rest.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
  for i in $(seq 0 11); do timeout 60 tools/probe_umma case $i || echo "case $i exit=$?"; done
  timeout 120 tools/probe_umma bench
} > gpurun_out/probe.txt 2>&1
cat gpurun_out/probe.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/probe_dw_rate 2>&1 | tee gpurun_out/probe_dw_rate.txt

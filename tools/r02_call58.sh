#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:nrc_infer_kernel -s 2 -c 1 -o gpurun_out/r02f_infer_records -f python tools/prof_infer_records.py > gpurun_out/ncu_ir.log 2>&1; tail -2 gpurun_out/ncu_ir.log; ls -la gpurun_out/r02f_infer_records.ncu-rep

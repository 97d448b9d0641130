#!/bin/bash
# producer warps / spare input buffer in the training kernel: tests, timings, traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python tools/lab_train.py one 2>&1 | tail -3
timeout 300 python tools/bench_records.py 2>&1 | tail -3
timeout 60 tools/trace_grad 1048576 1 1 260 > gpurun_out/trace_train_prod_unpacked.txt 2>&1; head -4 gpurun_out/trace_train_prod_unpacked.txt
timeout 60 tools/trace_grad 1048576 1 0 260 > gpurun_out/trace_train_prod_encoded.txt 2>&1; head -4 gpurun_out/trace_train_prod_encoded.txt

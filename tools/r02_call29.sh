#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/lab_train.py run 2>&1 | tee gpurun_out/lab_train_tgt.txt
timeout 60 tools/trace_grad 1048576 1 0 600 > gpurun_out/trace_train_tgt_encoded.txt 2>&1; head -4 gpurun_out/trace_train_tgt_encoded.txt

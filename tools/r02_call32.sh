#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/trace_grad 1048576 1 0 700 > gpurun_out/trace_train_ifence_encoded.txt 2>&1; head -4 gpurun_out/trace_train_ifence_encoded.txt

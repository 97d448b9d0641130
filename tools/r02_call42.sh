#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/trace_grad 16384 4 1 700 > gpurun_out/trace_train_frame_final.txt 2>&1; head -4 gpurun_out/trace_train_frame_final.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/lab_train.py run 2>&1 | tee gpurun_out/lab_train_skipstores.txt

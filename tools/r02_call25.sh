#!/bin/bash
# ReLU masks in registers / deferred forward smem copy: tests on the default build (both on), timings of the four variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python tools/lab_train.py run 2>&1 | tee gpurun_out/lab_train_regmask.txt

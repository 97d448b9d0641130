#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 600 python tools/lab_train.py run 2>&1 | tee gpurun_out/lab_train_ifence.txt

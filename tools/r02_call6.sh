#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -12
python tools/lab_train.py run 2>&1 | tail -8
python tools/lab_train.py run 2>&1 | tail -8

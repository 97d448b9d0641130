#!/bin/bash
mkdir -p gpurun_out
NRC_B200_LIB=$PWD/tools/lab_lib_ts.so timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python tools/lab_train.py run 2>&1 | grep -E "^==|frame"

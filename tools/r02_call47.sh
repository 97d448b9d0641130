#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python tools/lab_train.py run 2>&1; done | tee gpurun_out/lab_train_f16da.txt
echo "== GPU tests on the fp16-dA variant"
NRC_B200_LIB=$PWD/tools/lab_lib_f16da.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_v2.py tests/test_gpu_records.py tests/test_gpu_frame.py -q -m gpu 2>&1 | tail -12 | tee -a gpurun_out/lab_train_f16da.txt
echo "== product build, full-size property test"
timeout 600 python -m pytest tests/test_gpu_records.py -q -m gpu -k full_size 2>&1 | tail -2

#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
for tool in memcheck initcheck synccheck; do
timeout 900 compute-sanitizer --tool $tool --target-processes all python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py tests/test_gpu_records.py -q -m gpu -k "not 1080p and not full_size and not learn_an_image and not sweep and not loss_curve" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tail -2
done
timeout 300 python tools/lab_train.py one 2>&1 | tail -1

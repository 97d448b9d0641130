#!/bin/bash
for so in tools/lab_lib_v*.so; do
  for rep in 1 2; do
    echo "== $so memcheck #$rep"
    NRC_B200_LIB=$PWD/$so timeout 600 compute-sanitizer --tool memcheck --target-processes all python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_tile_encoded_gradient_pipeline" 2>&1 | grep -E "passed|failed" | tail -1
  done
done
echo "== head build without sanitizer, 20 repetitions"
for i in $(seq 1 20); do NRC_B200_LIB=$PWD/tools/lab_lib_v0head.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_tile_encoded_gradient_pipeline" 2>&1 | grep -E "passed|failed" | tail -1; done | sort | uniq -c

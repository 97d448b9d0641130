#!/bin/bash
# compute-sanitizer on the round-2 final build: memcheck / synccheck / initcheck over the parity + record + golden tests (long ones
# deselected), racecheck on the inference and encode kernels
mkdir -p gpurun_out
SEL='not 1080p and not full_size and not learn_an_image and not sweep and not loss_curve'
for tool in memcheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --target-processes all python -m pytest tests/test_gpu_parity.py tests/test_gpu_records.py tests/test_gpu_golden_v2.py -q -m gpu -x -k "$SEL" > gpurun_out/san_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_$tool.txt | tail -3
done
timeout 900 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_gpu_records.py tests/test_gpu_golden_v2.py -q -m gpu -k "infer or encoder" > gpurun_out/san_racecheck.txt 2>&1
echo "== racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_racecheck.txt | tail -3

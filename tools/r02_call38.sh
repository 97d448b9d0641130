#!/bin/bash
mkdir -p gpurun_out
for so in tools/lab_lib_kou*.so; do echo "== $so"; NRC_B200_LIB=$PWD/$so timeout 300 python tools/lab_gather.py 2>&1 | tail -1; done | tee gpurun_out/lab_gather_knockouts.txt
echo "== product build"; timeout 300 python tools/lab_gather.py 2>&1 | tail -1
timeout 900 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_gpu_records.py tests/test_gpu_golden_v2.py -q -m gpu -k "infer or encoder" > gpurun_out/san_racecheck.txt 2>&1
echo "== racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_racecheck.txt | tail -3
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python tools/lab_train.py one 2>&1 | tail -1

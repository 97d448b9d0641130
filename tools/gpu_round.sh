#!/bin/bash
# One GPU-box round trip: tests, smoke, bench, ncu launch list and full captures of the two dominant kernels.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json
if [ "$1" == "prof" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:nrc_infer_kernel -s 3 -c 1 -o gpurun_out/prof_infer -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:nrc_train_kernel -s 2 -c 1 -o gpurun_out/prof_train -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
  ls -la gpurun_out
fi

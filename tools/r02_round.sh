#!/bin/bash
# One single-GPU round trip on the product build: GPU tests, smoke, bench (+ reference arm), launch list, ncu captures, sweep.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python tools/sweep.py > gpurun_out/sweep_1gpu.txt 2>&1; tail -22 gpurun_out/sweep_1gpu.txt
if [ "$1" == "prof" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:nrc_train_kernel -s 2 -c 1 -o gpurun_out/r02_train_2p22_ts -f python tools/prof_train.py 22 unpacked 2 > gpurun_out/ncu_t1.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:nrc_infer_kernel -s 3 -c 1 -o gpurun_out/r02_infer_enc -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_i1.log 2>&1
fi
python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; head -c 600 gpurun_out/bench_reference.json

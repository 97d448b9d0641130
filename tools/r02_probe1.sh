#!/bin/bash
# round 2, call 1: state of the round-1 build on a fresh box - GPU tests, ncu of the training kernel at the throughput sizes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
python tools/prof_train.py 22 unpacked 3
python tools/prof_train.py 22 encoded 3
python tools/prof_train.py 20 unpacked 5
ncu --set full --clock-control none --import-source on -k regex:nrc_train_kernel -s 2 -c 1 -o gpurun_out/r02_train_2p22_unpacked -f python tools/prof_train.py 22 unpacked 2 > gpurun_out/ncu_t1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nrc_train_kernel -s 2 -c 1 -o gpurun_out/r02_train_2p22_encoded -f python tools/prof_train.py 22 encoded 2 > gpurun_out/ncu_t2.log 2>&1
tail -3 gpurun_out/ncu_t1.log gpurun_out/ncu_t2.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -3

#!/bin/bash
# 2-GPU box: GPU tests, bench at 1 and 2 GPUs, A/B of frame time between the round-1 and the current library
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench1 rc=$?"; tail -3 gpurun_out/bench_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench_2gpu.err
python tools/lab_train.py run 2>&1 | grep -E "^==|frame"
python tools/lab_train.py run 2>&1 | grep -E "^==|frame"

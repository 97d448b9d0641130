#!/bin/bash
# 8-GPU round trip on the round-2 final build
mkdir -p gpurun_out
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; grep -v "^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_${N}gpu.err | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/sweep.py > gpurun_out/sweep_${N}gpu.txt 2>&1; grep -v "^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/sweep_${N}gpu.txt | tail -18
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tests/mgpu_worker.py 2>&1 | grep MGPU_RESULT > gpurun_out/mgpu_${N}gpu.txt; cut -c1-400 gpurun_out/mgpu_${N}gpu.txt

#!/bin/bash
# 2-GPU: all GPU tests (incl. the multi-GPU worker at world 2) on the new exchange (multicast / grid-independent counts / timeout flag)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_worker.py > gpurun_out/mgpu2.log 2>&1; echo "worker rc=$?"
grep MGPU_RESULT gpurun_out/mgpu2.log || tail -30 gpurun_out/mgpu2.log

#!/bin/bash
# N-GPU: the multi-GPU worker (all exchange modes) at world N
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tests/mgpu_worker.py > gpurun_out/mgpu$N.log 2>&1; echo "worker rc=$?"
grep MGPU_RESULT gpurun_out/mgpu$N.log || tail -30 gpurun_out/mgpu$N.log

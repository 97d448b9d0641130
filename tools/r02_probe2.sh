#!/bin/bash
# round 2, 2-GPU call: correctness of the fused NVLink exchange on the current build + multicast capability probe
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -20
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py 2>&1 | grep -E "MGPU_RESULT|Error|error" | head -5
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/probe_multicast.py > gpurun_out/probe_multicast.log 2>&1
grep -E "\[probe\]|NVLS|nvls" gpurun_out/probe_multicast.log | head -40

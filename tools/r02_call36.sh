#!/bin/bash
# 2-GPU validation of the final build: multi-GPU worker tests (all three exchange set-ups), bench at 2 GPUs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_2gpu.json"))
print("value", d["value"], "e2e", d["e2e"]); print("multi", json.dumps(d["multi_gpu"])); print("train", json.dumps(d["train"]))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tests/mgpu_worker.py 2>&1 | grep MGPU_RESULT > gpurun_out/mgpu_2gpu.txt; cat gpurun_out/mgpu_2gpu.txt

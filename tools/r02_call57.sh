#!/bin/bash
mkdir -p gpurun_out
N=4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_4gpu.json"))
print("value %.3e ms %.4f e2e %.3e ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print(json.dumps(d["multi_gpu"])); print(d["train"]["frame_us"], d["train"]["step_2p20"]["us_per_step"])
PY

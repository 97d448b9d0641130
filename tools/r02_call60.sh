#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value %.3e ms %.4f frac %.3f e2e ms %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
print(d["train"]["frame_us"], d["train"]["frame_us_from_train_records"], d["train"]["step_2p20"]["us_per_step"], d["train"]["step_2p20"]["tflops_per_gpu"], d["train"]["step_2p20"]["roofline"]["frac"])
print({k:round(v["us"],1) for k,v in d["stages"].items()}, {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["extra"].items() if "us" in k or "ms_per" in k})
PY

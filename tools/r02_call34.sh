#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -c 600 gpurun_out/bench_r02b.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02b.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"])
print("stages", json.dumps(d["stages"]))
print("train", json.dumps(d["train"]))
print("extra", json.dumps(d["extra"]))
PY

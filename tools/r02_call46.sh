#!/bin/bash
# exactly what the driver runs at round end, with wall-clock times
mkdir -p gpurun_out
t0=$(date +%s); python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3; echo "pytest wall $(( $(date +%s) - t0 )) s"
t0=$(date +%s); python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke wall $(( $(date +%s) - t0 )) s"
t0=$(date +%s); python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -1 gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['steps'], d['warmup'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks'])"

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python tools/bench_records.py 2>&1 | tail -4

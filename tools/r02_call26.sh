#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15
examples/render_loop 48 2>&1 | tail -5

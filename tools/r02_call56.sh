#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_golden_v2.py -q -m gpu -k "standalone" 2>&1 | tail -3

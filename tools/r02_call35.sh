#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_golden_v2.py -q -m gpu -k "standalone_encoder" 2>&1 | tail -5

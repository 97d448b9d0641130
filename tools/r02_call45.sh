#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 600 python tools/lab_train.py run 2>&1; done

#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 300 python tools/lab_train.py one 2>&1 | tail -1

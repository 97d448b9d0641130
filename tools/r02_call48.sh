#!/bin/bash
for i in 1 2; do timeout 600 python tools/lab_train.py run 2>&1; done

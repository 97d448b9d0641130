#!/bin/bash
# runs tools/bench_records.py against every variant library tools/lab_lib_*.so (built by tools/lab_train.py build ...)
for so in tools/lab_lib_*.so; do echo "== $so"; NRC_B200_LIB=$PWD/$so python tools/bench_records.py 2>&1 | tail -5; done

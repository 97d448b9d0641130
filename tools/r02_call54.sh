#!/bin/bash
for so in tools/lab_lib_*.so; do echo "== $so"; NRC_B200_LIB=$PWD/$so timeout 300 python tools/lab_gather.py 2>&1 | tail -1; done

#!/bin/bash
echo "== early mask load + OLD (CTA-wide) drains under memcheck"; NRC_B200_LIB=$PWD/tools/lab_lib_early_olddrain.so timeout 600 compute-sanitizer --tool memcheck --target-processes all python tools/probe_repro.py 2>&1 | grep -E "n=|ERROR"

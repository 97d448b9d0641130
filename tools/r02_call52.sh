#!/bin/bash
for so in tools/lab_lib_h*.so; do echo "== $so"; NRC_B200_LIB=$PWD/$so timeout 300 python tools/probe_e2e_host.py 2>&1 | tail -3 | head -2 | tail -1; done
timeout 300 python tools/probe_e2e_host.py 2>&1 | tail -1

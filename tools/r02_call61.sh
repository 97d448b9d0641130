#!/bin/bash
lscpu | grep -E "Model name|MHz" | head -3
timeout 300 python tools/probe_e2e_host.py 2>&1 | tail -4

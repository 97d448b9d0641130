#!/bin/bash
timeout 300 python tools/probe_e2e_host.py 2>&1 | tail -6

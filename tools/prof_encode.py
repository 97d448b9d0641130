"""Profiling driver for ncu: a few launches of the stand-alone encode stage on a 1080p frame of 14-float records."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc

n = 1920 * 1080
rec = torch.rand((n, 14), device="cuda")
out = torch.empty((n, 64), device="cuda", dtype=torch.float16)
for _ in range(4):
    nrc.encode_inputs(rec, out=out)
torch.cuda.synchronize()

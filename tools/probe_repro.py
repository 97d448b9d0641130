"""Development probe: is the multi-tile encoded gradient bit-reproducible run to run, and if not, how do two runs differ?"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc
g = torch.Generator(device="cuda").manual_seed(77)
w = torch.from_numpy((np.random.default_rng(55).standard_normal(20672) * np.sqrt(2 / 64)).astype(np.float16)).cuda()
for n in (148 * 128 * 3, 148 * 128 * 5 + 128 * 37 + 19):
    x = torch.rand((n, 64), device="cuda", generator=g).half()
    t = torch.rand((n, 3), device="cuda", generator=g).half()
    runs = []
    for _ in range(4):
        dw = torch.zeros(nrc.WEIGHT_COUNT, dtype=torch.float32, device="cuda")
        nrc.mlp_gradient_encoded(w, dw, x, t)
        runs.append(dw.cpu().numpy())
    ref = runs[0]
    for i, r in enumerate(runs[1:], 1):
        d = np.abs(r - ref)
        nz = np.nonzero(d)[0]
        print(f"n={n} run {i} vs run 0: {nz.size} elements differ, max |diff| {d.max():.3e} (max |value| {np.abs(ref).max():.3e})",
              "layers:", sorted(set((nz // 4096).tolist()))[:8], "first idx:", nz[:6].tolist())

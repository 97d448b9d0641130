"""Profiling driver for ncu: a few launches of nrc_infer (20-byte eval records + scene gather + screen composite) on a 1080p frame."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc
from vknrc_b200 import synth

st = nrc.NrcState(0, (1920, 1080), seed=1)
sa = synth.make_scene_arrays(7, n_prims=20000, n_instances=8, n_materials=64, n_textures=8)
sc = nrc.DeviceScene(sa["vertices"], sa["vertex_indices"], sa["texcoords"], sa["texcoord_indices"], sa["materials"], sa["material_ids"], sa["transforms"], sa["textures"])
n = 1920 * 1080
ev = torch.from_numpy(synth.eval_records_screen(11, 1920, 1080, 20000, 8).view(np.uint8).reshape(-1)).cuda()
bf, gb = torch.rand((1080, 1920, 4), device="cuda"), torch.rand((1080, 1920, 2), device="cuda")
trs = [torch.zeros(nrc.TRAIN_BATCH_SIZE * 40, dtype=torch.uint8, device="cuda") for _ in range(4)]
cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
for _ in range(4):
    st.infer(ev, cnt, sc, bf, gb, 1920, trs, max_count=n)
torch.cuda.synchronize()

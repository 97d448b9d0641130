"""Times the record-format paths (packed 16/20/40-byte records + scene gather) with and without the per-primitive rows."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc
from vknrc_b200 import synth

def timed(fn, steps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

st = nrc.NrcState(0, (1920, 1080), seed=1)
sa = synth.make_scene_arrays(7, n_prims=20000, n_instances=8, n_materials=64, n_textures=8)
args = (sa["vertices"], sa["vertex_indices"], sa["texcoords"], sa["texcoord_indices"], sa["materials"], sa["material_ids"], sa["transforms"], sa["textures"])
pi = synth.prim_instance_ids(20000, 8)
n = 1920 * 1080
ev = torch.from_numpy(synth.eval_records_screen(11, 1920, 1080, 20000, 8).view(np.uint8).reshape(-1)).cuda()
out = torch.empty((n, 3), dtype=torch.float16, device="cuda")
bf, gb = torch.rand((1080, 1920, 4), device="cuda"), torch.rand((1080, 1920, 2), device="cuda")
trs = [torch.zeros(nrc.TRAIN_BATCH_SIZE * 40, dtype=torch.uint8, device="cuda") for _ in range(4)]
cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
nb = nrc.TRAIN_BATCH_SIZE
trec = [torch.from_numpy(synth.train_records(100 + b, nb, 20000, 8).view(np.uint8).reshape(-1)).cuda() for b in range(4)]
big = torch.from_numpy(synth.train_records(200, 1 << 20, 20000, 8).view(np.uint8).reshape(-1)).cuda()
for name, kw in (("prim_table", dict()), ("index buffers only", dict(prim_table=False))):
    sc = nrc.DeviceScene(*args, **kw)
    t1 = timed(lambda: st.infer_packed(ev[4:], sc, outputs=out, stride_bytes=20, max_count=n))
    t2 = timed(lambda: st.infer(ev, cnt, sc, bf, gb, 1920, trs, max_count=n))
    t3 = timed(lambda: st.train_frame(trec, sc, max_count=nb))
    t4 = timed(lambda: st.train_batch(big, sc, max_count=1 << 20), steps=8, warm=2)
    print(f"{name:20s}: infer_packed {t1:7.1f} us | nrc_infer (scatter) {t2:7.1f} us | train_frame (records) {t3:6.1f} us | train 2^20 records {t4:7.1f} us")

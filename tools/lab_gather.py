"""Development tool: times the gather paths (stand-alone unpack, unpack + encode, fused nrc_infer) of a (variant) build."""
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
sc = nrc.DeviceScene(sa["vertices"], sa["vertex_indices"], sa["texcoords"], sa["texcoord_indices"], sa["materials"], sa["material_ids"], sa["transforms"], sa["textures"])
n = 1920 * 1080
ev = torch.from_numpy(synth.eval_records_screen(11, 1920, 1080, 20000, 8).view(np.uint8).reshape(-1)).cuda()
out = torch.empty((n, 3), dtype=torch.float16, device="cuda")
enc = torch.empty((n, 64), dtype=torch.float16, device="cuda")
bf, gb = torch.rand((1080, 1920, 4), device="cuda"), torch.rand((1080, 1920, 2), device="cuda")
trs = [torch.zeros(nrc.TRAIN_BATCH_SIZE * 40, dtype=torch.uint8, device="cuda") for _ in range(4)]
cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
t0 = timed(lambda: nrc.unpack_inputs(ev[4:], sc, stride_bytes=20, n=n))
t1 = timed(lambda: nrc.encode_packed_inputs(ev[4:], sc, stride_bytes=20, n=n, out=enc))
t2 = timed(lambda: st.infer_packed(ev[4:], sc, outputs=out, stride_bytes=20, max_count=n))
t3 = timed(lambda: st.infer(ev, cnt, sc, bf, gb, 1920, trs, max_count=n))
print(f"unpack {t0:6.1f} us | unpack+encode {t1:6.1f} us | infer_packed {t2:6.1f} us | nrc_infer (scatter) {t3:6.1f} us")

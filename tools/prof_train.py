"""Runs a few training steps of one size / input mode (profiling driver for ncu): prof_train.py LOG2N MODE [STEPS]
MODE: unpacked (14-float records + fused Adam) | encoded (pre-encoded gradient only) | records (40-byte NRCTrainRecord + scene gather)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc

e, mode = int(sys.argv[1]), sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = 1 << e
st = nrc.NrcState(0, (1920, 1080), seed=1)
g = torch.Generator(device="cuda").manual_seed(1)
if mode == "encoded":
    x = torch.rand((n, 64), device="cuda", generator=g).half()
    t16 = torch.rand((n, 3), device="cuda", generator=g).half()
    fn = lambda: st.gradient_encoded(x, t16)
elif mode == "unpacked":
    rec = torch.rand((n, 14), device="cuda", generator=g)
    tgt = torch.rand((n, 3), device="cuda", generator=g)
    fn = lambda: st.train_batch_unpacked(rec, tgt, write_use_weights=True)
else:
    from vknrc_b200 import synth
    sa = synth.make_scene_arrays(7, n_prims=20000, n_instances=8, n_materials=64, n_textures=8)
    scene = nrc.DeviceScene(sa["vertices"], sa["vertex_indices"], sa["texcoords"], sa["texcoord_indices"], sa["materials"],
                            sa["material_ids"], sa["transforms"], sa["textures"], device=0)
    d = torch.from_numpy(synth.train_records(3, n, 20000, 8).view(np.uint8).reshape(-1)).cuda()
    fn = lambda: st.train_batch(d, scene, max_count=n)
for _ in range(steps):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    fn()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / steps * 1e-3
print(f"train 2^{e} {mode}: {t*1e6:.1f} us/step, {n/t:.3e} records/s, {n*115840/t/1e12:.1f} TFLOP/s")

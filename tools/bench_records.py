"""Development tool: where does the time of the record-format inference path go? Times, on the 1080p workload, the stand-alone
unpack kernel, packed -> fp16 outputs, unpacked -> scatter, and the full nrc_infer (packed -> scatter)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc
from vknrc_b200 import synth

def timed(fn, steps=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

n = 1920 * 1080
st = nrc.NrcState(0, (1920, 1080), seed=1)
sa = synth.make_scene_arrays(7, n_prims=20000, n_instances=8, n_materials=64, n_textures=8)
scene = nrc.DeviceScene(sa["vertices"], sa["vertex_indices"], sa["texcoords"], sa["texcoord_indices"], sa["materials"], sa["material_ids"], sa["transforms"], sa["textures"])
ev = synth.eval_records_screen(11, 1920, 1080, 20000, 8)
d_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).cuda()
d_bf = torch.rand((1080, 1920, 4), device="cuda"); d_gb = torch.rand((1080, 1920, 2), device="cuda")
d_trs = [torch.zeros(16384 * 40, dtype=torch.uint8, device="cuda") for _ in range(4)]
out = torch.empty((n, 3), dtype=torch.float16, device="cuda")
unp = nrc.unpack_inputs(d_ev[4:], scene, stride_bytes=20, n=n)
dst = torch.from_numpy(np.ascontiguousarray(ev["dst"])).cuda()
print("unpack kernel alone            %.1f us" % timed(lambda: nrc.unpack_inputs(d_ev[4:], scene, stride_bytes=20, n=n)))
print("unpacked -> fp16 out           %.1f us" % timed(lambda: st.infer_unpacked(unp, outputs=out)))
print("packed   -> fp16 out           %.1f us" % timed(lambda: st.infer_packed(d_ev[4:], scene, outputs=out, stride_bytes=20, max_count=n)))
print("unpacked -> scatter            %.1f us" % timed(lambda: st.infer_scatter_unpacked(dst, unp, None, d_bf, d_gb, 1920, d_trs)))
print("nrc_infer (packed -> scatter)  %.1f us" % timed(lambda: st.infer(d_ev, None, scene, d_bf, d_gb, 1920, d_trs, max_count=n)))

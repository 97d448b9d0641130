"""Development probe: host-side enqueue time vs device time of the e2e call (nrc_infer_eval_records_host)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc
from vknrc_b200 import synth

n = 1920 * 1080
st = nrc.NrcState(0, (1920, 1080), seed=1)
sa = synth.make_scene_arrays(7, n_prims=20000, n_instances=8, n_materials=64, n_textures=8)
scene = nrc.DeviceScene(sa["vertices"], sa["vertex_indices"], sa["texcoords"], sa["texcoord_indices"], sa["materials"], sa["material_ids"], sa["transforms"], sa["textures"], device=0)
ev = synth.eval_records_screen(11, 1920, 1080, 20000, 8)
h_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).pin_memory()
hout = torch.empty((n, 3), dtype=torch.float16).pin_memory()
d_stage = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
for rep in range(3):
    for _ in range(3):
        st.infer_eval_records_host(h_ev, scene, hout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    e0.record()
    for _ in range(20):
        t0 = time.perf_counter()
        st.infer_eval_records_host(h_ev, scene, hout)
        host.append(time.perf_counter() - t0)
    e1.record()
    torch.cuda.synchronize()
    print(f"e2e {e0.elapsed_time(e1)/20*1e3:7.1f} us/step | host enqueue per call: median {np.median(host)*1e6:6.1f} us, max {np.max(host)*1e6:6.1f} us")
# plain one-shot H2D of the same bytes, for the PCIe rate of this box
for _ in range(3):
    d_stage.copy_(h_ev, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    d_stage.copy_(h_ev, non_blocking=True)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e-3
print(f"plain H2D of {n*20/1e6:.1f} MB: {t*1e6:.1f} us = {n*20/t/1e9:.1f} GB/s")
# the same loop with an NVML sampler thread like bench.py's ClockSampler running beside it (5 ms period)
import threading
import pynvml
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = threading.Event()
def poll(period):
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)
        time.sleep(period)
for period in (0.005, 0.05):
    stop.clear()
    th = threading.Thread(target=poll, args=(period,), daemon=True)
    th.start()
    for rep in range(2):
        for _ in range(12):
            st.infer_eval_records_host(h_ev, scene, hout)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            st.infer_eval_records_host(h_ev, scene, hout)
        e1.record()
        torch.cuda.synchronize()
        print(f"with NVML sampler every {period*1e3:.0f} ms: e2e {e0.elapsed_time(e1)/20*1e3:7.1f} us/step")
    stop.set()
    th.join()

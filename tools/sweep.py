"""BASELINE config 5 on one GPU: inference from 2^20 to 2^26 pre-encoded queries per step and training steps of 2^18..2^22
records (pre-encoded gradient + 14-float records with the fused optimizer). Prints a table (profiles/r01_sweep.txt)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e-3


st = nrc.NrcState(0, (1920, 1080), seed=1)
g = torch.Generator(device="cuda").manual_seed(1)
print("inference, pre-encoded queries (41 344 FLOP, 134 B per query)")
print(f"{'queries':>10} {'us/step':>10} {'queries/s':>12} {'TFLOP/s':>9} {'HBM GB/s':>9}")
for e in range(20, 27):
    n = 1 << e
    x = torch.rand((n, 64), device="cuda", generator=g).half()
    out = torch.empty((n, 3), device="cuda", dtype=torch.float16)
    t = timed(lambda: st.infer_encoded(x, out, clamp=True), max(3, min(50, (1 << 27) // n)))
    print(f"2^{e:<8} {t * 1e6:10.1f} {n / t:12.3e} {n * 41344 / t / 1e12:9.1f} {n * 134 / t / 1e9:9.0f}")
    del x, out
print("training (115 840 FLOP per record)")
print(f"{'records':>10} {'enc grad us':>12} {'TFLOP/s':>9} {'records+Adam us':>16} {'records/s':>12} {'TFLOP/s':>9}")
for e in range(14, 23, 2):
    n = 1 << e
    x = torch.rand((n, 64), device="cuda", generator=g).half()
    t16 = torch.rand((n, 3), device="cuda", generator=g).half()
    rec = torch.rand((n, 14), device="cuda", generator=g)
    tgt = torch.rand((n, 3), device="cuda", generator=g)
    steps = max(3, min(40, (1 << 24) // n))
    te = timed(lambda: st.gradient_encoded(x, t16), steps)
    tr = timed(lambda: st.train_batch_unpacked(rec, tgt, write_use_weights=True), steps)
    print(f"2^{e:<8} {te * 1e6:12.1f} {n * 115840 / te / 1e12:9.1f} {tr * 1e6:16.1f} {n / tr:12.3e} {n * 115840 / tr / 1e12:9.1f}")
# BASELINE config 2: learn-an-image (test/mlp_learning_an_image): 16384 random-uv samples per SGD step, 640x640 inference per frame
img = torch.randint(0, 256, (512, 512, 4), dtype=torch.uint8, device="cuda", generator=g)
st2 = nrc.NrcState(0, (640, 640), seed=2)
out = torch.empty((640, 640, 4), dtype=torch.uint8, device="cuda")
seed = [0]


def image_frame():
    seed[0] += 1
    st2.image_train_step(img, 1234 + seed[0], 99 + 7 * seed[0])
    st2.image_infer(640, out)


ts = timed(lambda: st2.image_train_step(img, 5, 6), 50)
ti = timed(lambda: st2.image_infer(640, out), 50)
tf = timed(image_frame, 50)
print("learn-an-image (config 2)")
print(f"train step (16384 samples, SGD) {ts * 1e6:7.1f} us | inference 640x640 {ti * 1e6:7.1f} us ({409600 / ti:.3e} queries/s) | frame (step + inference) {tf * 1e6:7.1f} us = {1 / tf:7.0f} frames/s")

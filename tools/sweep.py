"""BASELINE config 5: inference from 2^20 to 2^26 pre-encoded queries per step and training steps of 2^18..2^22 records,
data-parallel over the GPUs of the job (the GLOBAL step size is split by index range; training all-reduces the gradient
inside the kernel). One process per GPU:
    python tools/sweep.py                                                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sweep.py
Prints a table (profiles/r02_sweep_<N>gpu.txt); times are CUDA events, max over ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc  # noqa: E402
from vknrc_b200.dist import shard_range  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


st = nrc.NrcState(local, (1920, 1080), seed=1)
exchange = "-"
if world > 1:
    try:
        exchange = "multimem.st (NVLS)" if st.comm_attach_symmetric() else "unicast (symmetric memory)"
    except Exception:
        st.comm_connect()
        exchange = "unicast (CUDA IPC)"
g = torch.Generator(device=dev).manual_seed(1 + rank)
say(f"config 5 sweep on {world} GPU(s), global step sizes split by index range; gradient exchange: {exchange}")
say("inference, pre-encoded queries (41 344 FLOP, 134 B per query)")
say(f"{'queries':>10} {'us/step':>10} {'queries/s':>12} {'TFLOP/s all':>12} {'TFLOP/s/GPU':>12} {'HBM GB/s/GPU':>13}")
for e in range(20, 27):
    n_glob = 1 << e
    lo, hi = shard_range(n_glob, rank, world, align=128)
    n = hi - lo
    x = torch.rand((n, 64), device=dev, generator=g).half()
    out = torch.empty((n, 3), device=dev, dtype=torch.float16)
    t = timed(lambda: st.infer_encoded(x, out, clamp=True), max(3, min(50, (1 << 27) // n_glob)))
    say(f"2^{e:<8} {t * 1e6:10.1f} {n_glob / t:12.3e} {n_glob * 41344 / t / 1e12:12.1f} {n * 41344 / t / 1e12:12.1f} {n * 134 / t / 1e9:13.0f}")
    del x, out
say("training (115 840 FLOP per record): 14-float records, gradient + in-kernel all-reduce + Adam in one launch")
say(f"{'records':>10} {'us/step':>10} {'records/s':>12} {'TFLOP/s all':>12} {'TFLOP/s/GPU':>12}")
for e in range(14, 23, 2):
    n_glob = 1 << e
    lo, hi = shard_range(n_glob, rank, world, align=128)
    n = hi - lo
    rec = torch.rand((max(n, 1), 14), device=dev, generator=g)
    tgt = torch.rand((max(n, 1), 3), device=dev, generator=g)
    steps = max(3, min(40, (1 << 24) // n_glob))
    tr = timed(lambda: st.train_batch_unpacked(rec, tgt, write_use_weights=True, max_count=n), steps)
    say(f"2^{e:<8} {tr * 1e6:10.1f} {n_glob / tr:12.3e} {n_glob * 115840 / tr / 1e12:12.1f} {n * 115840 / tr / 1e12:12.1f}")
    del rec, tgt
if world > 1:
    st.comm_status()
if world == 1:
    # BASELINE config 2: learn-an-image (test/mlp_learning_an_image): 16384 random-uv samples per SGD step, 640x640 inference per frame
    img = torch.randint(0, 256, (512, 512, 4), dtype=torch.uint8, device=dev, generator=g)
    st2 = nrc.NrcState(local, (640, 640), seed=2)
    out = torch.empty((640, 640, 4), dtype=torch.uint8, device=dev)
    seed = [0]

    def image_frame():
        seed[0] += 1
        st2.image_train_step(img, 1234 + seed[0], 99 + 7 * seed[0])
        st2.image_infer(640, out)

    ts = timed(lambda: st2.image_train_step(img, 5, 6), 50)
    ti = timed(lambda: st2.image_infer(640, out), 50)
    tf = timed(image_frame, 50)
    say("learn-an-image (config 2)")
    say(f"train step (16384 samples, SGD) {ts * 1e6:7.1f} us | inference 640x640 {ti * 1e6:7.1f} us ({409600 / ti:.3e} queries/s) | frame (step + inference) {tf * 1e6:7.1f} us = {1 / tf:7.0f} frames/s")
if world > 1:
    dist.destroy_process_group()

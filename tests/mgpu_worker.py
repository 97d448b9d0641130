"""Multi-GPU worker (launched by tests/test_multi_gpu.py or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_worker.py
Every rank owns a shard of each training batch. Checks that the all-reduce fused into the training kernel (peer-mapped
inboxes over NVLink) gives (1) bit-identical weights / optimizer state on all ranks, (2) the same result as the split
path gradient -> NCCL all-reduce -> nrc_adam_step (bit-exact for 2 ranks, where the sum order cannot differ), and
(3) the same result as ONE GPU training on the concatenated batch, within the gradient tolerance. Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vknrc_b200 as nrc  # noqa: E402
from util import he_weights, random_records  # noqa: E402
from vknrc_b200.dist import shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    w32 = he_weights(5)
    n_global = 4 * 4096 * world // world * world  # divisible by world
    frames = 2
    # the global batches (same on every rank), each rank trains on its index range
    recs = [random_records(100 + b, n_global) for b in range(4)]
    tgts = [np.random.default_rng(200 + b).uniform(0, 1, (n_global, 3)).astype(np.float32) for b in range(4)]
    lo, hi = shard_range(n_global, rank, world, align=128)
    d_recs = [torch.from_numpy(r[lo:hi].copy()).cuda() for r in recs]
    d_tgts = [torch.from_numpy(t[lo:hi].copy()).cuda() for t in tgts]
    n_local = hi - lo

    def same_on_all_ranks(arr: np.ndarray) -> bool:
        t = torch.from_numpy(arr.view(np.uint8).copy()).cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return all(torch.equal(out[0], o) for o in out)

    # ---- fused: in-kernel NVLink all-reduce, one launch per frame
    st = nrc.NrcState(local, (64, 64), seed=1)
    st.set_weights(w32)
    st.set_use_ema_weights(True)
    st.comm_connect()
    assert st.comm_world() == world
    for _ in range(frames):
        st.train_frame_unpacked(d_recs, d_tgts, max_count=n_local)
    fused = st.download()
    ok_replicated = all(same_on_all_ranks(np.ascontiguousarray(fused[k])) for k in ("weights", "use_weights", "optimizer_entries", "gradients"))

    # ---- split: gradient -> NCCL all-reduce -> Adam
    st2 = nrc.NrcState(local, (64, 64), seed=1)
    st2.set_weights(w32)
    st2.set_use_ema_weights(True)
    gt = st2.gradient_tensor()
    for _ in range(frames):
        for b in range(4):
            st2.gradient_unpacked(d_recs[b], d_tgts[b], max_count=n_local)
            dist.all_reduce(gt)
            st2.adam_step(write_use_weights=(b == 3))
    split = st2.download()
    bit_equal_split = all(np.array_equal(np.ascontiguousarray(fused[k]).view(np.uint8), np.ascontiguousarray(split[k]).view(np.uint8))
                          for k in ("weights", "use_weights", "optimizer_entries", "gradients"))
    wdiff_split = float(np.abs(fused["weights"].astype(np.float32) - split["weights"].astype(np.float32)).max())
    count_ok = fused["gradients"][nrc.GRAD_COUNT_SLOT] == n_global and int(fused["optimizer_state"]["t"]) == 4 * frames

    # ---- one GPU on the whole batch (rank 0 only): same maths, different summation tree
    wdiff_single = 0.0
    if rank == 0:
        st3 = nrc.NrcState(local, (64, 64), seed=1)
        st3.set_weights(w32)
        st3.set_use_ema_weights(True)
        full_r = [torch.from_numpy(r).cuda() for r in recs]
        full_t = [torch.from_numpy(t).cuda() for t in tgts]
        for _ in range(frames):
            st3.train_frame_unpacked(full_r, full_t)
        single = st3.download()
        wdiff_single = float(np.abs(fused["weights"].astype(np.float32) - single["weights"].astype(np.float32)).max())
        st3.close()

    # ---- timing: frame time with the fused exchange vs the split NCCL path (CUDA events, max over ranks)
    def timed(fn, steps=30, warm=5):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def split_frame():
        for b in range(4):
            st2.gradient_unpacked(d_recs[b], d_tgts[b], max_count=n_local)
            dist.all_reduce(gt)
            st2.adam_step(write_use_weights=(b == 3))
    ms_fused = timed(lambda: st.train_frame_unpacked(d_recs, d_tgts, max_count=n_local))
    ms_split = timed(split_frame)

    st.comm_shutdown()
    res = {"world": world, "n_global_per_batch": n_global, "replicated_bit_identical": bool(ok_replicated),
           "fused_equals_nccl_split_bitwise": bool(bit_equal_split), "max_weight_diff_vs_split": wdiff_split,
           "max_weight_diff_vs_single_gpu": wdiff_single, "count_ok": bool(count_ok), "ms_frame_fused": ms_fused, "ms_frame_nccl_split": ms_split}
    ok = ok_replicated and count_ok and (bit_equal_split if world == 2 else wdiff_split <= 2e-3) and wdiff_single <= 4e-3
    res["ok"] = bool(ok)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(res), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

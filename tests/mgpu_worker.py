"""Multi-GPU worker (launched by tests/test_multi_gpu.py, bench.py's driver-visible checks, or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_worker.py
Every rank owns a shard of each training batch (UNEQUAL shards: the ranks launch different grids). For each way of setting
up the exchange - CUDA-IPC inboxes + unicast stores ("ipc"), a symmetric allocation + unicast stores ("symm"), the same
allocation through its NVSwitch multicast mapping, one multimem.st per word ("multicast") - checks that the all-reduce fused
into the training kernel gives (1) bit-identical weights / optimizer state / gradients on all ranks, (2) the same result as
the split path gradient -> NCCL all-reduce -> nrc_adam_step (bit-exact for 2 ranks, where the sum order cannot differ),
(3) the same result as ONE GPU training on the concatenated batch within the gradient tolerance, and (4) that a peer which
never shows up is reported as NRC_ERR_PEER_TIMEOUT instead of hanging or killing the context. Prints one JSON line."""
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

KEYS = ("weights", "use_weights", "optimizer_entries", "gradients")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    w32 = he_weights(5)
    n_global = 4 * 4096 * world + 5 * 128  # NOT divisible: the last rank's shard is smaller -> the ranks launch different grids
    frames = 2
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

    # ---- split: gradient -> NCCL all-reduce -> Adam (the stock-collective version of the same frames)
    st2 = nrc.NrcState(local, (64, 64), seed=1)
    st2.set_weights(w32)
    st2.set_use_ema_weights(True)
    gt = st2.gradient_tensor()

    def split_frame():
        for b in range(4):
            st2.gradient_unpacked(d_recs[b], d_tgts[b], max_count=n_local)
            dist.all_reduce(gt)
            st2.adam_step(write_use_weights=(b == 3))
    for _ in range(frames):
        split_frame()
    split = st2.download()

    res = {"world": world, "n_global_per_batch": n_global, "shard_sizes_differ": bool(n_global % (world * 128) != 0), "modes": {}}
    ok = True
    fused_ref = None
    for mode in ("ipc", "symm", "multicast"):
        st = nrc.NrcState(local, (64, 64), seed=1)
        st.set_weights(w32)
        st.set_use_ema_weights(True)
        if mode == "ipc":
            st.comm_connect()
        else:
            try:
                has_mc = st.comm_attach_symmetric(multicast=(mode == "multicast"))
            except Exception as e:  # no symmetric memory on this box: say so, do not pass silently
                res["modes"][mode] = {"unavailable": repr(e)[:200]}
                continue
            if mode == "multicast" and not has_mc:
                res["modes"][mode] = {"unavailable": "no multicast mapping (NVLS) on this box"}
                continue
        assert st.comm_world() == world
        for _ in range(frames):
            st.train_frame_unpacked(d_recs, d_tgts, max_count=n_local)
        st.comm_status()
        fused = st.download()
        m = {"replicated_bit_identical": bool(all(same_on_all_ranks(np.ascontiguousarray(fused[k])) for k in KEYS))}
        m["equals_nccl_split_bitwise"] = bool(all(np.array_equal(np.ascontiguousarray(fused[k]).view(np.uint8), np.ascontiguousarray(split[k]).view(np.uint8)) for k in KEYS))
        m["max_weight_diff_vs_split"] = float(np.abs(fused["weights"].astype(np.float32) - split["weights"].astype(np.float32)).max())
        m["count_ok"] = bool(fused["gradients"][nrc.GRAD_COUNT_SLOT] == n_global and int(fused["optimizer_state"]["t"]) == 4 * frames)
        if fused_ref is None:
            fused_ref = fused
        m["equals_first_mode_bitwise"] = bool(all(np.array_equal(np.ascontiguousarray(fused[k]).view(np.uint8), np.ascontiguousarray(fused_ref[k]).view(np.uint8)) for k in KEYS))
        m["ms_frame"] = timed(lambda: st.train_frame_unpacked(d_recs, d_tgts, max_count=n_local))
        st.comm_status()
        m["ok"] = bool(m["replicated_bit_identical"] and m["count_ok"] and m["equals_first_mode_bitwise"]
                       and (m["equals_nccl_split_bitwise"] if world == 2 else m["max_weight_diff_vs_split"] <= 2e-3))
        ok = ok and m["ok"]
        res["modes"][mode] = m
        dist.barrier()
        st.comm_shutdown()
        st.close()
    res["ms_frame_nccl_split"] = timed(split_frame)

    # ---- one GPU on the whole batch (rank 0 only): same maths, different summation tree
    wdiff_single = 0.0
    if rank == 0 and fused_ref is not None:
        st3 = nrc.NrcState(local, (64, 64), seed=1)
        st3.set_weights(w32)
        st3.set_use_ema_weights(True)
        full_r = [torch.from_numpy(r).cuda() for r in recs]
        full_t = [torch.from_numpy(t).cuda() for t in tgts]
        for _ in range(frames):
            st3.train_frame_unpacked(full_r, full_t)
        single = st3.download()
        wdiff_single = float(np.abs(fused_ref["weights"].astype(np.float32) - single["weights"].astype(np.float32)).max())
        st3.close()
    res["max_weight_diff_vs_single_gpu"] = wdiff_single
    ok = ok and wdiff_single <= 4e-3 and fused_ref is not None

    # ---- a peer that never shows up: the last rank skips one training call; everybody else must get NRC_ERR_PEER_TIMEOUT
    # from nrc_comm_status (no hang, no trap, context intact), the skipping rank stays clean
    dist.barrier()
    st = nrc.NrcState(local, (64, 64), seed=1)
    st.comm_connect()
    st.comm_set_timeout(1 << 12)
    timeout_ok = True
    if rank != world - 1:
        st.train_frame_unpacked(d_recs, d_tgts, max_count=n_local)
        try:
            st.comm_status()
            timeout_ok = False
        except nrc.NrcError as e:
            timeout_ok = "-5" in str(e)
        y = st.infer_unpacked(d_recs[0][:256])  # the context is still usable
        torch.cuda.synchronize()
        timeout_ok = timeout_ok and bool(torch.isfinite(y.float()).all())
    else:
        st.comm_status()
    t = torch.tensor([1 if timeout_ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    res["peer_timeout_reported"] = bool(int(t.item()))
    ok = ok and res["peer_timeout_reported"]
    dist.barrier()
    st.comm_shutdown()

    res["ok"] = bool(ok)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(res), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

"""Host-side logic of the multi-GPU path on CPU (gloo, world_size 2): index-range sharding, the handle exchange, and the
arithmetic contract of the exchange step - per-shard gradients summed over ranks (dW, loss sum, record count in ONE
buffer) followed by a replicated optimizer step equal the single-process result. The per-shard gradients come from the
CPU oracle here (the CUDA path is checked on GPUs by tests/test_multi_gpu.py)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vknrc_b200.dist import shard_range  # noqa: E402


def test_shard_range_partitions_exactly():
    for n in (0, 1, 127, 128, 129, 16384, 2073600, 65536 + 7):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 128):
                spans = [shard_range(n, r, world, align) for r in range(world)]
                assert spans[0][0] == 0 and spans[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
                assert all(0 <= lo <= hi <= n for lo, hi in spans)
                assert all((hi - lo) % align == 0 for lo, hi in spans[:-1] if hi < n)
                assert max(hi - lo for lo, hi in spans) <= -(-(-(-n // world)) // align) * align
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import oracle
    from vknrc_b200.dist import allreduce_gradient_reference, exchange_handles, shard_range
    # 1. handle exchange: rank order, exact bytes
    mine = bytes([(rank * 7 + i) % 256 for i in range(64)])
    got = exchange_handles(mine)
    ok_handles = got == [bytes([(r * 7 + i) % 256 for i in range(64)]) for r in range(world)]
    # 2. sharded gradient + all-reduce + replicated Adam == single process
    rng = np.random.default_rng(3)
    n = 600
    w32 = (rng.standard_normal(20672) * np.sqrt(2 / 64)).astype(np.float32)
    rec = np.concatenate([rng.uniform(-2, 2, (n, 3)), rng.uniform(0, 1, (n, 11))], axis=1).astype(np.float32)
    tgt = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    enc = oracle.encode(rec)
    lo, hi = shard_range(n, rank, world, align=128)
    buf = np.zeros(20736, np.float32)
    if hi > lo:
        buf[:20672] = oracle.gradient(w32.astype(np.float16), enc[lo:hi], tgt[lo:hi], oracle.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle.ACC_FP32)
    buf[20673] = hi - lo
    t = torch.from_numpy(buf)
    allreduce_gradient_reference(t)
    full = oracle.gradient(w32.astype(np.float16), enc, tgt, oracle.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle.ACC_FP32)
    g = t.numpy()
    ok_count = g[20673] == n
    err = float(np.abs(g[:20672] - full).max() / np.abs(full).max())
    opt = oracle.Optimizer(w32)
    opt.step(g[:20672], int(g[20673]), True, False)
    # replicated step: every rank must hold identical weights
    wt = torch.from_numpy(opt.weights.astype(np.int32))
    outs = [torch.empty_like(wt) for _ in range(world)]
    dist.all_gather(outs, wt)
    ok_replicated = all(torch.equal(outs[0], o) for o in outs)
    q.put((rank, ok_handles, bool(ok_count), err, ok_replicated))
    dist.destroy_process_group()


def test_gloo_world2_sharded_gradient_allreduce_matches_single_process():
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_handles, ok_count, err, ok_replicated in res:
        assert ok_handles and ok_count and ok_replicated, (rank, ok_handles, ok_count, ok_replicated)
        assert err <= 1e-5, err  # fp32 sums of two shard gradients vs one pass over the batch

#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native NRC MLP (BASELINE.json metric: NRC MLP inference queries/s and
training records/s, with tensor-pipe roofline fraction), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[2] -- 1920x1080 = 2 073 600 synthetic pre-encoded path-vertex queries
([n][64] fp16, the layout of test/evaluate_NV.comp) through the 64-wide fp16 MLP, random He-normal weights; one "step"
= one pass over that batch. `value` = queries/s with inputs resident in HBM (CUDA events around exactly K launches,
max over ranks); `e2e` = the same call fed from pinned HOST buffers with the H2D / D2H copies inside the timed region.
`extra` carries the other half of the metric (training records/s on configs[3], 4 x 16384 records per frame) and the
fused-encode inference path. Multi-GPU: queries (and training records) are sharded by index range, one process per
GPU, weak scaling; training all-reduces the 82 944-byte gradient buffer INSIDE the training kernel (peer-mapped inboxes
over NVLink) before a replicated Adam step; the NCCL version of the same frame is timed beside it.

--impl reference times the reference's OWN CPU implementation of the same path (test/main.cpp `Evaluate`, compiled
unmodified into oracle/_ref) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_QUERIES = 1920 * 1080
FLOP_PER_QUERY = 41344           # 2 * (5*64*64 + 3*64), SURVEY 8d (padding not counted)
FLOP_PER_TRAIN_RECORD = 115840   # fwd 41344 + dA 33152 + dW 41344
BYTES_PER_QUERY = 128 + 6        # pre-encoded input + fp16x3 output
WORKLOAD = "nrc_inference_1080p_preencoded"
METRIC, UNIT = "nrc_mlp_inference_queries_per_s", "queries/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": p["bf16_tflops"], "tflops_sustained": p.get("bf16_tflops_sustained"), "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"tflops": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}  # B200_PROFILING.md fallback


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(index: int):
    """Best effort: run this rank (and therefore first-touch its pinned host buffers) on the CPUs of the NUMA node its GPU
    hangs off, so that N ranks do not all stream their e2e inputs out of one socket's memory. Returns the node or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def cpu_reference_rate(sample: int, repeats: int, threads: int):
    """queries/s of the reference's CPU `Evaluate` (oracle/_ref) on `sample` queries, best-of-median over repeats."""
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    import oracle
    rng = np.random.default_rng(0)
    w16 = (rng.standard_normal(20672) * np.sqrt(2 / 64)).astype(np.float32).astype(np.float16)
    x = rng.uniform(0, 1, (sample, 64)).astype(np.float16)
    kind = "reference" if oracle.ref_available() else "port"
    fn = (lambda: oracle.ref_evaluate(w16, x)) if kind == "reference" else (lambda: oracle.evaluate(w16, x, oracle.ACC_FP16_CHUNK16))
    fn()
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return sample / float(np.median(ts)), kind, ts


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # exactly K timed steps after W warm-up steps, as for the GPU arm; each step is a bounded sample of the frame's queries,
    # sized so that the whole run stays within about two minutes (~0.3 s per 65 536 queries on 16 host cores)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    sample = 65536 if steps + warm <= 320 else max(4096, (65536 * 320 // (steps + warm)) // 128 * 128)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    import oracle
    rng = np.random.default_rng(0)
    w16 = (rng.standard_normal(20672) * np.sqrt(2 / 64)).astype(np.float32).astype(np.float16)
    x = rng.uniform(0, 1, (sample, 64)).astype(np.float16)
    kind = "reference" if oracle.ref_available() else "port"
    fn = (lambda: oracle.ref_evaluate(w16, x)) if kind == "reference" else (lambda: oracle.evaluate(w16, x, oracle.ACC_FP16_CHUNK16))
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    value = sample * steps / dt
    desc = f"{sample} of {N_QUERIES} queries per step ({'test/main.cpp Evaluate, Eigen fp16, -O3 -mavx2 -mf16c -mfma -fopenmp' if kind == 'reference' else 'oracle port'})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic", "config": {"workload": WORKLOAD, "queries_per_step_sampled": sample, "queries_per_frame": N_QUERIES},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import vknrc_b200 as nrc

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback; use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None  # pinned e2e buffers then live next to this rank's GPU
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    K, W = args.steps, max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    st = nrc.NrcState(local, (1920, 1080), seed=1234)  # replicated weights: same seed on every rank
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    n = N_QUERIES  # per GPU (weak scaling: every rank owns its own index range of a world*n query buffer)
    x = torch.rand((n, 64), device=dev, generator=g).half()
    out = torch.empty((n, 3), device=dev, dtype=torch.float16)
    infer = lambda: st.infer_encoded(x, out, clamp=True)

    with ClockSampler(local) as clocks:
        ms = timed(infer, K, W)
        value = world * n / (ms * 1e-3)

        # ---- e2e: same call, HOST buffers, copies inside the timed region
        hx = torch.empty((n, 64), dtype=torch.float16).pin_memory()
        hx.copy_(x.cpu())
        hout = torch.empty((n, 3), dtype=torch.float16).pin_memory()

        def e2e_step():  # the C-ABI call on host buffers: chunked H2D / MLP / D2H overlapped inside nrc_infer_encoded_host
            st.infer_encoded_host(hx, hout, clamp=True)
        e2e_steps = max(3, min(K, 20))
        ms_e2e = timed(e2e_step, e2e_steps, 3)

        extra = {}
        if not args.no_extra:
            # fused-encode inference from 56-byte UnpackedNRCInput-shaped records
            rec = torch.rand((n, 14), device=dev, generator=g)
            ms_u = timed(lambda: st.infer_unpacked(rec, outputs=out), max(10, K // 4), 3)
            extra["infer_unpacked_queries_per_s"] = world * n / (ms_u * 1e-3)
            extra["infer_unpacked_ms_per_step"] = ms_u
            # the reference's own formats: 20-byte NRCEvalRecord per pixel + scene gather (UnpackNRCInput) -> composite into the
            # rgba32f / rg32f screen images: the exact nrc_inference.comp pass (nrc_infer), from device and from host records
            from vknrc_b200 import synth
            sa = synth.make_scene_arrays(7, n_prims=20000, n_instances=8, n_materials=64, n_textures=8)
            scene = nrc.DeviceScene(sa["vertices"], sa["vertex_indices"], sa["texcoords"], sa["texcoord_indices"], sa["materials"],
                                    sa["material_ids"], sa["transforms"], sa["textures"], device=local)
            ev = synth.eval_records_screen(11 + rank, 1920, 1080, 20000, 8)
            h_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).pin_memory()
            d_ev = h_ev.to(dev)
            d_bf = torch.rand((1080, 1920, 4), device=dev, generator=g)
            d_gb = torch.rand((1080, 1920, 2), device=dev, generator=g)
            d_trs = [torch.zeros(nrc.TRAIN_BATCH_SIZE * 40, dtype=torch.uint8, device=dev) for _ in range(4)]
            cnt = torch.tensor([n], dtype=torch.int32, device=dev)
            ms_r = timed(lambda: st.infer(d_ev, cnt, scene, d_bf, d_gb, 1920, d_trs, max_count=n), max(10, K // 4), 3)
            extra["infer_eval_records_scatter_queries_per_s"] = world * n / (ms_r * 1e-3)
            extra["infer_eval_records_scatter_ms_per_step"] = ms_r
            h_bf = torch.empty((1080, 1920, 4), dtype=torch.float32).pin_memory()

            def e2e_records():  # host records in, composited image back: 41.5 MB up, 33.2 MB down per frame
                d_ev.copy_(h_ev, non_blocking=True)
                st.infer(d_ev, cnt, scene, d_bf, d_gb, 1920, d_trs, max_count=n)
                h_bf.copy_(d_bf, non_blocking=True)
            ms_re = timed(e2e_records, max(3, min(K, 20)), 3)
            extra["e2e_eval_records_queries_per_s"] = world * n / (ms_re * 1e-3)
            extra["e2e_eval_records_ms_per_step"] = ms_re
            # training: one frame = 4 dependent batches of 16384 records (configs[3]); records sharded per GPU
            nb = nrc.TRAIN_BATCH_SIZE
            trec = torch.rand((4, nb, 14), device=dev, generator=g)
            ttgt = torch.rand((4, nb, 3), device=dev, generator=g)

            trecs, ttgts = [trec[b] for b in range(4)], [ttgt[b] for b in range(4)]
            if world > 1:
                st.comm_connect()  # peer-mapped inboxes: the gradient all-reduce runs inside the training kernel

            def train_frame():
                # the whole frame (4 x [gradient -> reduce -> (NVLink all-reduce) -> Adam]) is ONE cooperative kernel launch
                st.train_frame_unpacked(trecs, ttgts)
            ms_t = timed(train_frame, max(10, K // 4), 3)
            # the same frame on 40-byte NRCTrainRecord buffers (scene gather fused), the reference's NNTrain input
            d_trec = [torch.from_numpy(synth.train_records(100 + 10 * rank + b, nb, 20000, 8).view(np.uint8).reshape(-1)).to(dev) for b in range(4)]
            ms_tr = timed(lambda: st.train_frame(d_trec, scene, max_count=nb), max(10, K // 4), 3)
            extra["train_records_frame_ms_4x16384"] = ms_tr
            extra["train_records_per_s"] = world * 4 * nb / (ms_t * 1e-3)
            extra["train_ms_per_frame_4x16384"] = ms_t
            extra["train_tflops"] = world * 4 * nb * FLOP_PER_TRAIN_RECORD / (ms_t * 1e-3) / 1e12
            # throughput point of the sweep: 2^20 records in one step
            big = 1 << 20
            brec = torch.rand((big, 14), device=dev, generator=g)
            btgt = torch.rand((big, 3), device=dev, generator=g)

            def train_big():
                st.train_batch_unpacked(brec, btgt, write_use_weights=True)
            ms_b = timed(train_big, 5, 2)
            if world > 1:  # the stock-collective version of the same frame, for comparison: gradient -> NCCL all-reduce -> Adam
                st_split = nrc.NrcState(local, (1920, 1080), seed=1234)
                gt = st_split.gradient_tensor()

                def train_frame_nccl():
                    for b in range(4):
                        st_split.gradient_unpacked(trec[b], ttgt[b])
                        dist.all_reduce(gt)  # 20 736 fp32: dW + loss + record count
                        st_split.adam_step(write_use_weights=(b == 3))
                extra["train_ms_per_frame_nccl_split"] = timed(train_frame_nccl, max(10, K // 4), 3)
            extra["train_2p20_records_per_s"] = world * big / (ms_b * 1e-3)
            extra["train_2p20_tflops_per_gpu"] = big * FLOP_PER_TRAIN_RECORD / (ms_b * 1e-3) / 1e12

    if not args.no_extra:
        # The headline number above is a short burst (K launches). On random data this kernel is limited by the 1 kW board
        # power cap once the power integrator catches up (profiles/r01_power_cap_probe.txt): ~1 s of back-to-back launches
        # shows the sustained regime, with its own clock / throttle-reason record.
        with ClockSampler(local) as clocks_sustained:
            n_sus = max(200, int(0.9 / (ms * 1e-3)))
            ms_sus = timed(infer, n_sus, 3)
        extra["infer_sustained_ms_per_step"] = ms_sus
        extra["infer_sustained_launches"] = n_sus
        extra["infer_sustained_tflops"] = n * FLOP_PER_QUERY / (ms_sus * 1e-3) / 1e12
        extra["infer_sustained_clocks"] = clocks_sustained.summary()

    peaks = measured_peaks()
    tflops = n * FLOP_PER_QUERY / (ms * 1e-3) / 1e12  # per GPU (the kernel of one rank)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("nrc_infer_kernel_dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_per_gpu_per_step": n, "network": "64->5x(64,ReLU)->3, fp16 weights, fp32 TMEM accumulate",
                   "l2_policy": "inputs (265 MB per step) exceed the 126 MB L2; no flush needed", "parallelism": f"index-range x{world}",
                   "host_numa_node_rank0": numa},
        "e2e": {"value": world * n / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * 128, "d2h_bytes_per_step": n * 6,
                "ms_per_step": ms_e2e},
        "gpu_launches": K,  # one nrc_infer_kernel launch per step inside the timed region
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "achieved": tflops, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tflops / peaks["tflops"],
                     "traffic": traffic, "peak_source": f"{peaks['source']} cuBLAS bf16 burst (sustained {peaks['tflops_sustained']})",
                     "flop_per_query": FLOP_PER_QUERY,
                     "hbm_gbs_achieved": n * BYTES_PER_QUERY / (ms * 1e-3) / 1e9, "hbm_gbs_peak": peaks["hbm_gbs"]},
        "extra": extra,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate16k, kind, _ = cpu_reference_rate(16384, 1, threads)
        sample = int(min(1 << 20, max(16384, rate16k * 12)) // 128 * 128)  # ~12 s per repeat
        rate, kind, ts = cpu_reference_rate(sample, 2, threads)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": f"{sample} of {N_QUERIES} pre-encoded queries, median of 2 runs, "
                                          f"{'test/main.cpp Evaluate (Eigen fp16) from oracle/_ref' if kind == 'reference' else 'oracle port'}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""2-GPU probe: is NVLink multicast (NVLS) usable from this container? torch symmetric memory rendezvous + raw driver attribute.
torchrun --nproc-per-node 2 tools/probe_multicast.py"""
import os, sys, traceback
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
def say(*a):
    if rank == 0:
        print("[probe]", *a, flush=True)
try:
    try:
        from cuda.bindings import driver as cu
    except Exception:
        from cuda import cuda as cu
    cu.cuInit(0)
    err, dev = cu.cuDeviceGet(local)
    for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED"):
        a = getattr(cu.CUdevice_attribute, name, None)
        if a is not None:
            say(name, cu.cuDeviceGetAttribute(a, dev))
except Exception:
    say("cuda-python probe failed:", traceback.format_exc())
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f"cuda:{local}")
    h = symm_mem.rendezvous(t, dist.group.WORLD)
    say("symm_mem ok: multicast_ptr", hex(h.multicast_ptr), "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs])
    say("has_multicast_support", symm_mem.has_multicast_support("cuda", local) if hasattr(symm_mem, "has_multicast_support") else "n/a")
    t.fill_(rank + 1.0)
    dist.barrier(); torch.cuda.synchronize()
    if h.multicast_ptr:
        out = torch.ops.symm_mem.multimem_all_reduce_(t[:1024], "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        say("multimem_all_reduce_ result[0] =", float(t[0]), "(expect", world * (world + 1) / 2, ")")
except Exception:
    say("symm_mem probe failed:", traceback.format_exc())
dist.barrier()
dist.destroy_process_group()

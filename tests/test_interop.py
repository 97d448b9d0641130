"""Vulkan interop entry points (include/nrc_b200.h "Vulkan interop").

There is no Vulkan loader / ICD in this environment, so a VkDeviceMemory cannot be exported here. What is tested:
argument and error paths (CPU and GPU: a bad fd is reported, never thrown, and leaves the library usable), and - when the
driver accepts it - a round trip through a POSIX fd exported from a CUDA virtual-memory allocation, which goes through
the same cudaImportExternalMemory / cudaExternalMemoryGetMappedBuffer calls a Vulkan fd would."""
import ctypes as C
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L():
    import vknrc_b200
    return vknrc_b200.lib()


def test_interop_argument_errors(L):
    out = C.c_void_p()
    assert L.nrc_import_vulkan_memory_fd(0, -1, 4096, 0, C.byref(out)) == -1 and not out.value
    assert L.nrc_import_vulkan_memory_fd(0, 3, 0, 0, C.byref(out)) == -1
    assert L.nrc_import_vulkan_memory_fd(0, 3, 4096, 0, None) == -1
    assert b"nrc_import_vulkan_memory_fd" in L.nrc_last_error()
    assert L.nrc_import_vulkan_timeline_semaphore_fd(0, -1, C.byref(out)) == -1 and not out.value
    ptr = C.c_void_p()
    assert L.nrc_external_memory_map_buffer(None, 0, 16, C.byref(ptr)) == -1
    assert L.nrc_external_semaphore_wait(None, 1, None) == -1
    assert L.nrc_external_semaphore_signal(None, 1, None) == -1
    # releasing nothing is a no-op, like free(NULL)
    assert L.nrc_external_memory_release(None) == 0
    assert L.nrc_external_semaphore_release(None) == 0


@pytest.mark.gpu
def test_bad_fd_is_reported_and_library_stays_usable(L):
    import torch
    from vknrc_b200 import api
    out = C.c_void_p()
    # (a fresh fd per call: the CUDA driver may close an fd it was handed even when the import fails)
    fd = os.open("/dev/null", os.O_RDONLY)
    assert L.nrc_import_vulkan_memory_fd(0, fd, 1 << 20, 0, C.byref(out)) in (-2, -4) and not out.value
    assert len(L.nrc_last_error()) > 0
    fd = os.open("/dev/null", os.O_RDONLY)
    assert L.nrc_import_vulkan_timeline_semaphore_fd(0, fd, C.byref(out)) == -2 and not out.value
    # the failed imports must not leave a sticky CUDA error behind
    st = api.NrcState(0, (64, 64), 7)
    x = torch.rand(256, 64, device="cuda").half()
    y = st.infer_encoded(x)
    torch.cuda.synchronize()
    assert y.shape == (256, 3) and torch.isfinite(y.float()).all()


@pytest.mark.gpu
def test_external_memory_round_trip_through_a_cuda_exported_fd(L):
    """Records living in imported external memory feed the kernels exactly like cudaMalloc'ed ones."""
    import torch
    from cuda.bindings import driver as cu
    from vknrc_b200 import api
    torch.cuda.init()
    torch.zeros(1, device="cuda")

    def ok(res):
        err, *rest = res
        if int(err) != 0:
            pytest.skip(f"CUDA VMM export not available here: {err}")
        return rest[0] if len(rest) == 1 else rest

    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = ok(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    n = 4096
    size = ((n * 128 + gran - 1) // gran) * gran
    handle = ok(cu.cuMemCreate(size, prop, 0))
    try:
        fd = int(ok(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)))
        mem = C.c_void_p()
        rc = L.nrc_import_vulkan_memory_fd(0, fd, size, 0, C.byref(mem))
        if rc != 0:
            os.close(fd)
            pytest.skip("this driver does not accept a CUDA-exported fd as an opaque external-memory fd: "
                        + L.nrc_last_error().decode())
        ptr = C.c_void_p()
        assert L.nrc_external_memory_map_buffer(mem, 0, size + 1, C.byref(ptr)) == -1  # range check
        assert L.nrc_external_memory_map_buffer(mem, 0, n * 128, C.byref(ptr)) == 0 and ptr.value
        st = api.NrcState(0, (64, 64), 11)
        x = torch.rand(n, 64, device="cuda").half()
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        assert rt.cudaMemcpy(ptr, x.data_ptr(), n * 128, 3) == 0  # device -> device
        y_ref = st.infer_encoded(x)
        y = torch.empty(n, 3, device="cuda", dtype=torch.float16)
        api._check(L.nrc_infer_encoded(st._h, ptr, y.data_ptr(), n, 0, None))
        torch.cuda.synchronize()
        assert torch.equal(y, y_ref)
        rt.cudaFree.argtypes = [C.c_void_p]
        assert rt.cudaFree(ptr) == 0
        assert L.nrc_external_memory_release(mem) == 0
    finally:
        cu.cuMemRelease(handle)

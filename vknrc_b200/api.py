"""ctypes binding of ``libnrc_b200.so`` (the C ABI in ``include/nrc_b200.h``) for tests, bench and examples.

This is plumbing, not the product: the product is the CUDA/C++ library. PyTorch is used only to own device memory,
streams and (for multi-GPU) ``torch.distributed``. There is NO fallback: if the library is missing or the device is
not an sm_100 part, every call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (NRC_B200_LIB: development override used by tools/lab_train.py to time a variant build of the same library)
LIB_PATH = os.environ.get("NRC_B200_LIB") or os.path.join(_HERE, "libnrc_b200.so")

WEIGHT_COUNT = 20672
GRADIENT_FLOATS = 20736
GRAD_LOSS_SLOT = 20672
GRAD_COUNT_SLOT = 20673
TRAIN_BATCH_SIZE = 16384
TRAIN_BATCH_COUNT = 4

OPT_ENTRY_DTYPE = np.dtype([("m", "<f4"), ("v", "<f4"), ("weight", "<f4"), ("ema_weight", "<f4")])
OPT_STATE_DTYPE = np.dtype([("t", "<u4"), ("beta1_t", "<f4"), ("beta2_t", "<f4"), ("alpha_t", "<f4"), ("alpha_t_1", "<f4")])
# shader/src/NRCRecord.glsl:6-38
PACKED_INPUT_DTYPE = np.dtype([("primitive_id", "<u4"), ("flip_bit_instance_id", "<u4"), ("barycentric_2x16U", "<u4"),
                               ("scattered_dir_2x16U", "<u4")])
EVAL_RECORD_DTYPE = np.dtype([("dst", "<u4"), ("packed_input", PACKED_INPUT_DTYPE)])
TRAIN_RECORD_DTYPE = np.dtype([("bias", "<f4", 3), ("factor", "<f4", 3), ("packed_input", PACKED_INPUT_DTYPE)])
assert EVAL_RECORD_DTYPE.itemsize == 20 and TRAIN_RECORD_DTYPE.itemsize == 40


MATERIAL_DTYPE = np.dtype([("diffuse", "<f4", 3), ("diffuse_texture_id", "<u4"), ("specular", "<f4", 3), ("specular_texture_id", "<u4"),
                           ("emission", "<f4", 3), ("emission_texture_id", "<u4"), ("metallic", "<f4"), ("roughness", "<f4"),
                           ("ior", "<f4"), ("_pad", "<u4")])  # shader/src/Scene.glsl:12-20, 64 B std430 stride
assert MATERIAL_DTYPE.itemsize == 64


class nrc_texture_t(C.Structure):  # include/nrc_b200_types.h NrcTexture
    _fields_ = [("texels_rgba8_srgb", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class nrc_scene_t(C.Structure):  # include/nrc_b200_types.h NrcScene (device pointers)
    _fields_ = [("vertices", C.c_void_p), ("vertex_indices", C.c_void_p), ("texcoords", C.c_void_p), ("texcoord_indices", C.c_void_p),
                ("materials", C.c_void_p), ("material_ids", C.c_void_p), ("transforms", C.c_void_p), ("textures", C.c_void_p),
                ("texture_count", C.c_uint32), ("prim_table", C.c_void_p)]


class DeviceScene:
    """Uploads scene buffers given in the reference's layouts (numpy arrays; see shader/src/Scene.glsl:8-71) and holds
    the NrcScene struct of device pointers that the record-format entry points take."""

    def __init__(self, vertices, vertex_indices, texcoords, texcoord_indices, materials, material_ids, transforms, textures, device=0,
                 prim_table: bool = True):
        """prim_table: also build the per-primitive 64-byte rows (nrc_scene_build_prim_table) that turn the two-level
        index -> attribute gather of UnpackNRCInput into one aligned line (results are bit-identical)."""
        import torch
        dev = f"cuda:{device}"

        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dt).view(np.uint8).reshape(-1)).to(dev)
        self._keep = {"vertices": up(vertices, np.float32), "vertex_indices": up(vertex_indices, np.uint32), "texcoords": up(texcoords, np.float32),
                      "texcoord_indices": up(texcoord_indices, np.uint32), "materials": up(materials, MATERIAL_DTYPE),
                      "material_ids": up(material_ids, np.uint32), "transforms": up(transforms, np.float32)}
        self._tex = [up(t, np.uint8) for t in textures]
        table = (nrc_texture_t * max(1, len(textures)))()
        for i, t in enumerate(textures):
            table[i] = nrc_texture_t(self._tex[i].data_ptr(), t.shape[1], t.shape[0])
        self._table = torch.from_numpy(np.frombuffer(bytes(table), np.uint8).copy()).to(dev)
        k = self._keep
        self.c = nrc_scene_t(k["vertices"].data_ptr(), k["vertex_indices"].data_ptr(), k["texcoords"].data_ptr(), k["texcoord_indices"].data_ptr(),
                             k["materials"].data_ptr(), k["material_ids"].data_ptr(), k["transforms"].data_ptr(), self._table.data_ptr(),
                             len(textures), None)
        if prim_table:
            n_prims = int(np.asarray(material_ids).shape[0])
            self._prims = torch.empty(max(64, lib().nrc_scene_prim_table_bytes(n_prims)), dtype=torch.uint8, device=dev)
            assert self._prims.data_ptr() % 64 == 0
            with torch.cuda.device(device):
                _check(lib().nrc_scene_build_prim_table(C.byref(self.c), n_prims, self._prims.data_ptr(), _stream()))
            self.c.prim_table = self._prims.data_ptr()


class NrcError(RuntimeError):
    pass


class nrc_config_t(C.Structure):
    _fields_ = [("extent_width", C.c_uint32), ("extent_height", C.c_uint32), ("seed", C.c_uint64)]


_lib = None

# name -> (restype, argtypes); also the list tests check against the symbols include/nrc_b200.h declares
SIGNATURES = {
    "nrc_last_error": (C.c_char_p, []),
    "nrc_get_eval_record_buffer_size": (C.c_uint64, [C.c_uint32, C.c_uint32]),
    "nrc_get_batch_train_record_buffer_size": (C.c_uint64, []),
    "nrc_get_train_batch_count": (C.c_uint32, []),
    "nrc_get_train_batch_size": (C.c_uint32, []),
    "nrc_get_weight_count": (C.c_uint32, []),
    "nrc_get_default_train_probability": (C.c_float, []),
    "nrc_create": (C.c_int, [C.POINTER(nrc_config_t), C.c_int, C.POINTER(C.c_void_p)]),
    "nrc_destroy": (None, [C.c_void_p]),
    "nrc_reset_mlp_buffers": (C.c_int, [C.c_void_p, C.c_uint64]),
    "nrc_set_weights": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nrc_get_weight_buffer": (C.c_void_p, [C.c_void_p]),
    "nrc_get_use_weight_buffer": (C.c_void_p, [C.c_void_p]),
    "nrc_get_optimizer_entry_buffer": (C.c_void_p, [C.c_void_p]),
    "nrc_get_optimizer_state_buffer": (C.c_void_p, [C.c_void_p]),
    "nrc_get_gradient_buffer": (C.c_void_p, [C.c_void_p]),
    "nrc_download": (C.c_int, [C.c_void_p] * 7),
    "nrc_set_use_ema_weights": (None, [C.c_void_p, C.c_int]),
    "nrc_is_use_ema_weights": (C.c_int, [C.c_void_p]),
    "nrc_set_train_probability": (None, [C.c_void_p, C.c_float]),
    "nrc_get_train_probability": (C.c_float, [C.c_void_p]),
    "nrc_next_frame": (C.c_uint32, [C.c_void_p]),
    "nrc_get_seed": (C.c_uint32, [C.c_void_p]),
    "nrc_mlp_evaluate_encoded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "nrc_mlp_gradient_encoded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "nrc_infer_encoded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "nrc_infer_unpacked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "nrc_infer_scatter_unpacked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                             C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_void_p]),
    "nrc_gradient_unpacked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                        C.c_void_p]),
    "nrc_gradient_encoded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]),
    "nrc_adam_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "nrc_train_batch_unpacked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                           C.c_int, C.c_void_p]),
    "nrc_train_frame_unpacked": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32,
                                           C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p]),
    "nrc_infer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                            C.POINTER(C.c_void_p), C.c_void_p]),
    "nrc_infer_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrc_unpack_inputs": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrc_encode_inputs": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "nrc_encode_packed_inputs": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrc_gradient": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "nrc_train_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]),
    "nrc_train_frame": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p, C.c_void_p]),
    "nrc_comm_handle_bytes": (C.c_uint32, []),
    "nrc_comm_init": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "nrc_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nrc_comm_shutdown": (C.c_int, [C.c_void_p]),
    "nrc_comm_world": (C.c_uint32, [C.c_void_p]),
    "nrc_comm_buffer_bytes": (C.c_uint64, []),
    "nrc_comm_attach": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_void_p]),
    "nrc_comm_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nrc_comm_set_timeout": (None, [C.c_void_p, C.c_uint32]),
    "nrc_frame_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "nrc_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "nrc_set_prediction_capture": (None, [C.c_void_p, C.c_void_p]),
    "nrc_image_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_float, C.c_void_p]),
    "nrc_infer_encoded_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "nrc_infer_eval_records_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrc_scene_prim_table_bytes": (C.c_uint64, [C.c_uint32]),
    "nrc_scene_build_prim_table": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "nrc_image_infer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "nrc_import_vulkan_memory_fd": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "nrc_external_memory_map_buffer": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]),
    "nrc_external_memory_release": (C.c_int, [C.c_void_p]),
    "nrc_import_vulkan_timeline_semaphore_fd": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "nrc_external_semaphore_wait": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "nrc_external_semaphore_signal": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "nrc_external_semaphore_release": (C.c_int, [C.c_void_p]),
}


def lib() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NrcError(f"{LIB_PATH} is missing: run `python -m vknrc_b200.build` (needs nvcc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError:
                if os.environ.get("NRC_B200_LIB"):  # an older development build timed beside the current one (tools/lab_train.py)
                    continue
                raise
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise NrcError(f"nrc_b200 error {rc}: {lib().nrc_last_error().decode()}")


def _ptr(t) -> int:
    """Device pointer of a torch CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return t.data_ptr()


def _stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


class _DevView:
    """Zero-copy torch view of library-owned device memory via __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def mlp_evaluate_encoded(weights, inputs, outputs=None):
    """launchKernel("evaluate_32.spv", ..., weights, inputs, outputs) (test/main.cpp:116-117)."""
    import torch
    n = inputs.shape[0]
    if outputs is None:
        outputs = torch.empty((n, 3), dtype=torch.float16, device=inputs.device)
    _check(lib().nrc_mlp_evaluate_encoded(_ptr(weights), _ptr(inputs), _ptr(outputs), n, _stream()))
    return outputs


def mlp_gradient_encoded(weights, dw, inputs, targets):
    """launchKernel("train_32.spv", ..., weights, dw, inputs, targets) (test/main.cpp:180-181); dw accumulates."""
    _check(lib().nrc_mlp_gradient_encoded(_ptr(weights), _ptr(dw), _ptr(inputs), _ptr(targets), inputs.shape[0], _stream()))
    return dw


def unpack_inputs(packed_inputs, scene: "DeviceScene", stride_bytes: int = 16, n=None):
    """UnpackNRCInput (NRCRecord.glsl:98-125) as a kernel of its own: -> [n,14] fp32."""
    import torch
    n = packed_inputs.numel() * packed_inputs.element_size() // stride_bytes if n is None else n
    out = torch.empty((n, 14), dtype=torch.float32, device=packed_inputs.device)
    _check(lib().nrc_unpack_inputs(_ptr(packed_inputs), stride_bytes, n, C.byref(scene.c), _ptr(out), _stream()))
    return out


def encode_inputs(records, stride_bytes: int = 56, n=None, out=None):
    """NRCInputEncode (NRCRecord.glsl:77-95) as a kernel of its own: [n,14] fp32 records -> [n,64] fp16 features."""
    import torch
    n = records.numel() * records.element_size() // stride_bytes if n is None else n
    out = torch.empty((n, 64), dtype=torch.float16, device=records.device) if out is None else out
    _check(lib().nrc_encode_inputs(_ptr(records), stride_bytes, n, _ptr(out), _stream()))
    return out


def encode_packed_inputs(packed_inputs, scene: "DeviceScene", stride_bytes: int = 16, n=None, out=None):
    """UnpackNRCInput + NRCInputEncode (NRCRecord.glsl:98-125, 77-95): PackedNRCInput words -> [n,64] fp16 features."""
    import torch
    n = packed_inputs.numel() * packed_inputs.element_size() // stride_bytes if n is None else n
    out = torch.empty((n, 64), dtype=torch.float16, device=packed_inputs.device) if out is None else out
    _check(lib().nrc_encode_packed_inputs(_ptr(packed_inputs), stride_bytes, n, C.byref(scene.c), _ptr(out), _stream()))
    return out


class NrcState:
    """Python handle on the C++ ``nrc::NrcState`` (shaped like VkNRCState, src/VkNRCState.hpp:17-90)."""

    def __init__(self, device: int = 0, extent=(1920, 1080), seed: int = 0):
        self._h = C.c_void_p()
        cfg = nrc_config_t(extent[0], extent[1], seed)
        _check(lib().nrc_create(C.byref(cfg), device, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().nrc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- statics (src/VkNRCState.hpp:83-89)
    @staticmethod
    def get_eval_record_buffer_size(extent) -> int:
        return lib().nrc_get_eval_record_buffer_size(extent[0], extent[1])

    @staticmethod
    def get_batch_train_record_buffer_size() -> int:
        return lib().nrc_get_batch_train_record_buffer_size()

    @staticmethod
    def get_train_batch_count() -> int:
        return lib().nrc_get_train_batch_count()

    @staticmethod
    def get_train_batch_size() -> int:
        return lib().nrc_get_train_batch_size()

    @staticmethod
    def get_weight_count() -> int:
        return lib().nrc_get_weight_count()

    # ---- state
    def reset_mlp_buffers(self, seed: int):
        _check(lib().nrc_reset_mlp_buffers(self._h, seed))

    def set_weights(self, fp32_weights: np.ndarray):
        w = np.ascontiguousarray(fp32_weights, np.float32).reshape(WEIGHT_COUNT)
        _check(lib().nrc_set_weights(self._h, w.ctypes.data))

    def set_use_ema_weights(self, v: bool):
        lib().nrc_set_use_ema_weights(self._h, int(v))

    def is_use_ema_weights(self) -> bool:
        return bool(lib().nrc_is_use_ema_weights(self._h))

    def set_train_probability(self, p: float):
        lib().nrc_set_train_probability(self._h, p)

    def get_train_probability(self) -> float:
        return lib().nrc_get_train_probability(self._h)

    def next_frame(self) -> int:
        return lib().nrc_next_frame(self._h)

    def download(self):
        """Synchronises the current stream and returns host copies of every persistent buffer."""
        w = np.empty(WEIGHT_COUNT, np.uint16)
        uw = np.empty(WEIGHT_COUNT, np.uint16)
        e = np.empty(WEIGHT_COUNT, OPT_ENTRY_DTYPE)
        s = np.empty(1, OPT_STATE_DTYPE)
        g = np.empty(GRADIENT_FLOATS, np.float32)
        _check(lib().nrc_download(self._h, w.ctypes.data, uw.ctypes.data, e.ctypes.data, s.ctypes.data, g.ctypes.data, _stream()))
        return {"weights": w.view(np.float16), "use_weights": uw.view(np.float16), "optimizer_entries": e,
                "optimizer_state": s[0], "gradients": g}

    def gradient_tensor(self):
        """Zero-copy fp32[20736] torch view of the gradient buffer (what a multi-GPU caller all-reduces)."""
        import torch
        return torch.as_tensor(_DevView(lib().nrc_get_gradient_buffer(self._h), (GRADIENT_FLOATS,), "<f4"), device=f"cuda:{self.device}")

    def weight_tensor(self, use: bool = False):
        import torch
        p = lib().nrc_get_use_weight_buffer(self._h) if use else lib().nrc_get_weight_buffer(self._h)
        return torch.as_tensor(_DevView(p, (WEIGHT_COUNT,), "<f2"), device=f"cuda:{self.device}")

    def set_prediction_capture(self, t):
        lib().nrc_set_prediction_capture(self._h, _ptr(t))

    # ---- inference
    def infer_encoded(self, inputs, outputs=None, clamp: bool = False):
        import torch
        n = inputs.shape[0]
        if outputs is None:
            outputs = torch.empty((n, 3), dtype=torch.float16, device=inputs.device)
        _check(lib().nrc_infer_encoded(self._h, _ptr(inputs), _ptr(outputs), n, int(clamp), _stream()))
        return outputs

    def infer_encoded_host(self, h_inputs, h_outputs, clamp: bool = False):
        """nrc_infer_encoded_host: pre-encoded queries in (pinned) HOST memory -> outputs in host memory, copies pipelined with
        the MLP inside the call (the reference harness' cudaMemcpy / launchKernel / cudaMemcpy, test/main.cpp:103-128)."""
        assert not h_inputs.is_cuda and not h_outputs.is_cuda and h_inputs.is_contiguous() and h_outputs.is_contiguous()
        n = h_inputs.shape[0]
        _check(lib().nrc_infer_encoded_host(self._h, h_inputs.data_ptr(), h_outputs.data_ptr(), n, int(clamp), _stream()))
        return h_outputs

    def infer_eval_records_host(self, h_eval_records, scene: "DeviceScene", h_outputs):
        """nrc_infer_eval_records_host: 20-byte NRCEvalRecords in (pinned) HOST memory -> fp16 x 3 radiance per query in host memory."""
        assert not h_eval_records.is_cuda and not h_outputs.is_cuda and h_eval_records.is_contiguous() and h_outputs.is_contiguous()
        n = h_eval_records.numel() * h_eval_records.element_size() // 20
        _check(lib().nrc_infer_eval_records_host(self._h, h_eval_records.data_ptr(), n, C.byref(scene.c), h_outputs.data_ptr(), _stream()))
        return h_outputs

    def infer_unpacked(self, records, count=None, outputs=None, stride_bytes: int = 56, max_count=None):
        import torch
        n = records.shape[0] if max_count is None else max_count
        if outputs is None:
            outputs = torch.empty((n, 3), dtype=torch.float16, device=records.device)
        _check(lib().nrc_infer_unpacked(self._h, _ptr(records), stride_bytes, _ptr(count), n, _ptr(outputs), _stream()))
        return outputs

    def infer_scatter_unpacked(self, dst, records, count, bias_factor_r, factor_gb, image_pitch, train_records,
                               dst_stride_bytes: int = 4, stride_bytes: int = 56, max_count=None):
        n = records.shape[0] if max_count is None else max_count
        ptrs = (C.c_void_p * 4)(*[_ptr(t) for t in train_records])
        _check(lib().nrc_infer_scatter_unpacked(self._h, _ptr(dst), dst_stride_bytes, _ptr(records), stride_bytes, _ptr(count), n,
                                                _ptr(bias_factor_r), _ptr(factor_gb), image_pitch, ptrs, _stream()))

    # ---- training
    def gradient_unpacked(self, inputs, targets, count=None, max_count=None, input_stride: int = 56, target_stride: int = 12):
        n = inputs.shape[0] if max_count is None else max_count
        _check(lib().nrc_gradient_unpacked(self._h, _ptr(inputs), input_stride, _ptr(targets), target_stride, _ptr(count), n, _stream()))

    def gradient_encoded(self, inputs, targets16, count=None, max_count=None, relative_loss: bool = False):
        n = inputs.shape[0] if max_count is None else max_count
        _check(lib().nrc_gradient_encoded(self._h, _ptr(inputs), _ptr(targets16), _ptr(count), n, int(relative_loss), _stream()))

    def adam_step(self, write_use_weights: bool = True):
        _check(lib().nrc_adam_step(self._h, int(write_use_weights), _stream()))

    def train_batch_unpacked(self, inputs, targets, count=None, max_count=None, write_use_weights=True, input_stride=56,
                             target_stride=12):
        n = inputs.shape[0] if max_count is None else max_count
        _check(lib().nrc_train_batch_unpacked(self._h, _ptr(inputs), input_stride, _ptr(targets), target_stride, _ptr(count), n,
                                              int(write_use_weights), _stream()))

    def train_frame_unpacked(self, inputs, targets, counts=None, max_count=None, input_stride=56, target_stride=12):
        """inputs / targets / counts: sequences of 4 tensors (the frame's batches); ONE kernel launch."""
        n = inputs[0].shape[0] if max_count is None else max_count
        ins = (C.c_void_p * 4)(*[_ptr(t) for t in inputs])
        tgs = (C.c_void_p * 4)(*[_ptr(t) for t in targets])
        cns = (C.c_void_p * 4)(*[_ptr(t) for t in counts]) if counts is not None else None
        _check(lib().nrc_train_frame_unpacked(self._h, ins, input_stride, tgs, target_stride, cns, n, _stream()))

    # ---- the reference's own record formats (NRCEvalRecord 20 B / NRCTrainRecord 40 B) + scene gather
    def infer(self, eval_records, count, scene: "DeviceScene", bias_factor_r, factor_gb, image_pitch, train_records, max_count=None):
        """nrc_inference.comp with the bindings of src/rg/NNInference.cpp:13-48. eval_records: uint8/uint32 tensor of 20 B records."""
        n = eval_records.numel() * eval_records.element_size() // 20 if max_count is None else max_count
        ptrs = (C.c_void_p * 4)(*[_ptr(t) for t in train_records])
        _check(lib().nrc_infer(self._h, _ptr(eval_records), _ptr(count), n, C.byref(scene.c), _ptr(bias_factor_r), _ptr(factor_gb), image_pitch,
                               ptrs, _stream()))

    def infer_packed(self, packed_inputs, scene: "DeviceScene", count=None, outputs=None, stride_bytes: int = 16, max_count=None):
        import torch
        n = packed_inputs.numel() * packed_inputs.element_size() // stride_bytes if max_count is None else max_count
        if outputs is None:
            outputs = torch.empty((n, 3), dtype=torch.float16, device=packed_inputs.device)
        _check(lib().nrc_infer_packed(self._h, _ptr(packed_inputs), stride_bytes, _ptr(count), n, C.byref(scene.c), _ptr(outputs), _stream()))
        return outputs

    def gradient(self, train_records, scene: "DeviceScene", count=None, max_count=None):
        n = train_records.numel() * train_records.element_size() // 40 if max_count is None else max_count
        _check(lib().nrc_gradient(self._h, _ptr(train_records), _ptr(count), n, C.byref(scene.c), _stream()))

    def train_batch(self, train_records, scene: "DeviceScene", count=None, max_count=None, write_use_weights=True):
        n = train_records.numel() * train_records.element_size() // 40 if max_count is None else max_count
        _check(lib().nrc_train_batch(self._h, _ptr(train_records), _ptr(count), n, C.byref(scene.c), int(write_use_weights), _stream()))

    def train_frame(self, train_records, scene: "DeviceScene", counts=None, max_count=None):
        n = train_records[0].numel() * train_records[0].element_size() // 40 if max_count is None else max_count
        recs = (C.c_void_p * 4)(*[_ptr(t) for t in train_records])
        cns = (C.c_void_p * 4)(*[_ptr(t) for t in counts]) if counts is not None else None
        _check(lib().nrc_train_frame(self._h, recs, cns, n, C.byref(scene.c), _stream()))

    # ---- one frame (src/rg/NRCRenderGraph.cpp:46-80, 100-113)
    def frame_begin(self, eval_count, train_counts):
        cns = (C.c_void_p * 4)(*[_ptr(t) for t in train_counts])
        _check(lib().nrc_frame_begin(self._h, _ptr(eval_count), cns, _stream()))

    def frame(self, eval_records, eval_count, scene: "DeviceScene", bias_factor_r, factor_gb, image_pitch, train_records, train_counts,
              max_eval_count=None):
        n = eval_records.numel() * eval_records.element_size() // 20 if max_eval_count is None else max_eval_count
        recs = (C.c_void_p * 4)(*[_ptr(t) for t in train_records])
        cns = (C.c_void_p * 4)(*[_ptr(t) for t in train_counts])
        _check(lib().nrc_frame(self._h, _ptr(eval_records), _ptr(eval_count), n, C.byref(scene.c), _ptr(bias_factor_r), _ptr(factor_gb),
                               image_pitch, recs, cns, _stream()))

    # ---- multi-GPU (one process per GPU)
    def comm_attach_symmetric(self, group=None, multicast: bool = True):
        """nrc_comm_attach on a torch symmetric-memory allocation (cuMem + NVSwitch multicast object, mapped into every rank by
        torch.distributed._symmetric_memory - plumbing): with `multicast` the in-kernel exchange pushes every word with ONE
        multimem.st through the switch instead of one store per peer. Returns True if a multicast mapping was attached."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = lib().nrc_comm_buffer_bytes()
        self._symm = symm_mem.empty(nbytes // 8, dtype=torch.int64, device=f"cuda:{self.device}")
        self._symm_handle = symm_mem.rendezvous(self._symm, group)
        ptrs = (C.c_void_p * world)(*[int(p) for p in self._symm_handle.buffer_ptrs])
        mc = int(self._symm_handle.multicast_ptr) if multicast else 0
        _check(lib().nrc_comm_attach(self._h, rank, world, ptrs, mc or None))
        dist.barrier(group)  # nobody pushes before everybody has zeroed and attached
        return bool(mc)

    def comm_status(self):
        _check(lib().nrc_comm_status(self._h, _stream()))

    def comm_set_timeout(self, polls: int):
        lib().nrc_comm_set_timeout(self._h, polls)

    def comm_connect(self, group=None):
        """Sets up the in-kernel NVLink all-reduce between the ranks of a torch.distributed group: every rank allocates
        its inbox, the 64-byte IPC handles are all-gathered (plumbing), every rank maps its peers' inboxes."""
        import torch.distributed as dist
        from .dist import exchange_handles
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = C.create_string_buffer(lib().nrc_comm_handle_bytes())
        _check(lib().nrc_comm_init(self._h, rank, world, buf))
        handles = exchange_handles(buf.raw, group)
        _check(lib().nrc_comm_connect(self._h, b"".join(handles)))
        dist.barrier(group)  # nobody pushes before everybody has mapped

    def comm_shutdown(self):
        _check(lib().nrc_comm_shutdown(self._h))

    def comm_world(self) -> int:
        return lib().nrc_comm_world(self._h)

    # ---- learn-an-image
    def image_train_step(self, image_rgba8, seed_x: int, seed_y: int, batch: int = 16384, lr: float = 0.01):
        h, w = image_rgba8.shape[0], image_rgba8.shape[1]
        _check(lib().nrc_image_train_step(self._h, _ptr(image_rgba8), w, h, seed_x & 0xFFFFFFFF, seed_y & 0xFFFFFFFF, batch, lr, _stream()))

    def image_infer(self, width: int, out=None):
        import torch
        if out is None:
            out = torch.empty((width, width, 4), dtype=torch.uint8, device=f"cuda:{self.device}")
        _check(lib().nrc_image_infer(self._h, _ptr(out), width, _stream()))
        return out

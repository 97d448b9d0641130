"""CPU oracle for the VkNRC MLP hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package. The product (``vknrc_b200``) never does; it fails loudly without its CUDA library.

Two libraries sit behind this module (built by ``oracle/Makefile``):

* ``_build/libnrc_oracle.so`` -- ``nrc_oracle.c``, our restatement of the reference's shaders
  (NN_nv.glsl, NRCRecord.glsl, nrc_*.comp), every function citing the lines it follows;
* ``_ref/libvknrc_ref.so``    -- the reference's OWN CPU ``Evaluate`` / ``Train`` (test/main.cpp:11-74) compiled
  unmodified from /root/reference (only buildable where that tree exists; the prebuilt file travels to the GPU box).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHT_COUNT = 20672  # src/VkNRCState.hpp:25
ACC_FP16_CHUNK16 = 0
ACC_FP32 = 1
LOSS_L2 = 0
LOSS_RELATIVE_L2_LUMINANCE = 1

_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


class OptimizerState(C.Structure):  # src/VkNRCState.cpp:25-28
    _fields_ = [("t", C.c_uint32), ("beta1_t", C.c_float), ("beta2_t", C.c_float), ("alpha_t", C.c_float),
                ("alpha_t_1", C.c_float)]

    @classmethod
    def initial(cls):  # src/VkNRCState.cpp:50
        return cls(0, 1.0, 1.0, 1.0, 0.0)


OPT_ENTRY_DTYPE = np.dtype([("m", "<f4"), ("v", "<f4"), ("weight", "<f4"), ("ema_weight", "<f4")])  # VkNRCState.cpp:29-31


def build(force: bool = False) -> None:
    """Compile the oracle (and, where /root/reference exists, the reference's own CPU MLP)."""
    args = ["make", "-C", _HERE] + (["-B"] if force else [])
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_build", "libnrc_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.nrc_oracle_encode_batch.argtypes = [_f32p, C.c_uint64, _u16p]
        L.nrc_oracle_encode_oneblob32.argtypes = [C.c_float, C.c_float, _u16p]
        L.nrc_oracle_learn_image_uv.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float),
                                                C.POINTER(C.c_float)]
        L.nrc_oracle_dst_screen.argtypes = [C.c_uint32, C.c_uint32]
        L.nrc_oracle_dst_screen.restype = C.c_uint32
        L.nrc_oracle_dst_train.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.nrc_oracle_dst_train.restype = C.c_uint32
        L.nrc_oracle_dst_decode.argtypes = [C.c_uint32] + [C.POINTER(C.c_uint32)] * 4
        L.nrc_oracle_forward.argtypes = [_u16p, _u16p, C.c_uint64, C.c_int, C.c_void_p, _f32p]
        L.nrc_oracle_evaluate.argtypes = [_u16p, _u16p, C.c_uint64, C.c_int, C.c_int, _u16p]
        L.nrc_oracle_loss.argtypes = [_f32p, _f32p, C.c_uint64, C.c_int]
        L.nrc_oracle_loss.restype = C.c_double
        L.nrc_oracle_gradient.argtypes = [_u16p, _u16p, _f32p, C.c_uint64, C.c_int, C.c_float, C.c_int, _f32p,
                                          C.c_void_p]
        L.nrc_oracle_prepare.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(OptimizerState)]
        L.nrc_oracle_prepare.restype = C.c_uint32
        L.nrc_oracle_optimize.argtypes = [C.c_uint32, C.POINTER(OptimizerState), C.c_void_p, _f32p, _u16p, C.c_void_p,
                                          C.c_int]
        L.nrc_oracle_sgd.argtypes = [_f32p, _f32p, _u16p, C.c_float, C.c_float]
        L.nrc_oracle_scatter.argtypes = [_f32p, _u32p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32,
                                         C.POINTER(C.c_void_p)]
        L.nrc_oracle_unpack_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, _f32p]
        _lib = L
    return _lib


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libvknrc_ref.so"))


def ref() -> C.CDLL:
    """The reference's own CPU MLP (test/main.cpp:11-74), compiled unmodified."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libvknrc_ref.so")
        if not os.path.exists(path):
            build()
        R = C.CDLL(path)
        R.vknrc_ref_evaluate.argtypes = [_u16p, _u16p, C.c_uint64, _u16p]
        R.vknrc_ref_train.argtypes = [_u16p, _u16p, _u16p, C.c_uint64, _f32p]
        _ref = R
    return _ref


# ---------------------------------------------------------------------------------------------------------------------
# numpy-facing helpers. fp16 buffers are passed as uint16 bit patterns or np.float16 (viewed).
# ---------------------------------------------------------------------------------------------------------------------
def _bits(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype == np.float16:
        a = a.view(np.uint16)
    assert a.dtype == np.uint16, a.dtype
    return a


def encode(unpacked14: np.ndarray) -> np.ndarray:
    """NRCInputEncode (NRCRecord.glsl:78-95): [n,14] fp32 -> [n,64] fp16."""
    x = np.ascontiguousarray(unpacked14, dtype=np.float32).reshape(-1, 14)
    out = np.empty((x.shape[0], 64), np.uint16)
    lib().nrc_oracle_encode_batch(x, x.shape[0], out)
    return out.view(np.float16)


def encode_oneblob32(uv: np.ndarray) -> np.ndarray:
    """learn-an-image encoding (mlp_learning_an_image/gradient.comp:33-44): [n,2] -> [n,64] fp16."""
    uv = np.asarray(uv, np.float32).reshape(-1, 2)
    out = np.empty((uv.shape[0], 64), np.uint16)
    for i in range(uv.shape[0]):
        lib().nrc_oracle_encode_oneblob32(float(uv[i, 0]), float(uv[i, 1]), out[i])
    return out.view(np.float16)


def learn_image_uv(seed_x: int, seed_y: int, n: int) -> np.ndarray:
    """uv of sample gid (mlp_learning_an_image/gradient.comp:15-24, 47-48)."""
    out = np.empty((n, 2), np.float32)
    u, v = C.c_float(), C.c_float()
    for g in range(n):
        lib().nrc_oracle_learn_image_uv(seed_x & 0xFFFFFFFF, seed_y & 0xFFFFFFFF, g, C.byref(u), C.byref(v))
        out[g] = (u.value, v.value)
    return out


def forward(weights, inputs, mode=ACC_FP32, want_acts=False):
    """Returns y [n,3] fp32 (exactly widened fp16) and optionally the six activation sets [6,n,64] fp16."""
    w, x = _bits(weights).reshape(-1), _bits(inputs).reshape(-1, 64)
    n = x.shape[0]
    y = np.empty((n, 3), np.float32)
    acts = np.empty((6, n, 64), np.uint16) if want_acts else None
    lib().nrc_oracle_forward(w, x, n, mode, acts.ctypes.data if want_acts else None, y)
    return (y, acts.view(np.float16)) if want_acts else y


def evaluate(weights, inputs, mode=ACC_FP32, clamp=False) -> np.ndarray:
    """test/evaluate_NV.comp semantics: [n,3] fp16 outputs (clamp=True adds nrc_inference.comp:48)."""
    w, x = _bits(weights).reshape(-1), _bits(inputs).reshape(-1, 64)
    out = np.empty((x.shape[0], 3), np.uint16)
    lib().nrc_oracle_evaluate(w, x, x.shape[0], mode, int(clamp), out)
    return out.view(np.float16)


def gradient(weights, inputs, targets, loss=LOSS_L2, loss_scale=1.0, mode=ACC_FP32, want_y=False):
    """dW [20672] fp32 summed over the batch (un-normalised, like the reference's `gradients` buffer)."""
    w, x = _bits(weights).reshape(-1), _bits(inputs).reshape(-1, 64)
    t = np.ascontiguousarray(np.asarray(targets, np.float32).reshape(-1, 3))
    assert t.shape[0] == x.shape[0]
    dw = np.zeros(WEIGHT_COUNT, np.float32)
    y = np.empty((x.shape[0], 3), np.float32) if want_y else None
    lib().nrc_oracle_gradient(w, x, t, x.shape[0], loss, loss_scale, mode, dw, y.ctypes.data if want_y else None)
    return (dw, y) if want_y else dw


def loss_value(y, targets, loss=LOSS_L2) -> float:
    y = np.ascontiguousarray(y, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(targets, np.float32).reshape(-1, 3)
    return float(lib().nrc_oracle_loss(y, t, y.shape[0], loss))


class Optimizer:
    """Adam + EMA exactly as nrc_train_prepare.comp:16-28 + nrc_optimize.comp:32-54, state as VkNRCState.cpp:46-58."""

    def __init__(self, fp32_weights: np.ndarray):
        w = np.asarray(fp32_weights, np.float32).reshape(WEIGHT_COUNT)
        self.state = OptimizerState.initial()
        self.entries = np.zeros(WEIGHT_COUNT, OPT_ENTRY_DTYPE)
        self.entries["weight"] = w
        self.entries["ema_weight"] = w
        self.weights = w.astype(np.float16).view(np.uint16).copy()  # half_float RTNE, VkNRCState.cpp:53
        self.use_weights = self.weights.copy()

    def step(self, gradients: np.ndarray, count: int, write_use_weights: bool, use_ema: bool, batch_cap=16384) -> int:
        count = lib().nrc_oracle_prepare(count, batch_cap, C.byref(self.state))
        g = np.ascontiguousarray(gradients, np.float32)
        lib().nrc_oracle_optimize(count, C.byref(self.state), self.entries.ctypes.data, g, self.weights,
                                  self.use_weights.ctypes.data if write_use_weights else None, int(use_ema))
        return count


def sgd(fp_weights: np.ndarray, gradients: np.ndarray, weights16: np.ndarray, lr=0.01, batch=16384.0) -> None:
    lib().nrc_oracle_sgd(fp_weights, np.ascontiguousarray(gradients, np.float32), _bits(weights16), lr, batch)


def dst_screen(x: int, y: int) -> int:
    return lib().nrc_oracle_dst_screen(x, y)


def dst_train(b: int, l: int, r: int) -> int:
    return lib().nrc_oracle_dst_train(b, l, r)


def dst_decode(e: int):
    t, a, b, c = (C.c_uint32() for _ in range(4))
    lib().nrc_oracle_dst_decode(e, C.byref(t), C.byref(a), C.byref(b), C.byref(c))
    return t.value, a.value, b.value, c.value


def scatter(predict, dst, bias_factor_r, factor_gb, width, train_records):
    """nrc_inference.comp:48-73 applied in place. train_records: list of 4 float32 arrays [cap,10] (40 B records)."""
    p = np.ascontiguousarray(predict, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(dst, np.uint32)
    ptrs = (C.c_void_p * 4)(*[r.ctypes.data for r in train_records])
    lib().nrc_oracle_scatter(p, d, p.shape[0], bias_factor_r.ctypes.data, factor_gb.ctypes.data, width, ptrs)


# --- the reference's own CPU code ------------------------------------------------------------------------------------
def ref_evaluate(weights, inputs) -> np.ndarray:
    """Reference `Evaluate` (test/main.cpp:11-27): [n,3] fp16."""
    w, x = _bits(weights).reshape(-1), _bits(inputs).reshape(-1, 64)
    out = np.empty((x.shape[0], 3), np.uint16)
    rc = ref().vknrc_ref_evaluate(w, x, x.shape[0], out)
    assert rc == 0
    return out.view(np.float16)


def ref_train(weights, inputs, targets16) -> np.ndarray:
    """Reference `Train` (test/main.cpp:29-74): [20672] fp32. A debugging sketch; prints to stdout."""
    w, x, t = _bits(weights).reshape(-1), _bits(inputs).reshape(-1, 64), _bits(targets16).reshape(-1)
    dw = np.empty(WEIGHT_COUNT, np.float32)
    rc = ref().vknrc_ref_train(w, x, t, x.shape[0], dw)
    assert rc == 0
    return dw


# ---------------------------------------------------------------------------------------------------------------------
# Scene buffers + UnpackNRCInput (shader/src/Scene.glsl:8-71, shader/src/NRCRecord.glsl:98-125)
# ---------------------------------------------------------------------------------------------------------------------
MATERIAL_DTYPE = np.dtype([("diffuse", "<f4", 3), ("diffuse_texture_id", "<u4"), ("specular", "<f4", 3), ("specular_texture_id", "<u4"),
                           ("emission", "<f4", 3), ("emission_texture_id", "<u4"), ("metallic", "<f4"), ("roughness", "<f4"),
                           ("ior", "<f4"), ("_pad", "<u4")])
assert MATERIAL_DTYPE.itemsize == 64  # std430 array stride of Scene.glsl's Material


class _CTexture(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class _CScene(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertex_indices", C.c_void_p), ("texcoords", C.c_void_p), ("texcoord_indices", C.c_void_p),
                ("materials", C.c_void_p), ("material_ids", C.c_void_p), ("transforms", C.c_void_p), ("textures", C.c_void_p),
                ("texture_count", C.c_uint32)]


class Scene:
    """Host-side scene in the reference's buffer layouts (numpy arrays, kept alive here)."""

    def __init__(self, vertices, vertex_indices, texcoords, texcoord_indices, materials, material_ids, transforms, textures):
        self.vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        self.vertex_indices = np.ascontiguousarray(vertex_indices, np.uint32).reshape(-1, 3)
        self.texcoords = np.ascontiguousarray(texcoords, np.float32).reshape(-1, 2)
        self.texcoord_indices = np.ascontiguousarray(texcoord_indices, np.uint32).reshape(-1, 3)
        self.materials = np.ascontiguousarray(materials, MATERIAL_DTYPE)
        self.material_ids = np.ascontiguousarray(material_ids, np.uint32)
        self.transforms = np.ascontiguousarray(transforms, np.float32).reshape(-1, 12)
        self.textures = [np.ascontiguousarray(t, np.uint8) for t in textures]  # [H, W, 4] sRGB
        self._ctex = (_CTexture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            self._ctex[i] = _CTexture(t.ctypes.data, t.shape[1], t.shape[0])
        self._c = _CScene(self.vertices.ctypes.data, self.vertex_indices.ctypes.data, self.texcoords.ctypes.data,
                          self.texcoord_indices.ctypes.data, self.materials.ctypes.data, self.material_ids.ctypes.data,
                          self.transforms.ctypes.data, C.addressof(self._ctex), len(self.textures))


def unpack(scene: Scene, packed: np.ndarray) -> np.ndarray:
    """UnpackNRCInput: [n,4] uint32 PackedNRCInput -> [n,14] fp32 (UnpackedNRCInput order)."""
    pk = np.ascontiguousarray(packed, np.uint32).reshape(-1, 4)
    out = np.empty((pk.shape[0], 14), np.float32)
    lib().nrc_oracle_unpack_batch(C.addressof(scene._c), pk.ctypes.data, pk.shape[0], 16, out)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own GLSL shaders compiled as C++ (oracle/glsl, built into _ref_glsl/libvknrc_glsl.so by the Makefile
# where /root/reference exists; the prebuilt file travels to the GPU box). This is what pins encode / unpack / dst codec /
# loss gradients / backward / dW / optimizer: reference SOURCE executed, not a restatement of it.
# ---------------------------------------------------------------------------------------------------------------------
_glsl = None


class _GlslTexture(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class _GlslScene(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertex_indices", C.c_void_p), ("texcoords", C.c_void_p), ("texcoord_indices", C.c_void_p),
                ("materials", C.c_void_p), ("material_ids", C.c_void_p), ("transforms", C.c_void_p), ("textures", C.c_void_p),
                ("texture_count", C.c_uint32), ("material_count", C.c_uint32)]


def glsl_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref_glsl", "libvknrc_glsl.so"))


def glsl() -> C.CDLL:
    global _glsl
    if _glsl is None:
        path = os.path.join(_HERE, "_ref_glsl", "libvknrc_glsl.so")
        if not os.path.exists(path):
            build()
        G = C.CDLL(path)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        G.glsl_NRCInputEncode.argtypes = [_f32p, u64, _u16p]
        G.glsl_UnpackNRCInput.argtypes = [vp, _u32p, u64, _f32p]
        G.glsl_EncodeNRCEvalDstScreen.argtypes = [u32, u32]
        G.glsl_EncodeNRCEvalDstScreen.restype = u32
        G.glsl_EncodeNRCEvalDstTrain.argtypes = [u32, u32, u32]
        G.glsl_EncodeNRCEvalDstTrain.restype = u32
        G.glsl_DecodeNRCEvalDst.argtypes = [u32] + [C.POINTER(u32)] * 4
        G.glsl_evaluate_NV.argtypes = [_u16p, _u16p, u64, _u16p, C.c_int]
        G.glsl_train_NV.argtypes = [_u16p, _f32p, _u16p, _u16p, u64, C.c_int]
        G.glsl_nrc_inference.argtypes = [vp, vp, u32, _u16p, _f32p, _f32p, u32, u32, C.POINTER(vp), C.c_int]
        G.glsl_nrc_gradient.argtypes = [vp, vp, u32, _u16p, _f32p, C.c_int]
        G.glsl_nrc_train_prepare.argtypes = [C.POINTER(u32), C.POINTER(u32 * 3), C.POINTER(OptimizerState)]
        G.glsl_nrc_optimize.argtypes = [_u16p, vp, _f32p, vp, u32, C.POINTER(OptimizerState), u32]
        G.glsl_image_gradient.argtypes = [_u16p, _f32p, vp, u32, u32, u32, u32, u32, C.c_int]
        G.glsl_image_optimize.argtypes = [_u16p, _f32p, _f32p]
        G.glsl_image_inference.argtypes = [_u16p, vp, C.c_int]
        G.glsl_image_uv.argtypes = [u32, u32, u32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        G.glsl_image_oneblob32.argtypes = [C.c_float, C.c_float, _u16p]
        _glsl = G
    return _glsl


def _glsl_scene(scene: "Scene"):
    tex = (_GlslTexture * max(1, len(scene.textures)))()
    for i, t in enumerate(scene.textures):
        tex[i] = _GlslTexture(t.ctypes.data, t.shape[1], t.shape[0])
    s = _GlslScene(scene.vertices.ctypes.data, scene.vertex_indices.ctypes.data, scene.texcoords.ctypes.data, scene.texcoord_indices.ctypes.data,
                   scene.materials.ctypes.data, scene.material_ids.ctypes.data, scene.transforms.ctypes.data, C.addressof(tex),
                   len(scene.textures), scene.materials.shape[0])
    return s, tex  # (keep `tex` alive while `s` is in use)


def glsl_encode(unpacked14) -> np.ndarray:
    """NRCInputEncode of shader/src/NRCRecord.glsl:78-95, the reference's source: [n,14] fp32 -> [n,64] fp16."""
    x = np.ascontiguousarray(unpacked14, np.float32).reshape(-1, 14)
    out = np.empty((x.shape[0], 64), np.uint16)
    glsl().glsl_NRCInputEncode(x, x.shape[0], out)
    return out.view(np.float16)


def glsl_unpack(scene: "Scene", packed) -> np.ndarray:
    """UnpackNRCInput of shader/src/NRCRecord.glsl:98-125 over Scene.glsl, the reference's source."""
    pk = np.ascontiguousarray(packed, np.uint32).reshape(-1, 4)
    out = np.empty((pk.shape[0], 14), np.float32)
    s, keep = _glsl_scene(scene)
    glsl().glsl_UnpackNRCInput(C.addressof(s), pk, pk.shape[0], out)
    return out


def glsl_dst_screen(x, y):
    return glsl().glsl_EncodeNRCEvalDstScreen(x, y)


def glsl_dst_train(b, l, r):
    return glsl().glsl_EncodeNRCEvalDstTrain(b, l, r)


def glsl_dst_decode(e):
    t, a, b, c = (C.c_uint32() for _ in range(4))
    glsl().glsl_DecodeNRCEvalDst(e, C.byref(t), C.byref(a), C.byref(b), C.byref(c))
    return t.value, a.value, b.value, c.value


def glsl_evaluate_nv(weights, inputs, parallel=True) -> np.ndarray:
    """test/evaluate_NV.comp dispatched over n/128 workgroups: [n,3] fp16."""
    w, x = _bits(weights).reshape(-1), _bits(inputs).reshape(-1, 64)
    out = np.empty((x.shape[0], 3), np.uint16)
    assert glsl().glsl_evaluate_NV(w, x, x.shape[0], out, int(parallel)) == 0
    return out.view(np.float16)


def glsl_train_nv(weights, inputs, targets16, dw=None, parallel=False) -> np.ndarray:
    """test/train_NV.comp (L2 loss): dW [20672] fp32, accumulated into `dw` by fp32 atomics as the shader does."""
    w, x, t = _bits(weights).reshape(-1), _bits(inputs).reshape(-1, 64), _bits(targets16).reshape(-1)
    dw = np.zeros(WEIGHT_COUNT, np.float32) if dw is None else dw
    assert glsl().glsl_train_NV(w, dw, x, t, x.shape[0], int(parallel)) == 0
    return dw


def glsl_nrc_inference(scene, eval_records, eval_count, weights, bias_factor_r, factor_gb, train_records, parallel=True):
    """shader/src/nrc_inference.comp in place on bias_factor_r [H,W,4] f32 / train_records (4 byte arrays of 40 B records)."""
    s, keep = _glsl_scene(scene)
    h, w_ = bias_factor_r.shape[0], bias_factor_r.shape[1]
    ptrs = (C.c_void_p * 4)(*[r.ctypes.data for r in train_records])
    ev = np.ascontiguousarray(eval_records)
    assert glsl().glsl_nrc_inference(C.addressof(s), ev.ctypes.data, eval_count, _bits(weights).reshape(-1), bias_factor_r.reshape(-1),
                                     np.ascontiguousarray(factor_gb, np.float32).reshape(-1), w_, h, ptrs, int(parallel)) == 0


def glsl_nrc_gradient(scene, train_records, count, weights, dw=None, parallel=False) -> np.ndarray:
    """shader/src/nrc_gradient.comp: dW accumulated by fp32 atomics (un-normalised)."""
    s, keep = _glsl_scene(scene)
    dw = np.zeros(WEIGHT_COUNT, np.float32) if dw is None else dw
    tr = np.ascontiguousarray(train_records)
    assert glsl().glsl_nrc_gradient(C.addressof(s), tr.ctypes.data, count, _bits(weights).reshape(-1), dw, int(parallel)) == 0
    return dw


class GlslOptimizer:
    """nrc_train_prepare.comp + nrc_optimize.comp of the reference driven like src/rg/NNTrain.cpp: same interface as Optimizer."""

    def __init__(self, fp32_weights):
        w = np.asarray(fp32_weights, np.float32).reshape(WEIGHT_COUNT)
        self.state = OptimizerState.initial()
        self.entries = np.zeros(WEIGHT_COUNT, OPT_ENTRY_DTYPE)
        self.entries["weight"] = w
        self.entries["ema_weight"] = w
        self.weights = w.astype(np.float16).view(np.uint16).copy()
        self.use_weights = self.weights.copy()

    def step(self, gradients, count, write_use_weights, use_ema) -> int:
        c, cmd = C.c_uint32(count), (C.c_uint32 * 3)()
        glsl().glsl_nrc_train_prepare(C.byref(c), C.byref(cmd), C.byref(self.state))
        self.last_command = tuple(cmd)
        g = np.ascontiguousarray(gradients, np.float32)
        glsl().glsl_nrc_optimize(self.weights, self.use_weights.ctypes.data if write_use_weights else None, g, self.entries.ctypes.data,
                                 c.value, C.byref(self.state), int(use_ema))
        return c.value


def glsl_image_uv(seed_x, seed_y, n) -> np.ndarray:
    out = np.empty((n, 2), np.float32)
    u, v = C.c_float(), C.c_float()
    for g in range(n):
        glsl().glsl_image_uv(seed_x & 0xFFFFFFFF, seed_y & 0xFFFFFFFF, g, C.byref(u), C.byref(v))
        out[g] = (u.value, v.value)
    return out


def glsl_image_oneblob32(uv) -> np.ndarray:
    uv = np.asarray(uv, np.float32).reshape(-1, 2)
    out = np.empty((uv.shape[0], 64), np.uint16)
    for i in range(uv.shape[0]):
        glsl().glsl_image_oneblob32(float(uv[i, 0]), float(uv[i, 1]), out[i])
    return out.view(np.float16)


def glsl_image_gradient(weights, image_rgba8, seed_x, seed_y, n=16384, dw=None, parallel=False) -> np.ndarray:
    img = np.ascontiguousarray(image_rgba8, np.uint8)
    dw = np.zeros(WEIGHT_COUNT, np.float32) if dw is None else dw
    assert glsl().glsl_image_gradient(_bits(weights).reshape(-1), dw, img.ctypes.data, img.shape[1], img.shape[0], seed_x & 0xFFFFFFFF,
                                      seed_y & 0xFFFFFFFF, n, int(parallel)) == 0
    return dw


def glsl_image_optimize(weights16, fp_weights, gradients):
    glsl().glsl_image_optimize(_bits(weights16), fp_weights, np.ascontiguousarray(gradients, np.float32))


def glsl_image_inference(weights, parallel=True) -> np.ndarray:
    out = np.empty((640, 640, 4), np.uint8)
    assert glsl().glsl_image_inference(_bits(weights).reshape(-1), out.ctypes.data, int(parallel)) == 0
    return out

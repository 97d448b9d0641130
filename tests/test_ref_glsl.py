"""CPU: the oracle (oracle/nrc_oracle.c, a restatement) against outputs of the reference's OWN GLSL shaders, compiled as C++
by oracle/Makefile (oracle/glsl) and stored in tests/golden/nrc_golden_v2.npz by tools/make_golden_v2.py. This is what pins
the rows the reference's CPU `Evaluate` cannot: encoding, scene gather, dst codec, loss gradients, backward pass, dW
reduction, optimizer, scatter, learn-an-image. Tolerances are stated per row; "bit-exact" means np.array_equal on the bits."""
import os

import numpy as np
import pytest

from util import layer_rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g2():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "nrc_golden_v2.npz")))


@pytest.fixture(scope="module")
def scene(g2, oracle_mod):
    return golden_scene(g2, oracle_mod)


def golden_scene(g2, oracle_mod):
    texs = [g2[f"scene_texture{i}"] for i in range(3)]
    mats = np.ascontiguousarray(g2["scene_materials"]).view(oracle_mod.MATERIAL_DTYPE).reshape(-1)
    return oracle_mod.Scene(g2["scene_vertices"], g2["scene_vertex_indices"], g2["scene_texcoords"], g2["scene_texcoord_indices"], mats,
                            g2["scene_material_ids"], g2["scene_transforms"], texs)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint16 if a.dtype.itemsize == 2 else np.uint32)


def test_encode_is_bit_exact(g2, oracle_mod):  # A5: NRCRecord.glsl:47-95
    assert np.array_equal(bits(oracle_mod.encode(g2["records14"])), bits(g2["glsl_encoded"]))


def test_unpack(g2, oracle_mod, scene):  # A1, A4: NRCRecord.glsl:98-125 over Scene.glsl:50-64
    got, ref = oracle_mod.unpack(scene, g2["packed_inputs"]), g2["glsl_unpacked"]
    # position: the shader's transform / barycentric sums are written differently (vec4 * mat3x4 vs dot products): a few fp32 ulps of
    # coordinates of size ~4. scattered_dir: glm's unpackUnorm2x16 multiplies by fl(1/65535), the oracle divides: 1 ulp.
    # normal, roughness: identical operations. colours: the bilinear blend is associated differently (E2): ulps of [0,1].
    assert np.abs(got[:, 0:3] - ref[:, 0:3]).max() <= 2e-6
    assert np.abs(got[:, 3:5] - ref[:, 3:5]).max() <= 6e-8
    assert np.abs(got[:, 5:7] - ref[:, 5:7]).max() <= 1.2e-7 and np.array_equal(got[:, 7], ref[:, 7])
    assert np.abs(got[:, 8:14] - ref[:, 8:14]).max() <= 2e-5


def test_dst_codec_is_bit_exact(g2, oracle_mod):  # A2: NRCRecord.glsl:19-33
    assert [oracle_mod.dst_screen(int(a), int(b)) for a, b in g2["dst_xy"]] == g2["glsl_dst_screen"].tolist()
    assert [oracle_mod.dst_train(int(a), int(b), int(c)) for a, b, c in g2["dst_blr"]] == g2["glsl_dst_train"].tolist()
    words = np.concatenate([g2["glsl_dst_screen"], g2["glsl_dst_train"]])
    assert [list(oracle_mod.dst_decode(int(e))) for e in words] == g2["glsl_dst_decoded"].tolist()


def test_forward_shader_precision_is_bit_exact(g2, oracle_mod):  # A7-A9, A19: test/evaluate_NV.comp
    w16 = g2["weights_fp32"].astype(np.float16)
    assert np.array_equal(bits(oracle_mod.evaluate(w16, g2["inputs"], oracle_mod.ACC_FP16_CHUNK16)), bits(g2["glsl_evaluate_nv"]))


def test_backward_and_dw_shader_precision(g2, oracle_mod):  # A10-A12: test/train_NV.comp (NNLoadDA3_L2Loss, NNBackwardDA*, NNUpdateDW*)
    w16 = g2["weights_fp32"].astype(np.float16)
    dw = oracle_mod.gradient(w16, g2["inputs"], g2["targets"].astype(np.float32), oracle_mod.LOSS_L2, 1.0, oracle_mod.ACC_FP16_CHUNK16)
    # identical fp16 per-warp partials; only the order of the workgroups' fp32 atomic adds differs
    assert max(layer_rel_err(dw, g2["glsl_train_nv_dw"])) <= 1e-6


def test_gradient_shader_on_train_records(g2, oracle_mod, scene):  # A14: nrc_gradient.comp (+ relative-L2-luminance loss, tail lanes)
    w16 = g2["weights_fp32"].astype(np.float16)
    tr = np.ascontiguousarray(g2["train_records"]).view(np.float32).reshape(-1, 10)
    cnt = int(g2["train_count"])
    pk = np.ascontiguousarray(tr[:cnt, 6:10]).view(np.uint32)
    enc = oracle_mod.encode(oracle_mod.unpack(scene, pk))
    dw = oracle_mod.gradient(w16, enc, np.ascontiguousarray(tr[:cnt, 0:3]), oracle_mod.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle_mod.ACC_FP16_CHUNK16)
    # the gather differs by fp32 ulps (test_unpack), which the top frequency octaves and the fp16 roundings amplify
    assert max(layer_rel_err(dw, g2["glsl_nrc_gradient_dw"])) <= 2e-2
    # with the reference's own unpacked inputs the restated loss / backward / dW agree to the atomics' summation order
    if oracle_mod.glsl_available():
        enc_ref = oracle_mod.glsl_encode(oracle_mod.glsl_unpack(scene, pk))
        dw2 = oracle_mod.gradient(w16, enc_ref, np.ascontiguousarray(tr[:cnt, 0:3]), oracle_mod.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle_mod.ACC_FP16_CHUNK16)
        assert max(layer_rel_err(dw2, g2["glsl_nrc_gradient_dw"])) <= 1e-6


def test_optimizer_is_bit_exact(g2, oracle_mod):  # A15, A16: nrc_train_prepare.comp + nrc_optimize.comp (both variants), Q6-Q10
    opt = oracle_mod.Optimizer(g2["weights_fp32"])
    for i, (cnt, wu, ema) in enumerate(g2["opt_steps"]):
        clamped = opt.step(g2["opt_gradient"] * np.float32(i + 1), int(cnt), bool(wu), bool(ema))
        s = opt.state
        assert [s.t, s.beta1_t, s.beta2_t, s.alpha_t, s.alpha_t_1] == g2["glsl_opt_states"][i].tolist()
        assert (clamped + 127) // 128 == g2["glsl_opt_commands"][i][0]  # the indirect dispatch size of nrc_train_prepare.comp:21
    assert np.array_equal(opt.entries.view(np.uint32).reshape(-1, 4), bits(g2["glsl_opt_entries"]))
    assert np.array_equal(opt.weights, g2["glsl_opt_weights"]) and np.array_equal(opt.use_weights, g2["glsl_opt_use_weights"])


def test_inference_scatter(g2, oracle_mod, scene):  # A13: nrc_inference.comp:30-74 (screen composite + feedback into the train targets)
    w16 = g2["weights_fp32"].astype(np.float16)
    ev = np.ascontiguousarray(g2["frame_eval_records"]).view(np.uint32).reshape(-1, 5)
    pred = oracle_mod.evaluate(w16, oracle_mod.encode(oracle_mod.unpack(scene, np.ascontiguousarray(ev[:, 1:5]))), oracle_mod.ACC_FP16_CHUNK16, clamp=True)
    bf = g2["frame_bias_factor_r"].copy()
    trs = [np.ascontiguousarray(g2[f"frame_train_records{b}"]).view(np.float32).reshape(-1, 10).copy() for b in range(4)]
    oracle_mod.scatter(pred.astype(np.float32), np.ascontiguousarray(ev[:, 0]), bf, g2["frame_factor_gb"], bf.shape[1], trs)
    scale = float(np.abs(pred.astype(np.float32)).max())
    # indexing is exact: every pixel / record the shader touched is touched, nothing else moves (alpha is zeroed by the store)
    ref_bf = g2["glsl_frame_bias_factor_r"]
    assert np.array_equal(bf[..., 3], ref_bf[..., 3]) and np.abs(bf - ref_bf).max() <= 1e-2 * scale
    for b in range(4):
        ref = np.ascontiguousarray(g2[f"glsl_frame_train_records{b}"]).view(np.float32).reshape(-1, 10)
        assert np.array_equal(trs[b][:, 3:].view(np.uint32), ref[:, 3:].view(np.uint32))  # factor + packed input untouched
        touched_ref = ref[:, :3] != np.ascontiguousarray(g2[f"frame_train_records{b}"]).view(np.float32).reshape(-1, 10)[:, :3]
        touched = trs[b][:, :3] != np.ascontiguousarray(g2[f"frame_train_records{b}"]).view(np.float32).reshape(-1, 10)[:, :3]
        assert np.array_equal(touched.any(axis=1), touched_ref.any(axis=1))
        assert np.abs(trs[b][:, :3] - ref[:, :3]).max() <= 1e-2 * max(scale, float(np.abs(ref[:, :3]).max()))


def test_learn_an_image_kernels(g2, oracle_mod):  # A20: test/mlp_learning_an_image/{gradient,optimize,inference}.comp
    sx, sy = (int(v) for v in g2["image_seed"])
    uv = oracle_mod.learn_image_uv(sx, sy, 256)
    assert np.array_equal(uv, g2["glsl_image_uv"])  # pcg2d stream
    assert np.array_equal(bits(oracle_mod.encode_oneblob32(uv)), bits(g2["glsl_image_encoded"]))  # one-blob-32 incl. the 32 / 4 radii (Q14)
    fp, h16 = g2["weights_fp32"].copy(), g2["weights_fp32"].astype(np.float16).view(np.uint16).copy()
    oracle_mod.sgd(fp, g2["glsl_image_gradient_dw"], h16, 0.01, 16384.0)
    assert np.array_equal(fp, g2["glsl_image_sgd_fp"]) and np.array_equal(h16, g2["glsl_image_sgd_weights"])  # optimize.comp:21-29


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref_glsl", "libvknrc_glsl.so")), reason="oracle/_ref_glsl not built")
def test_golden_v2_is_reproducible_from_the_compiled_shaders(g2, oracle_mod, scene):
    """The committed fixture equals what the reference-compiled library produces now (tools/make_golden_v2.py is current), and
    the per-subgroup product memo of the emulator changes nothing."""
    w16 = g2["weights_fp32"].astype(np.float16)
    assert np.array_equal(bits(oracle_mod.glsl_encode(g2["records14"])), bits(g2["glsl_encoded"]))
    assert np.array_equal(oracle_mod.glsl_unpack(scene, g2["packed_inputs"]), g2["glsl_unpacked"])
    assert np.array_equal(bits(oracle_mod.glsl_evaluate_nv(w16, g2["inputs"][:256])), bits(g2["glsl_evaluate_nv"][:256]))
    assert np.array_equal(oracle_mod.glsl_train_nv(w16, g2["inputs"], g2["targets"]), g2["glsl_train_nv_dw"])
    os.environ["GLSL_EMU_NO_MEMO"] = "1"
    try:
        assert np.array_equal(bits(oracle_mod.glsl_evaluate_nv(w16, g2["inputs"][:128])), bits(g2["glsl_evaluate_nv"][:128]))
    finally:
        del os.environ["GLSL_EMU_NO_MEMO"]

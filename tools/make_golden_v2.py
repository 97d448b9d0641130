"""Generates tests/golden/nrc_golden_v2.npz IN THIS CONTAINER (needs /root/reference): outputs of the reference's OWN GLSL
shaders - NRCRecord.glsl, Scene.glsl, NN_nv.glsl, nrc_inference / nrc_gradient / nrc_optimize / nrc_train_prepare.comp, the
test kernels and the learn-an-image kernels - compiled as C++ by oracle/Makefile (oracle/glsl: lexical translation + glm +
a GLSL environment shim) and run on the CPU on seeded inputs. Every array named `glsl_*` was produced by reference source;
the inputs they were produced from are stored next to them. Emulation assumptions (E1)-(E4): oracle/glsl/glsl_shim.hpp."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from vknrc_b200 import synth  # noqa: E402


def main():
    oracle.build()
    assert oracle.glsl_available(), "needs /root/reference to build oracle/_ref_glsl"
    rng = np.random.default_rng(20261017)
    out = {}
    w32 = (rng.standard_normal(oracle.WEIGHT_COUNT) * np.sqrt(2.0 / 64.0)).astype(np.float32)
    w16 = w32.astype(np.float16)
    out["weights_fp32"] = w32
    # ---- A5 NRCInputEncode: positions far outside [0,1] (Q12), every other field in [0,1] incl. the exact end points
    n = 512
    rec = np.concatenate([rng.uniform(-40, 40, (n, 3)), rng.uniform(0, 1, (n, 11))], axis=1).astype(np.float32)
    rec[:8, 3:8] = np.array([0, 1, 0.25, 0.5, 0.75, 1e-7, 1 - 1e-7, 0.125])[:, None]
    out["records14"], out["glsl_encoded"] = rec, oracle.glsl_encode(rec)
    # ---- A1/A4 UnpackNRCInput over a small scene (stored, not re-generated: LAPACK QR is not bit-stable across machines)
    sa = synth.make_scene_arrays(7, n_prims=200, n_instances=3, n_materials=12, n_textures=3)
    for k in ("vertices", "vertex_indices", "texcoords", "texcoord_indices", "material_ids", "transforms"):
        out["scene_" + k] = sa[k]
    out["scene_materials"] = sa["materials"].view(np.uint8).reshape(-1, 64)
    for i, t in enumerate(sa["textures"]):
        out[f"scene_texture{i}"] = t
    sc = oracle.Scene(sa["vertices"], sa["vertex_indices"], sa["texcoords"], sa["texcoord_indices"], sa["materials"], sa["material_ids"],
                      sa["transforms"], sa["textures"])
    pk = synth.random_packed_inputs(8, n, 200, 3)
    pk[:4, 2] = [0, 0xFFFF, 0xFFFF0000, 0x7FFF8000]  # barycentric corners
    out["packed_inputs"], out["glsl_unpacked"] = pk, oracle.glsl_unpack(sc, pk)
    # ---- A2 dst codec
    xy = np.array([[0, 0], [1919, 1079], [32767, 32767], [5, 7], [1, 0], [0, 1]], np.uint32)
    blr = np.array([[0, 0, 0], [3, 16383, 16383], [2, 100, 7000], [1, 5, 5]], np.uint32)
    out["dst_xy"], out["dst_blr"] = xy, blr
    out["glsl_dst_screen"] = np.array([oracle.glsl_dst_screen(int(a), int(b)) for a, b in xy], np.uint32)
    out["glsl_dst_train"] = np.array([oracle.glsl_dst_train(int(a), int(b), int(c)) for a, b, c in blr], np.uint32)
    out["glsl_dst_decoded"] = np.array([oracle.glsl_dst_decode(int(e)) for e in np.concatenate([out["glsl_dst_screen"], out["glsl_dst_train"]])], np.uint32)
    # ---- A7-A9, A19 test/evaluate_NV.comp and A10-A12 test/train_NV.comp (L2 loss) on pre-encoded inputs
    x = rng.uniform(0, 1, (n, 64)).astype(np.float32).astype(np.float16)
    t16 = rng.uniform(0, 1, (n, 3)).astype(np.float32).astype(np.float16)
    out["inputs"], out["targets"] = x, t16
    out["glsl_evaluate_nv"] = oracle.glsl_evaluate_nv(w16, x)
    out["glsl_train_nv_dw"] = oracle.glsl_train_nv(w16, x, t16)
    # ---- A14 nrc_gradient.comp: 40-byte records + scene, relative-L2-luminance loss, count not a multiple of 128 (Q11)
    tr = synth.train_records(9, n, 200, 3)
    out["train_records"], out["train_count"] = tr.view(np.uint8).reshape(-1, 40), np.uint32(500)
    out["glsl_nrc_gradient_dw"] = oracle.glsl_nrc_gradient(sc, tr, 500, w16)
    # ---- A15/A16 nrc_train_prepare.comp + nrc_optimize.comp: empty, over-full and partial batches, both variants, EMA on / off
    g = (rng.standard_normal(oracle.WEIGHT_COUNT) * 50).astype(np.float32)
    g[:4] = [np.nan, np.inf, -np.inf, 0.0]  # nrc_optimize.comp:37-38
    steps = np.array([[16384, 0, 0], [0, 1, 1], [100000, 1, 1], [777, 1, 0], [5, 0, 1]], np.uint32)  # count, write_use_weights, use_ema
    opt = oracle.GlslOptimizer(w32)
    states, cmds = [], []
    for i, (cnt, wu, ema) in enumerate(steps):
        opt.step(g * np.float32(i + 1), int(cnt), bool(wu), bool(ema))
        states.append([opt.state.t, opt.state.beta1_t, opt.state.beta2_t, opt.state.alpha_t, opt.state.alpha_t_1])
        cmds.append(opt.last_command)
    out["opt_gradient"], out["opt_steps"] = g, steps
    out["glsl_opt_states"], out["glsl_opt_commands"] = np.array(states, np.float64), np.array(cmds, np.uint32)
    out["glsl_opt_entries"], out["glsl_opt_weights"], out["glsl_opt_use_weights"] = opt.entries.view(np.float32).reshape(-1, 4), opt.weights.copy(), opt.use_weights.copy()
    # ---- A13 nrc_inference.comp: a path-structured frame (screen queries + train-tail queries feeding back into the targets)
    W, H, cap = 48, 32, 256
    fr = synth.frame_records(10, W, H, 200, 3, train_probability=0.3, batch_size=cap)
    ev, n_ev = fr["eval_records"], int(fr["eval_count"])
    bf = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
    gb = rng.uniform(0, 1, (H, W, 2)).astype(np.float32)
    out["frame_eval_records"], out["frame_bias_factor_r"], out["frame_factor_gb"] = ev.view(np.uint8).reshape(-1, 20), bf.copy(), gb
    for b in range(4):
        out[f"frame_train_records{b}"] = fr["train_records"][b].view(np.uint8).reshape(-1, 40).copy()
    trs = [t.copy() for t in fr["train_records"]]
    oracle.glsl_nrc_inference(sc, ev, n_ev, w16, bf, gb, [t.view(np.uint8).reshape(-1) for t in trs])
    out["glsl_frame_bias_factor_r"] = bf
    for b in range(4):
        out[f"glsl_frame_train_records{b}"] = trs[b].view(np.uint8).reshape(-1, 40)
    # ---- A20 learn-an-image kernels
    img = rng.integers(0, 256, (29, 37, 4), dtype=np.uint8)
    out["image_rgba8"], out["image_seed"] = img, np.array([123, 456], np.uint32)
    uv = oracle.glsl_image_uv(123, 456, 256)
    out["glsl_image_uv"], out["glsl_image_encoded"] = uv, oracle.glsl_image_oneblob32(uv)
    out["glsl_image_gradient_dw"] = oracle.glsl_image_gradient(w16, img, 123, 456, 512)
    fp, h16 = w32.copy(), w16.view(np.uint16).copy()
    oracle.glsl_image_optimize(h16, fp, out["glsl_image_gradient_dw"])
    out["glsl_image_sgd_weights"], out["glsl_image_sgd_fp"] = h16, fp
    out["glsl_image_inference_16"] = oracle.glsl_image_inference(w16)[::16, ::16].copy()  # every 16th pixel of the 640 x 640 frame
    path = os.path.join(ROOT, "tests", "golden", "nrc_golden_v2.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

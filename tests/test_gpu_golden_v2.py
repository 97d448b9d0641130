"""GPU (run with -m gpu): the CUDA path, called through the C ABI, against outputs of the reference's OWN GLSL shaders
(tests/golden/nrc_golden_v2.npz, produced by tools/make_golden_v2.py from oracle/_ref_glsl = the reference's shader sources
compiled as C++). One test per SURVEY 8(a) row that the reference's CPU `Evaluate` cannot pin. Where the reference
accumulates in fp16 (Q1, Q4) and this implementation in fp32 (TMEM), the tolerance is the measured distance between the two
precisions of the same network, stated in the test."""
import os

import numpy as np
import pytest

from test_ref_glsl import golden_scene
from util import layer_rel_err, out_err, REF_EVALUATE_ABS_FRAC

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nrc():
    import vknrc_b200
    assert torch.cuda.is_available(), "these tests need a GPU"
    vknrc_b200.lib()
    return vknrc_b200


@pytest.fixture(scope="module")
def g2():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "nrc_golden_v2.npz")))


@pytest.fixture()
def state(nrc, g2):
    st = nrc.NrcState(0, (48, 32), seed=3)
    st.set_weights(g2["weights_fp32"])
    yield st
    st.close()


@pytest.fixture(scope="module")
def dscene(nrc, g2, oracle_mod):
    sc = golden_scene(g2, oracle_mod)
    return nrc.DeviceScene(sc.vertices, sc.vertex_indices, sc.texcoords, sc.texcoord_indices, sc.materials, sc.material_ids, sc.transforms,
                           sc.textures)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_A5_fused_encoder_vs_reference_shader(nrc, g2):
    """NRCInputEncode (NRCRecord.glsl:78-95) as the reference's source computes it: the frequency features and pass-through
    slots bit for bit, the one-blob features within one fp16 ulp (the shader source carries no `precise`: contraction, and
    with it the last bit, is the GLSL compiler's choice), the roughness slots within exp()'s last ulp on top."""
    rec, ref = g2["records14"], g2["glsl_encoded"].astype(np.float32)
    n = rec.shape[0]
    st = nrc.NrcState(0, (48, 32), seed=3)
    w = np.zeros(nrc.WEIGHT_COUNT, np.float32)
    for l in range(5):
        w[l * 4096:(l + 1) * 4096] = np.eye(64, dtype=np.float32).reshape(-1)
    for sign in (1.0, -1.0):
        w[0:4096] = sign * np.eye(64, dtype=np.float32).reshape(-1)
        got = np.zeros((n, 64), np.float32)
        for base in range(0, 64, 3):
            w5 = np.zeros((3, 64), np.float32)
            for c in range(3):
                if base + c < 64:
                    w5[c, base + c] = 1.0
            w[20480:] = w5.reshape(-1)
            st.set_weights(w)
            y = st.infer_unpacked(dev(rec)).float().cpu().numpy()
            for c in range(3):
                if base + c < 64:
                    got[:, base + c] = y[:, c]
        want = np.maximum(sign * ref, 0)
        exact = [i for i in range(64) if not 36 <= i < 56]
        assert np.array_equal(got[:, exact], want[:, exact])
        ulp16 = np.spacing(want[:, 36:52].astype(np.float16)).astype(np.float32)
        assert (np.abs(got[:, 36:52] - want[:, 36:52]) <= np.maximum(ulp16, 2.5e-7)).all()
        assert np.abs(got[:, 52:56] - want[:, 52:56]).max() <= 2.0 ** -11
    st.close()


def test_A5_standalone_encoder_vs_reference_shader(nrc, g2):
    """nrc_encode_inputs = NRCInputEncode as a kernel of its own: every feature, both signs, read directly (no network in
    between). Same bounds as the fused encoder above - it is the same device function."""
    ref = g2["glsl_encoded"].astype(np.float32)
    got = nrc.encode_inputs(dev(g2["records14"])).float().cpu().numpy()
    exact = [i for i in range(64) if not 36 <= i < 56]
    assert np.array_equal(got[:, exact], ref[:, exact])
    ulp16 = np.spacing(np.abs(ref[:, 36:52]).astype(np.float16)).astype(np.float32)
    assert (np.abs(got[:, 36:52] - ref[:, 36:52]) <= np.maximum(ulp16, 2.5e-7)).all()
    assert np.abs(got[:, 52:56] - ref[:, 52:56]).max() <= 2.0 ** -11


@pytest.mark.parametrize("n", [1, 127, 129, 1000])
def test_standalone_encoder_is_the_fused_encoder(nrc, g2, dscene, n):
    """Encoding on its own and then the pre-encoded path == the fused record paths, bit for bit (ragged tile counts, strided
    records); UnpackNRCInput + encode from packed words == nrc_unpack_inputs followed by nrc_encode_inputs, bit for bit."""
    st = nrc.NrcState(0, (48, 32), seed=11)
    rng = np.random.default_rng(n)
    rec = np.concatenate([rng.uniform(-4, 4, (n, 3)), rng.uniform(0, 1, (n, 11))], axis=1).astype(np.float32)
    enc = nrc.encode_inputs(dev(rec))
    assert enc.shape == (n, 64) and enc.dtype == torch.float16
    assert torch.equal(st.infer_encoded(enc, clamp=True), st.infer_unpacked(dev(rec)))  # (the record paths clamp, nrc_inference.comp:46)
    wide = np.zeros((n, 16), np.float32)  # 64-byte stride
    wide[:, :14] = rec
    assert torch.equal(nrc.encode_inputs(dev(wide), stride_bytes=64, n=n), enc)
    m = min(n, g2["packed_inputs"].shape[0])
    packed = dev(g2["packed_inputs"][:m])
    a = nrc.encode_packed_inputs(packed, dscene)
    b = nrc.encode_inputs(nrc.unpack_inputs(packed, dscene))
    assert torch.equal(a, b)
    # the same inputs inside 20-byte NRCEvalRecords (dst word in front, NRCRecord.glsl:11-14): stride 20, offset 4
    ev = np.zeros((m, 5), np.uint32)
    ev[:, 0] = 0xFFFFFFFF
    ev[:, 1:] = g2["packed_inputs"][:m].view(np.uint32).reshape(m, 4)
    d_ev = dev(ev.view(np.uint8).reshape(-1))
    assert torch.equal(nrc.encode_packed_inputs(d_ev[4:], dscene, stride_bytes=20, n=m), a)
    st.close()


def test_A1_A4_unpack_vs_reference_shader(nrc, g2, dscene):
    """UnpackNRCInput (NRCRecord.glsl:98-125): unorm16 decodes, barycentric interpolation through the index and transform
    buffers, face normal -> spherical, material / texture fetch. Tolerances: fp32 ulps of the quantity's range (positions
    ~4, the rest [0,1]); texture colours 2e-5 (bilinear blend associated differently, sRGB table vs pow)."""
    got = nrc.unpack_inputs(dev(g2["packed_inputs"]), dscene).cpu().numpy()
    ref = g2["glsl_unpacked"]
    assert np.abs(got[:, 0:3] - ref[:, 0:3]).max() <= 4e-6
    assert np.abs(got[:, 3:5] - ref[:, 3:5]).max() <= 6e-8
    assert np.abs(got[:, 5:7] - ref[:, 5:7]).max() <= 4e-6 and np.array_equal(got[:, 7], ref[:, 7])
    assert np.abs(got[:, 8:14] - ref[:, 8:14]).max() <= 2e-5


def test_A7_A9_forward_vs_reference_shader(nrc, g2):
    """test/evaluate_NV.comp (fp16 cooperative-matrix accumulators) vs the tcgen05 path (fp32 TMEM accumulators): north_star's
    1e-2 relative, with the absolute floor that fp16 accumulation of the reference itself needs (tests/util.py)."""
    w16 = dev(g2["weights_fp32"].astype(np.float16))
    y = nrc.mlp_evaluate_encoded(w16, dev(g2["inputs"])).float().cpu().numpy()
    assert out_err(y, g2["glsl_evaluate_nv"].astype(np.float32), abs_frac=REF_EVALUATE_ABS_FRAC) <= 1.0


def test_A10_A12_l2_gradient_vs_reference_shader(nrc, g2):
    """test/train_NV.comp: NNLoadDA3_L2Loss + NNBackwardDA*_ReLU + NNUpdateDW* of the reference (fp16 accumulators, fp16 per-warp dW
    partials) vs fp32 accumulation here. The two precisions of the SAME network differ by up to 8e-3 of a layer's max |dW| on
    this fixture (oracle fp16-mode vs fp32-mode, tests/test_ref_glsl.py pins the former to the shader); 2e-2 is the bound."""
    w16 = dev(g2["weights_fp32"].astype(np.float16))
    dw = torch.zeros(nrc.WEIGHT_COUNT, device="cuda")
    nrc.mlp_gradient_encoded(w16, dw, dev(g2["inputs"]), dev(g2["targets"]))
    errs = layer_rel_err(dw.cpu().numpy(), g2["glsl_train_nv_dw"])
    assert max(errs) <= 2e-2, errs


def test_A14_record_gradient_vs_reference_shader(nrc, g2, state, dscene):
    """nrc_gradient.comp on 40-byte NRCTrainRecords + scene: gather, encode, forward, relative-L2-luminance loss, backward, dW;
    500 of 512 records valid (tail lanes contribute exactly zero, Q11)."""
    tr = dev(np.ascontiguousarray(g2["train_records"]).reshape(-1))
    cnt = torch.tensor([int(g2["train_count"])], dtype=torch.int32, device="cuda")
    state.gradient(tr, dscene, count=cnt, max_count=512)
    d = state.download()
    assert d["gradients"][nrc.GRAD_COUNT_SLOT] == 500
    errs = layer_rel_err(d["gradients"][:nrc.WEIGHT_COUNT], g2["glsl_nrc_gradient_dw"])
    assert max(errs) <= 3e-2, errs  # fp16 vs fp32 accumulation (see above) + the gather's ulps through the top octaves


def test_A15_A16_optimizer_vs_reference_shaders_bit_exact(nrc, g2, state):
    """nrc_train_prepare.comp + nrc_optimize.comp (both variants) driven with the fixture's gradients: an over-full batch
    (clamped to 16384), an empty one (complete no-op), NaN / Inf gradients, EMA on and off, use_weights written or not -
    optimizer entries, fp16 weights, use_weights and the running products must equal the shaders' bit for bit."""
    gt = state.gradient_tensor()
    for i, (cnt, wu, ema) in enumerate(g2["opt_steps"]):
        g = np.zeros(nrc.GRADIENT_FLOATS, np.float32)
        g[:nrc.WEIGHT_COUNT] = g2["opt_gradient"] * np.float32(i + 1)
        g[nrc.GRAD_COUNT_SLOT] = min(int(cnt), nrc.TRAIN_BATCH_SIZE)  # nrc_train_prepare.comp:17-19 (done by the training kernel itself)
        gt.copy_(dev(g))
        state.set_use_ema_weights(bool(ema))
        state.adam_step(write_use_weights=bool(wu))
        s = state.download()["optimizer_state"]
        assert [int(s["t"]), float(s["beta1_t"]), float(s["beta2_t"]), float(s["alpha_t"]), float(s["alpha_t_1"])] == g2["glsl_opt_states"][i].tolist()
    d = state.download()
    assert np.array_equal(d["optimizer_entries"].view(np.uint32).reshape(-1, 4), g2["glsl_opt_entries"].view(np.uint32))
    assert np.array_equal(d["weights"].view(np.uint16), g2["glsl_opt_weights"])
    assert np.array_equal(d["use_weights"].view(np.uint16), g2["glsl_opt_use_weights"])


def test_A13_inference_pass_vs_reference_shader(nrc, g2, state, dscene):
    """nrc_inference.comp on a path-structured frame: which pixels / train records are written is bit-exact (dst codec,
    l..r ranges, invalid records skipped, alpha zeroed by the image store); the composited values follow the network tolerance."""
    ev = np.ascontiguousarray(g2["frame_eval_records"])
    n_ev = ev.shape[0]
    d_ev = dev(ev.reshape(-1))
    d_bf, d_gb = dev(g2["frame_bias_factor_r"]), dev(g2["frame_factor_gb"])
    d_tr = [dev(np.ascontiguousarray(g2[f"frame_train_records{b}"]).reshape(-1)) for b in range(4)]
    cnt = torch.tensor([n_ev], dtype=torch.int32, device="cuda")
    state.infer(d_ev, cnt, dscene, d_bf, d_gb, 48, d_tr, max_count=n_ev)
    bf, ref_bf = d_bf.cpu().numpy(), g2["glsl_frame_bias_factor_r"]
    scale = float(np.abs(ref_bf[..., :3] - g2["frame_bias_factor_r"][..., :3]).max())  # ~ the largest composited prediction
    assert np.array_equal(bf[..., 3], ref_bf[..., 3])
    assert np.abs(bf - ref_bf).max() <= 1e-2 * scale + 2e-3
    for b in range(4):
        before = np.ascontiguousarray(g2[f"frame_train_records{b}"]).view(np.float32).reshape(-1, 10)
        ref = np.ascontiguousarray(g2[f"glsl_frame_train_records{b}"]).view(np.float32).reshape(-1, 10)
        got = d_tr[b].cpu().numpy().view(np.float32).reshape(-1, 10)
        assert np.array_equal(got[:, 3:].view(np.uint32), ref[:, 3:].view(np.uint32))
        assert np.array_equal((got[:, :3] != before[:, :3]).any(axis=1), (ref[:, :3] != before[:, :3]).any(axis=1))
        assert np.abs(got[:, :3] - ref[:, :3]).max() <= 1e-2 * max(scale, float(np.abs(ref[:, :3]).max()))


def test_A20_learn_an_image_vs_reference_shaders(nrc, g2, state):
    """test/mlp_learning_an_image: gradient.comp (pcg2d uv stream, clamp-to-edge bilinear target, one-blob-32, L2) and
    inference.comp (640 x 640 grid -> rgba8)."""
    sx, sy = (int(v) for v in g2["image_seed"])
    state.image_train_step(dev(g2["image_rgba8"]), sx, sy, batch=512, lr=0.0)  # lr 0: leave the weights, read the gradient
    d = state.download()
    errs = layer_rel_err(d["gradients"][:nrc.WEIGHT_COUNT], g2["glsl_image_gradient_dw"])
    assert max(errs) <= 2e-2, errs
    img = state.image_infer(640).cpu().numpy()[::16, ::16]
    ref = g2["glsl_image_inference_16"]
    assert np.array_equal(img[..., 3], ref[..., 3]) and np.abs(img.astype(np.int32) - ref.astype(np.int32)).max() <= 3  # fp16 vs fp32 accumulate, in 1/255 steps

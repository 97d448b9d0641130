"""CPU tests: the oracle against the reference's own CPU code, the golden fixtures, and independent restatements."""
import numpy as np
import pytest

from util import REF_EVALUATE_ABS_FRAC, he_weights, layer_rel_err, out_err, random_records


def test_oracle_reproduces_golden_bit_exactly(golden, oracle_mod):
    o = oracle_mod
    x = golden["inputs"]
    for tag in ("he", "uniform"):
        w16 = golden[f"weights_{tag}_fp32"].astype(np.float16)
        assert np.array_equal(o.evaluate(w16, x, o.ACC_FP32).view(np.uint16), golden[f"oracle_evaluate_fp32acc_{tag}"].view(np.uint16))
        assert np.array_equal(o.evaluate(w16, x, o.ACC_FP16_CHUNK16).view(np.uint16),
                              golden[f"oracle_evaluate_fp16acc_{tag}"].view(np.uint16))
    w16 = golden["weights_he_fp32"].astype(np.float16)
    t = golden["targets"].astype(np.float32)
    assert np.array_equal(o.gradient(w16, x, t, o.LOSS_L2, 1.0, o.ACC_FP32), golden["oracle_dw_l2_fp32acc_he"])
    assert np.array_equal(o.encode(golden["records14"]).view(np.uint16), golden["oracle_encoded_records"].view(np.uint16))


def test_oracle_forward_matches_reference_evaluate_golden(golden):
    """The fixture holds outputs of the reference's own `Evaluate` (test/main.cpp:11-27)."""
    ref = golden["ref_evaluate_he"].astype(np.float32)
    for key in ("oracle_evaluate_fp32acc_he", "oracle_evaluate_fp16acc_he"):
        assert out_err(golden[key].astype(np.float32), ref, REF_EVALUATE_ABS_FRAC) <= 1.0, key
    # the two oracle modes themselves (fp16 accumulator per 16-wide MMA vs fp32) agree to the tight tolerance
    assert out_err(golden["oracle_evaluate_fp16acc_he"].astype(np.float32), golden["oracle_evaluate_fp32acc_he"].astype(np.float32)) <= 1.0


def _exact_forward(w16, x16):
    """The network evaluated in float64 with no intermediate rounding (fp16 weights and inputs)."""
    a = x16.astype(np.float64)
    w = w16.astype(np.float64)
    for l in range(5):
        a = np.maximum(a @ w[l * 4096:(l + 1) * 4096].reshape(64, 64).T, 0)
    return a @ w[20480:].reshape(3, 64).T


def test_fp32_accumulate_is_closer_to_exact_than_the_reference_cpu_path(golden):
    """Why parity against `Evaluate` carries a wider floor: its fp16 accumulation is the larger error source."""
    w16 = golden["weights_he_fp32"].astype(np.float16)
    exact = _exact_forward(w16, golden["inputs"])
    scale = np.abs(exact).max()
    e_ref = np.abs(golden["ref_evaluate_he"].astype(np.float64) - exact).max() / scale
    e_ours = np.abs(golden["oracle_evaluate_fp32acc_he"].astype(np.float64) - exact).max() / scale
    assert e_ours < e_ref and e_ours < 2e-3 and e_ref < 1e-2


def test_oracle_forward_matches_reference_evaluate_live(oracle_mod):
    o = oracle_mod
    if not o.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(5)
    for n in (128, 1024):
        w16 = he_weights(100 + n).astype(np.float16)
        x = rng.uniform(0, 1, (n, 64)).astype(np.float16)
        ref = o.ref_evaluate(w16, x).astype(np.float32)
        assert out_err(o.evaluate(w16, x, o.ACC_FP32).astype(np.float32), ref, REF_EVALUATE_ABS_FRAC) <= 1.0
        assert out_err(o.evaluate(w16, x, o.ACC_FP16_CHUNK16).astype(np.float32), ref, REF_EVALUATE_ABS_FRAC) <= 1.0


def test_reference_train_layer5_cross_check(golden, oracle_mod):
    """the reference's CPU `Train` is a sketch; its layer-5 dW equals sum_n (relu(y_n) - t'_n) a5_n^T with the
    targets read channel-major and fp16 accumulation. Rebuilt here from the ORACLE's activations it must agree."""
    o = oracle_mod
    w16 = golden["weights_he_fp32"].astype(np.float16)
    x, t = golden["inputs"], golden["targets"]
    y, acts = o.forward(w16, x, o.ACC_FP32, want_acts=True)
    n = x.shape[0]
    t_cm = t.reshape(-1).astype(np.float32).reshape(3, n).T  # row-major 3 x n map, as test/main.cpp:55-56 reads it
    dat = (np.maximum(y, 0).astype(np.float16).astype(np.float32) - t_cm)  # [n,3]
    dw5 = dat.T @ acts[5].astype(np.float32)  # [3,64]
    ref5 = golden["ref_train_he"][20480:20672].reshape(3, 64)
    assert np.abs(dw5 - ref5).max() <= 3e-2 * np.abs(ref5).max()  # the reference accumulates all n samples in fp16


def _encode_f64(rec):
    """Independent float64 restatement of NRCInputEncode (NRCRecord.glsl:47-95)."""
    rec = np.asarray(rec, np.float64)
    def tri(x):
        return 2.0 * np.abs(np.mod(x - 0.5, 2.0) - 1.0) - 1.0
    def cdf(x, ir):
        u = x * ir
        return np.clip(15.0 / 16.0 * u * (1 - 2.0 / 3.0 * u**2 + 0.2 * u**4) + 0.5, 0, 1)
    def ob4(x):
        l = np.array([0, .25, .5, .75]); r = l + .25
        return cdf(r[None] - x[:, None], 4) - cdf(l[None] - x[:, None], 4)
    cols = []
    for a in range(3):
        cols.append(tri(rec[:, a:a + 1] * (2.0 ** np.arange(12))[None]))
    for c in (3, 4, 5, 6):
        cols.append(ob4(rec[:, c]))
    cols.append(ob4(1 - np.exp(-rec[:, 7])))
    cols.append(rec[:, 8:14])
    cols.append(np.ones((rec.shape[0], 2)))
    return np.concatenate(cols, axis=1)


def test_encode_against_float64_restatement(oracle_mod):
    rec = random_records(11, 2000, pos_scale=1.0)  # |p| <= 1 keeps the fp32 argument reduction of 2048*p accurate
    got = oracle_mod.encode(rec).astype(np.float64)
    ref = _encode_f64(rec)
    assert got.shape == (2000, 64)
    # fp16 rounding (2^-11 relative) + fp32 evaluation error amplified by up to 2048 in the top octave
    assert np.abs(got - ref).max() <= 2e-3
    assert np.array_equal(got[:, 62:], np.ones((2000, 2)))
    assert np.array_equal(got[:, 56:62], rec[:, 8:14].astype(np.float16).astype(np.float64))


def test_encode_known_answers(oracle_mod):
    rec = np.zeros((1, 14), np.float32)
    e = oracle_mod.encode(rec)[0].astype(np.float32)
    assert np.all(e[0:36] == 0.0)          # tri(0) = 2|mod(-0.5, 2) - 1| - 1 = 2*0.5 - 1 = 0  (a triangle-wave sin(pi x))
    # one-blob of x = 0: bin 0 = cdf(1) - cdf(0) = 0.5, the rest 0
    for base in (36, 40, 44, 48, 52):
        assert e[base] == 0.5 and np.all(e[base + 1:base + 4] == 0.0)
    rec[0, 0] = 0.5                        # tri(0.5) = 2|mod(0, 2) - 1| - 1 = 1
    rec[0, 1] = 0.25
    e = oracle_mod.encode(rec)[0].astype(np.float32)
    assert e[0] == 1.0 and e[1] == 0.0 and e[12] == np.float16(0.5)  # tri(.5)=1, tri(1)=0, tri(.25)=.5
    f = lambda x: 2 * abs(((x - 0.5) % 2.0) - 1.0) - 1.0
    for k in range(12):
        assert e[k] == np.float16(f(0.5 * 2**k))
        assert e[12 + k] == np.float16(f(0.25 * 2**k))


def test_oneblob32_and_pcg(oracle_mod):
    o = oracle_mod
    enc = o.encode_oneblob32(np.array([[0.3, 0.9]], np.float32))[0].astype(np.float64)
    def cdf(x, ir):
        u = x * ir
        return min(max(15 / 16 * u * (1 - 2 / 3 * u * u + 0.2 * u**4) + 0.5, 0), 1)
    for i in range(32):
        assert abs(enc[i] - (cdf((i + 1) / 32 - 0.3, 32) - cdf(i / 32 - 0.3, 4))) < 1e-3
        assert abs(enc[32 + i] - (cdf((i + 1) / 32 - 0.9, 32) - cdf(i / 32 - 0.9, 4))) < 1e-3
    uv = o.learn_image_uv(123, 456, 300)
    assert uv.min() >= 0.0 and uv.max() <= 1.0 and len(np.unique(uv[:, 0])) > 290
    # pcg2d is pure integer arithmetic: pin one value computed by hand-evaluating gradient.comp:15-24
    def pcg(x, y):
        M = 0xFFFFFFFF
        x = (x * 1664525 + 1013904223) & M; y = (y * 1664525 + 1013904223) & M
        x = (x + y * 1664525) & M; y = (y + x * 1664525) & M
        x ^= x >> 16; y ^= y >> 16
        x = (x + y * 1664525) & M; y = (y + x * 1664525) & M
        x ^= x >> 16; y ^= y >> 16
        return x, y
    px, py = pcg(123 + 299 % 128, 456 + 299 // 128)
    assert uv[299, 0] == np.float32(np.float32(1.0) / np.float32(0xFFFFFFFF) * np.float32(px))
    assert uv[299, 1] == np.float32(np.float32(1.0) / np.float32(0xFFFFFFFF) * np.float32(py))


def test_dst_codec_bit_exact(oracle_mod):
    o = oracle_mod
    assert o.dst_screen(0, 0) == 0
    assert o.dst_screen(1919, 1079) == ((1919 | (1079 << 15)) << 1)
    assert o.dst_train(3, 16383, 16383) == (((3 | (16383 << 2) | (16383 << 16)) << 1) | 1)
    assert o.dst_train(3, 16383, 16383) != 0xFFFFFFFF
    rng = np.random.default_rng(1)
    for _ in range(200):
        x, y = int(rng.integers(0, 1 << 15)), int(rng.integers(0, 1 << 15))
        assert o.dst_decode(o.dst_screen(x, y)) == (0, x, y, 0)
        b, l, r = int(rng.integers(0, 4)), int(rng.integers(0, 1 << 14)), int(rng.integers(0, 1 << 14))
        assert o.dst_decode(o.dst_train(b, l, r)) == (1, b, l, r)


def test_backward_against_torch_autograd(oracle_mod):
    """Independent check of the restated backward pass: float64 autograd through the same network, with the same
    fp16 activations as straight-through values, must give the same dW up to the fp16 rounding of the deltas."""
    import torch
    o = oracle_mod
    n = 256
    rng = np.random.default_rng(9)
    w32 = he_weights(77)
    w16 = w32.astype(np.float16)
    x = rng.uniform(0, 1, (n, 64)).astype(np.float16)
    t = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    for loss in (o.LOSS_L2, o.LOSS_RELATIVE_L2_LUMINANCE):
        dw = o.gradient(w16, x, t, loss, 1.0, o.ACC_FP32)
        W = [torch.tensor(w16[l * 4096:(l + 1) * 4096].astype(np.float64).reshape(64, 64), requires_grad=True) for l in range(5)]
        W.append(torch.tensor(w16[20480:].astype(np.float64).reshape(3, 64), requires_grad=True))
        a = torch.tensor(x.astype(np.float64))
        r16 = lambda v: v + (v.detach().to(torch.float16).to(torch.float64) - v.detach())  # fp16 rounding, straight-through
        for l in range(5):
            a = r16(torch.relu(a @ W[l].T))
        y = r16(a @ W[5].T)
        tt = torch.tensor(t.astype(np.float64))
        if loss == o.LOSS_L2:
            L = ((y - tt) ** 2).sum()
        else:
            lum = (torch.clamp(y, min=0) * torch.tensor([0.299, 0.587, 0.114], dtype=torch.float64)).sum(1, keepdim=True)
            L = (((y - tt) ** 2) / (lum.detach() ** 2 + 0.01)).sum()  # denominator treated as constant (NN_nv.glsl:183-184)
        L.backward()
        ref = np.concatenate([w.grad.numpy().reshape(-1) for w in W])
        errs = layer_rel_err(dw, ref)
        assert max(errs) < 2e-2, (loss, errs)


def test_gradient_modes_agree_and_tail_is_exact(oracle_mod):
    o = oracle_mod
    rng = np.random.default_rng(3)
    w16 = he_weights(5).astype(np.float16)
    x = rng.uniform(0, 1, (200, 64)).astype(np.float16)
    t = rng.uniform(0, 1, (200, 3)).astype(np.float32)
    g32 = o.gradient(w16, x, t, o.LOSS_L2, 1.0, o.ACC_FP32)
    g16 = o.gradient(w16, x, t, o.LOSS_L2, 1.0, o.ACC_FP16_CHUNK16)
    assert max(layer_rel_err(g16, g32)) < 2e-2
    # zero input + zero target contributes exactly zero (nrc_gradient.comp:27-34)
    xz = np.concatenate([x, np.zeros((56, 64), np.float16)])
    tz = np.concatenate([t, np.zeros((56, 3), np.float32)])
    assert np.array_equal(o.gradient(w16, xz, tz, o.LOSS_L2, 1.0, o.ACC_FP32), g32)
    # loss scale is linear
    g2 = o.gradient(w16, x, t, o.LOSS_L2, 2.0, o.ACC_FP32)
    assert max(layer_rel_err(g2, 2 * g32)) < 2e-3


def test_optimizer_against_float64_adam(oracle_mod):
    o = oracle_mod
    w = he_weights(21)
    opt = o.Optimizer(w)
    rng = np.random.default_rng(2)
    m = np.zeros_like(w, np.float64); v = np.zeros_like(w, np.float64); wd = w.astype(np.float64); ema = wd.copy()
    b1t = b2t = at = 1.0
    for step in range(5):
        g = rng.standard_normal(o.WEIGHT_COUNT).astype(np.float32) * 100
        count = 1000 + step
        if step == 2:
            g[7] = np.nan; g[8] = np.inf  # guarded (nrc_optimize.comp:37-38)
        opt.step(g, count, write_use_weights=True, use_ema=True)
        gd = g.astype(np.float64) / count
        gd[~np.isfinite(gd)] = 0
        b1t *= 0.9; b2t *= 0.999; at_1 = at; at *= 0.99
        m = 0.9 * m + 0.1 * gd; v = 0.999 * v + 0.001 * gd * gd
        wd = wd - 0.002 * (m / (1 - b1t)) / (np.sqrt(v / (1 - b2t)) + 1e-8)
        ema = 0.01 / (1 - at) * wd + 0.99 * (1 - at_1) * ema
        assert np.abs(opt.entries["weight"] - wd).max() < 1e-5
        assert np.abs(opt.entries["ema_weight"] - ema).max() < 2e-5
    assert opt.state.t == 5 and abs(opt.state.beta1_t - 0.9**5) < 1e-6 and abs(opt.state.alpha_t_1 - 0.99**4) < 1e-6
    assert np.array_equal(opt.weights, opt.entries["weight"].astype(np.float16).view(np.uint16))
    assert np.array_equal(opt.use_weights, opt.entries["ema_weight"].astype(np.float16).view(np.uint16))
    # empty batch: nothing moves, running products do not advance (nrc_train_prepare.comp:22, nrc_optimize.comp:33-34)
    before = opt.entries.copy(); t0 = opt.state.t
    assert opt.step(np.ones(o.WEIGHT_COUNT, np.float32), 0, True, True) == 0
    assert np.array_equal(before, opt.entries) and opt.state.t == t0
    # over-full batches are clamped to 16384 (nrc_train_prepare.comp:18)
    assert opt.step(np.zeros(o.WEIGHT_COUNT, np.float32), 20000, False, False) == 16384


def test_scatter_semantics(oracle_mod):
    o = oracle_mod
    rng = np.random.default_rng(4)
    W, H = 16, 8
    bf = rng.uniform(0, 1, (H, W, 4)).astype(np.float32); gb = rng.uniform(0, 1, (H, W, 2)).astype(np.float32)
    recs = [rng.uniform(0, 1, (32, 10)).astype(np.float32) for _ in range(4)]
    bf0, recs0 = bf.copy(), [r.copy() for r in recs]
    pred = rng.uniform(0, 2, (4, 3)).astype(np.float32)
    dst = np.array([o.dst_screen(5, 3), o.dst_train(2, 4, 6), 0xFFFFFFFF, o.dst_screen(15, 7)], np.uint32)
    o.scatter(pred, dst, bf, gb, W, recs)
    exp = bf0.copy()
    for s, (x, y) in ((0, (5, 3)), (3, (15, 7))):
        f = np.array([bf0[y, x, 3], gb[y, x, 0], gb[y, x, 1]])
        exp[y, x, :3] = bf0[y, x, :3] + f * pred[s]; exp[y, x, 3] = 0
    assert np.allclose(bf, exp, rtol=1e-6) and np.array_equal(bf[0, 0], bf0[0, 0])
    for b in range(4):
        e = recs0[b].copy()
        if b == 2:
            e[4:7, :3] = e[4:7, :3] + e[4:7, 3:6] * pred[1]
        assert np.allclose(recs[b], e, rtol=1e-6)
        assert np.array_equal(recs[b][:, 3:], recs0[b][:, 3:])

"""GPU parity tests (run with -m gpu on a B200): the sm_100a kernels, called through the C ABI, against the CPU oracle,
the committed golden fixtures (outputs of the reference's own CPU code) and size-independent properties at the
BASELINE.json sizes. Tolerances are in tests/util.py."""
import numpy as np
import pytest

from util import (GRAD_REL_TOL, LOSS_CURVE_TOL, REF_EVALUATE_ABS_FRAC, he_weights, layer_rel_err, out_err, random_records)

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def nrc():
    import vknrc_b200
    assert torch.cuda.is_available(), "these tests need a GPU"
    vknrc_b200.lib()  # fails loudly if the CUDA library is missing: there is no fallback
    return vknrc_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def f32(t):
    return t.detach().cpu().numpy().astype(np.float32)


# ------------------------------------------------------------------------------------------------ config 1: raw MLP
@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 16384])
def test_evaluate_encoded_matches_oracle(nrc, oracle_mod, n):
    rng = np.random.default_rng(n)
    w16 = he_weights(n + 1).astype(np.float16)
    x = rng.uniform(0, 1, (n, 64)).astype(np.float16)
    y = nrc.mlp_evaluate_encoded(dev(w16), dev(x))
    ref = oracle_mod.evaluate(w16, x, oracle_mod.ACC_FP32).astype(np.float32)
    assert out_err(f32(y), ref) <= 1.0


def test_evaluate_encoded_matches_golden_reference_outputs(nrc, golden):
    """golden['ref_evaluate_*'] are outputs of the reference's own `Evaluate` (test/main.cpp:11-27)."""
    x = golden["inputs"]
    w16 = golden["weights_he_fp32"].astype(np.float16)
    y = f32(nrc.mlp_evaluate_encoded(dev(w16), dev(x)))
    assert out_err(y, golden["ref_evaluate_he"].astype(np.float32), REF_EVALUATE_ABS_FRAC) <= 1.0
    assert out_err(y, golden["oracle_evaluate_fp32acc_he"].astype(np.float32)) <= 1.0
    # reference test distribution (weights U(-0.02, 0.02)): outputs are ~1e-7, everything must stay finite and tiny
    wu = golden["weights_uniform_fp32"].astype(np.float16)
    yu = f32(nrc.mlp_evaluate_encoded(dev(wu), dev(x)))
    assert np.isfinite(yu).all() and np.abs(yu - golden["ref_evaluate_uniform"].astype(np.float32)).max() < 1e-6


def test_infer_encoded_host_equals_device_path(nrc):
    """nrc_infer_encoded_host (host buffers, chunked copy / compute overlap) returns bit for bit what the device-buffer
    call returns, for sizes below, at and far above one chunk per tile, pinned and pageable host memory."""
    st = nrc.NrcState(0, (64, 48), seed=21)
    g = torch.Generator(device="cuda").manual_seed(3)
    for n, pinned in ((1, True), (127, False), (1025, True), (300000, True), (70001, False)):
        x = torch.rand((n, 64), device="cuda", generator=g).half()
        ref = st.infer_encoded(x, clamp=True)
        hx = x.cpu().pin_memory() if pinned else x.cpu()
        hy = torch.full((n, 3), -1.0, dtype=torch.float16)
        hy = hy.pin_memory() if pinned else hy
        st.infer_encoded_host(hx, hy, clamp=True)
        torch.cuda.synchronize()
        assert torch.equal(hy, ref.cpu())
    st.close()


def test_evaluate_empty_and_determinism(nrc):
    w = dev(he_weights(3).astype(np.float16))
    out = nrc.mlp_evaluate_encoded(w, torch.empty((0, 64), dtype=torch.float16, device="cuda"))
    assert out.shape == (0, 3)
    x = torch.rand((5000, 64), device="cuda").half()
    a = nrc.mlp_evaluate_encoded(w, x).clone()
    b = nrc.mlp_evaluate_encoded(w, x)
    assert torch.equal(a, b)


@pytest.mark.parametrize("n", [1, 128, 200, 1024, 16384])
def test_gradient_encoded_matches_oracle(nrc, oracle_mod, n):
    rng = np.random.default_rng(100 + n)
    w16 = he_weights(n + 7).astype(np.float16)
    x = rng.uniform(0, 1, (n, 64)).astype(np.float16)
    t = rng.uniform(0, 1, (n, 3)).astype(np.float16)
    dw = torch.zeros(nrc.WEIGHT_COUNT, dtype=torch.float32, device="cuda")
    nrc.mlp_gradient_encoded(dev(w16), dw, dev(x), dev(t))
    ref = oracle_mod.gradient(w16, x, t.astype(np.float32), oracle_mod.LOSS_L2, 1.0, oracle_mod.ACC_FP32)
    errs = layer_rel_err(f32(dw), ref)
    assert max(errs) <= GRAD_REL_TOL, errs
    # the shader-like rounding mode (fp16 accumulators / per-warp fp16 dW partials) is the reference's own precision:
    ref16 = oracle_mod.gradient(w16, x, t.astype(np.float32), oracle_mod.LOSS_L2, 1.0, oracle_mod.ACC_FP16_CHUNK16)
    assert max(layer_rel_err(f32(dw), ref16)) <= 5e-2


def test_multi_tile_encoded_gradient_pipeline(nrc, oracle_mod):
    """Sizes at which every CTA runs several tiles: the forward pass of tile r overlaps the backward pass of tile r - 1
    (two tiles in flight, rotating activation buffers, TMA input tiles landing in recycled buffers). The batch gradient is
    additive over any split of the records, so the one-launch result must equal the sum over single-tile-per-CTA chunks
    (each of those paths is checked against the oracle above) up to fp32 reassociation; it must be bit-reproducible; and a
    ragged size that leaves some CTAs one tile short must agree too."""
    g = torch.Generator(device="cuda").manual_seed(77)
    w = dev(he_weights(55).astype(np.float16))
    for n in (148 * 128 * 3, 148 * 128 * 5 + 128 * 37 + 19, 1 << 18):
        x = torch.rand((n, 64), device="cuda", generator=g).half()
        t = torch.rand((n, 3), device="cuda", generator=g).half()
        dw = torch.zeros(nrc.WEIGHT_COUNT, dtype=torch.float32, device="cuda")
        nrc.mlp_gradient_encoded(w, dw, x, t)
        again = torch.zeros_like(dw)
        nrc.mlp_gradient_encoded(w, again, x, t)
        assert torch.equal(dw, again), "batch reduction must be bit-reproducible run to run"
        parts = torch.zeros(nrc.WEIGHT_COUNT, dtype=torch.float64, device="cuda")
        for lo in range(0, n, 16384):
            d = torch.zeros(nrc.WEIGHT_COUNT, dtype=torch.float32, device="cuda")
            nrc.mlp_gradient_encoded(w, d, x[lo:lo + 16384].contiguous(), t[lo:lo + 16384].contiguous())
            parts += d.double()
        assert max(layer_rel_err(f32(dw), parts.cpu().numpy())) <= 1e-4


def test_gradient_encoded_accumulates_and_is_deterministic(nrc, golden):
    w, x, t = dev(golden["weights_he_fp32"].astype(np.float16)), dev(golden["inputs"]), dev(golden["targets"])
    dw = torch.zeros(nrc.WEIGHT_COUNT, dtype=torch.float32, device="cuda")
    nrc.mlp_gradient_encoded(w, dw, x, t)
    once = dw.clone()
    assert max(layer_rel_err(f32(once), golden["oracle_dw_l2_fp32acc_he"])) <= GRAD_REL_TOL
    nrc.mlp_gradient_encoded(w, dw, x, t)  # the reference's kernel atomically ADDS into uDWeights
    assert torch.equal(dw, once + once)
    dw2 = torch.zeros_like(dw)
    nrc.mlp_gradient_encoded(w, dw2, x, t)
    assert torch.equal(dw2, once), "batch reduction must be bit-reproducible run to run"


def test_reference_cpu_train_layer5_against_gpu(nrc, golden):
    """Only the layer-5 dW of the reference's CPU `Train` is meaningful: it is sum_n (relu(y)-t') a5^T,
    no factor 2, targets read channel-major. Feed the GPU kernel targets arranged the same way and compare."""
    x, t = golden["inputs"], golden["targets"]
    n = x.shape[0]
    w16 = golden["weights_he_fp32"].astype(np.float16)
    t_cm = np.ascontiguousarray(t.reshape(-1).reshape(3, n).T)
    dw = torch.zeros(nrc.WEIGHT_COUNT, dtype=torch.float32, device="cuda")
    nrc.mlp_gradient_encoded(dev(w16), dw, dev(x), dev(t_cm))
    y = f32(nrc.mlp_evaluate_encoded(dev(w16), dev(x)))
    ours5 = f32(dw)[20480:20672].reshape(3, 64) / 2.0
    ref5 = golden["ref_train_he"][20480:20672].reshape(3, 64)
    # rows whose outputs are positive for every sample are directly comparable (the reference applies ReLU to y)
    ok = [c for c in range(3) if (y[:, c] > 0).all()]
    for c in ok:
        assert np.abs(ours5[c] - ref5[c]).max() <= 3e-2 * np.abs(ref5[c]).max()


# ------------------------------------------------------------------------------------------------ NRC state paths
@pytest.fixture()
def state(nrc):
    st = nrc.NrcState(0, (64, 48), seed=11)
    yield st
    st.close()


def test_reset_mlp_buffers_is_seeded_he_normal(nrc, state):
    d = state.download()
    w = d["optimizer_entries"]["weight"]
    assert abs(w.std() - np.sqrt(2 / 64)) < 0.01 and abs(w.mean()) < 0.01
    assert np.array_equal(d["weights"].view(np.uint16), w.astype(np.float16).view(np.uint16))  # RNE fp16 copy
    assert np.array_equal(d["weights"].view(np.uint16), d["use_weights"].view(np.uint16))
    assert np.array_equal(d["optimizer_entries"]["ema_weight"], w) and not d["optimizer_entries"]["m"].any()
    s = d["optimizer_state"]
    assert (s["t"], s["beta1_t"], s["beta2_t"], s["alpha_t"]) == (0, 1.0, 1.0, 1.0)
    state.reset_mlp_buffers(11)
    assert np.array_equal(state.download()["optimizer_entries"]["weight"], w)
    state.reset_mlp_buffers(12)
    assert not np.array_equal(state.download()["optimizer_entries"]["weight"], w)


@pytest.mark.parametrize("n", [1, 130, 4097])
def test_infer_unpacked_matches_oracle(nrc, oracle_mod, state, n):
    w32 = he_weights(31)
    state.set_weights(w32)
    rec = random_records(n, n)
    y = f32(state.infer_unpacked(dev(rec)))
    enc = oracle_mod.encode(rec)
    ref = oracle_mod.evaluate(w32.astype(np.float16), enc, oracle_mod.ACC_FP32, clamp=True).astype(np.float32)
    assert (y >= 0).all()
    assert out_err(y, ref) <= 1.0


def test_fused_encoding_is_bit_exact(nrc, oracle_mod, state):
    """Push the encoded features through an identity-like network: W0 = I (exact in fp16), so a_1 = relu(x); comparing
    against relu(oracle encode) checks the fused encoder. Frequency features (0..35) and the pass-through slots (56..63)
    must match bit for bit. The one-blob features (36..55) are evaluated with FMAs in Horner form (GLSL leaves contraction
    to the compiler, so the reference's own last bit is not defined): they must be within one fp16 ulp of the unfused
    evaluation (2.5e-7 absolute near zero); the roughness slots additionally see exp()'s last ulp."""
    n = 1000
    rec = random_records(5, n)
    w = np.zeros(nrc.WEIGHT_COUNT, np.float32)
    for l in range(5):
        w[l * 4096:(l + 1) * 4096] = np.eye(64, dtype=np.float32).reshape(-1)
    enc = oracle_mod.encode(rec).astype(np.float32)
    for sign in (1.0, -1.0):  # relu(+x) exposes the positive features, relu(-x) the negative ones
        w[0:4096] = sign * np.eye(64, dtype=np.float32).reshape(-1)
        got = np.zeros((n, 64), np.float32)
        for base in range(0, 64, 3):  # read 3 features at a time through the 3-row output layer
            w5 = np.zeros((3, 64), np.float32)
            for c in range(3):
                if base + c < 64:
                    w5[c, base + c] = 1.0
            w[20480:] = w5.reshape(-1)
            state.set_weights(w)
            y = f32(state.infer_unpacked(dev(rec)))
            for c in range(3):
                if base + c < 64:
                    got[:, base + c] = y[:, c]
        ref = np.maximum(sign * enc, 0)
        exact = [i for i in range(64) if not 36 <= i < 56]
        assert np.array_equal(got[:, exact], ref[:, exact])
        ulp16 = np.spacing(ref[:, 36:52].astype(np.float16)).astype(np.float32)
        assert (np.abs(got[:, 36:52] - ref[:, 36:52]) <= np.maximum(ulp16, 2.5e-7)).all()
        assert np.abs(got[:, 52:56] - ref[:, 52:56]).max() <= 2.0 ** -11


def test_device_resident_count_limits_work(nrc, oracle_mod, state):
    w32 = he_weights(41)
    state.set_weights(w32)
    rec = random_records(6, 1000)
    out = torch.full((1000, 3), -7.0, dtype=torch.float16, device="cuda")
    count = torch.tensor([333], dtype=torch.int32, device="cuda")
    state.infer_unpacked(dev(rec), count=count, outputs=out)
    y = f32(out)
    assert (y[333:] == -7.0).all() and (y[:333] >= 0).all()
    ref = oracle_mod.evaluate(w32.astype(np.float16), oracle_mod.encode(rec[:333]), oracle_mod.ACC_FP32, clamp=True)
    assert out_err(y[:333], ref.astype(np.float32)) <= 1.0
    count.zero_()
    out.fill_(-7.0)
    state.infer_unpacked(dev(rec), count=count, outputs=out)
    assert (f32(out) == -7.0).all()


def test_scatter_indexing_is_bit_exact(nrc, oracle_mod, state):
    """nrc_inference.comp:48-73: screen composite and train-record feedback, addressed by the dst bit codes."""
    rng = np.random.default_rng(8)
    w32 = he_weights(51)
    state.set_weights(w32)
    W, H, n = 64, 48, 3000
    rec = random_records(9, n)
    perm = rng.permutation(W * H)[:n]
    dst = np.empty(n, np.uint32)
    bf = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
    gb = rng.uniform(0, 1, (H, W, 2)).astype(np.float32)
    trecs = [rng.uniform(0, 1, (16384, 10)).astype(np.float32) for _ in range(4)]
    ranges, used = [], [0, 0, 0, 0]
    for i in range(n):
        kind = i % 7
        if kind == 3:
            dst[i] = 0xFFFFFFFF
        elif kind == 5:
            b = int(rng.integers(0, 4)); ln = int(rng.integers(1, 6))
            l = used[b]; r = l + ln - 1; used[b] += ln  # disjoint ranges, as the path tracer emits them
            dst[i] = oracle_mod.dst_train(b, l, r)
        else:
            dst[i] = oracle_mod.dst_screen(int(perm[i] % W), int(perm[i] // W))
    enc = oracle_mod.encode(rec)
    pred_ref = np.maximum(oracle_mod.forward(w32.astype(np.float16), enc, oracle_mod.ACC_FP32), 0).astype(np.float32)
    # run on the GPU
    d_bf, d_gb = dev(bf), dev(gb)
    d_tr = [dev(t) for t in trecs]
    state.infer_scatter_unpacked(dev(dst), dev(rec), None, d_bf, d_gb, W, d_tr)
    g_bf, g_tr = f32(d_bf), [f32(t) for t in d_tr]
    # (1) indexing: exactly the addressed pixels / record rows changed, nothing else
    touched_px = np.zeros((H, W), bool)
    touched_rec = [np.zeros(16384, bool) for _ in range(4)]
    for i in range(n):
        t, a, b, c = oracle_mod.dst_decode(int(dst[i])) if dst[i] != 0xFFFFFFFF else (2, 0, 0, 0)
        if t == 0:
            touched_px[b, a] = True
        elif t == 1:
            touched_rec[a][b:c + 1] = True
    changed_px = (g_bf != bf).any(axis=2)
    assert np.array_equal(changed_px | (touched_px & ~changed_px), touched_px)  # no untouched pixel changed
    assert (g_bf[touched_px][:, 3] == 0).all()                                  # alpha cleared where written
    for b in range(4):
        assert np.array_equal(g_tr[b][~touched_rec[b]], trecs[b][~touched_rec[b]])
        assert np.array_equal(g_tr[b][:, 3:], trecs[b][:, 3:])                  # factors / packed input untouched
    # (2) values: against the oracle's scatter driven by the GPU's own predictions is exact up to fma contraction;
    #     against the oracle's predictions within the output tolerance
    exp_bf, exp_tr = bf.copy(), [t.copy() for t in trecs]
    oracle_mod.scatter(pred_ref, dst, exp_bf, gb, W, exp_tr)
    scale = np.abs(pred_ref).max()
    assert np.abs(g_bf - exp_bf).max() <= 1e-2 * scale
    for b in range(4):
        assert np.abs(g_tr[b] - exp_tr[b]).max() <= 1e-2 * scale


def test_gradient_unpacked_and_adam_match_oracle(nrc, oracle_mod, state):
    w32 = he_weights(61)
    state.set_weights(w32)
    n = 3000
    rec = random_records(10, n)
    tgt = np.random.default_rng(10).uniform(0, 1, (n, 3)).astype(np.float32)
    cap = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
    state.set_prediction_capture(cap)
    count = torch.tensor([n], dtype=torch.int32, device="cuda")
    state.gradient_unpacked(dev(rec), dev(tgt), count=count)
    state.set_prediction_capture(None)
    d = state.download()
    enc = oracle_mod.encode(rec)
    gref, yref = oracle_mod.gradient(w32.astype(np.float16), enc, tgt, oracle_mod.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle_mod.ACC_FP32,
                                     want_y=True)
    assert max(layer_rel_err(d["gradients"][:nrc.WEIGHT_COUNT], gref)) <= GRAD_REL_TOL
    assert d["gradients"][nrc.GRAD_COUNT_SLOT] == n
    assert out_err(f32(cap), yref) <= 1.0
    loss_ref = oracle_mod.loss_value(yref, tgt, oracle_mod.LOSS_RELATIVE_L2_LUMINANCE)
    assert abs(d["gradients"][nrc.GRAD_LOSS_SLOT] / n - loss_ref) <= LOSS_CURVE_TOL * loss_ref
    # Adam + EMA: given the SAME gradient the optimizer must agree with the oracle bit for bit
    for use_ema in (False, True):
        state.set_weights(w32)
        state.set_use_ema_weights(use_ema)
        state.gradient_unpacked(dev(rec), dev(tgt))
        g = state.download()["gradients"]
        opt = oracle_mod.Optimizer(w32)
        for step in range(3):
            state.adam_step(write_use_weights=(step == 2))
            opt.step(g[:nrc.WEIGHT_COUNT], n, write_use_weights=(step == 2), use_ema=use_ema)
        d = state.download()
        assert np.array_equal(d["optimizer_entries"].view(np.uint32), opt.entries.view(np.uint32))
        assert np.array_equal(d["weights"].view(np.uint16), opt.weights)
        assert np.array_equal(d["use_weights"].view(np.uint16), opt.use_weights)
        s = d["optimizer_state"]
        assert (s["t"], s["beta1_t"], s["beta2_t"], s["alpha_t"], s["alpha_t_1"]) == (
            opt.state.t, opt.state.beta1_t, opt.state.beta2_t, opt.state.alpha_t, opt.state.alpha_t_1)


def test_fused_train_batch_equals_split_path(nrc, state):
    """nrc_train_batch_unpacked (gradient, reduction and Adam in ONE launch) must equal nrc_gradient_unpacked followed by
    nrc_adam_step (the multi-GPU path, with the all-reduce in between) bit for bit."""
    w32 = he_weights(63)
    rec, tgt = dev(random_records(14, 5000)), torch.rand((5000, 3), device="cuda")
    results = []
    for fused in (True, False):
        state.set_weights(w32)
        state.set_use_ema_weights(True)
        for step in range(3):
            if fused:
                state.train_batch_unpacked(rec, tgt, write_use_weights=(step == 2))
            else:
                state.gradient_unpacked(rec, tgt)
                state.adam_step(write_use_weights=(step == 2))
        results.append(state.download())
    a, b = results
    assert np.array_equal(a["optimizer_entries"].view(np.uint32), b["optimizer_entries"].view(np.uint32))
    assert np.array_equal(a["weights"].view(np.uint16), b["weights"].view(np.uint16))
    assert np.array_equal(a["use_weights"].view(np.uint16), b["use_weights"].view(np.uint16))
    assert a["optimizer_state"] == b["optimizer_state"] and a["optimizer_state"]["t"] == 3
    assert np.array_equal(a["gradients"], b["gradients"])


def test_train_frame_equals_four_batches(nrc, state):
    """nrc_train_frame_unpacked (the frame's four dependent batches in ONE launch, NRCRenderGraph.cpp:57-70) must equal
    four nrc_train_batch_unpacked calls bit for bit - including an empty batch in the middle (skipped entirely,
    nrc_optimize.comp:33-34), an over-full one (clamped, nrc_train_prepare.comp:17-19) and use_weights written by the
    last batch only."""
    w32 = he_weights(67)
    nb = nrc.TRAIN_BATCH_SIZE
    recs = [dev(random_records(20 + b, nb)) for b in range(4)]
    tgts = [torch.rand((nb, 3), device="cuda", generator=torch.Generator(device="cuda").manual_seed(b)) for b in range(4)]
    counts_host = [nb, 0, 100000, 777]
    results = []
    for frame in (True, False):
        state.set_weights(w32)
        state.set_use_ema_weights(True)
        counts = [torch.tensor([c], dtype=torch.int32, device="cuda") for c in counts_host]
        for _ in range(2):  # two frames
            if frame:
                state.train_frame_unpacked(recs, tgts, counts)
            else:
                for b in range(4):
                    state.train_batch_unpacked(recs[b], tgts[b], count=counts[b], write_use_weights=(b == 3))
        results.append((state.download(), [int(c.item()) for c in counts]))
    (a, ca), (b, cb) = results
    assert ca == cb == [nb, 0, nb, 777]
    assert np.array_equal(a["optimizer_entries"].view(np.uint32), b["optimizer_entries"].view(np.uint32))
    assert np.array_equal(a["weights"].view(np.uint16), b["weights"].view(np.uint16))
    assert np.array_equal(a["use_weights"].view(np.uint16), b["use_weights"].view(np.uint16))
    assert a["optimizer_state"] == b["optimizer_state"] and a["optimizer_state"]["t"] == 6
    assert np.array_equal(a["gradients"], b["gradients"]) and a["gradients"][nrc.GRAD_COUNT_SLOT] == 777
    assert not np.array_equal(a["weights"].view(np.uint16), a["use_weights"].view(np.uint16))  # EMA weights were published


def test_empty_and_overfull_batches(nrc, oracle_mod, state):
    w32 = he_weights(71)
    state.set_weights(w32)
    before = state.download()
    rec, tgt = dev(random_records(1, 256)), torch.rand((256, 3), device="cuda")
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    state.train_batch_unpacked(rec, tgt, count=count)  # empty batch: a complete no-op (nrc_optimize.comp:33-34)
    after = state.download()
    assert np.array_equal(before["optimizer_entries"], after["optimizer_entries"]) and after["optimizer_state"]["t"] == 0
    assert not after["gradients"][:nrc.WEIGHT_COUNT].any()
    count.fill_(100000)  # over-full: clamped to the buffer capacity and written back (nrc_train_prepare.comp:17-19)
    state.train_batch_unpacked(rec, tgt, count=count, max_count=256)
    assert int(count.item()) == 256 and state.download()["gradients"][nrc.GRAD_COUNT_SLOT] == 256


def test_training_loss_curve_tracks_oracle(nrc, oracle_mod, state):
    """'loss curves within 2 % after N steps': 30 Adam steps on a fixed synthetic batch, GPU vs oracle-driven loop."""
    w32 = he_weights(81)
    n, steps = 2048, 30
    rec = random_records(12, n, pos_scale=1.0)
    rng = np.random.default_rng(12)
    tgt = (0.5 + 0.5 * np.sin(rec[:, :3] * 3)).astype(np.float32) * rng.uniform(0.8, 1.2, (n, 3)).astype(np.float32)
    enc = oracle_mod.encode(rec)
    state.set_weights(w32)
    d_rec, d_tgt = dev(rec), dev(tgt)
    gpu_losses = []
    for _ in range(steps):
        state.train_batch_unpacked(d_rec, d_tgt, write_use_weights=True)
        gpu_losses.append(state.download()["gradients"][nrc.GRAD_LOSS_SLOT] / n)
    opt = oracle_mod.Optimizer(w32)
    ora_losses = []
    for _ in range(steps):
        g, y = oracle_mod.gradient(opt.weights, enc, tgt, oracle_mod.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle_mod.ACC_FP32, want_y=True)
        ora_losses.append(oracle_mod.loss_value(y, tgt, oracle_mod.LOSS_RELATIVE_L2_LUMINANCE))
        opt.step(g, n, True, False)
    gpu_losses, ora_losses = np.array(gpu_losses), np.array(ora_losses)
    assert ora_losses[-1] < 0.7 * ora_losses[0], "the oracle itself must be learning"
    assert np.abs(gpu_losses - ora_losses).max() <= LOSS_CURVE_TOL * ora_losses.max()
    assert abs(gpu_losses[-1] - ora_losses[-1]) <= LOSS_CURVE_TOL * ora_losses[-1] + 1e-3 * ora_losses[0]


# ------------------------------------------------------------------------------------------------ config 2: learn an image
def _bilinear(img, u, v):
    h, w = img.shape[:2]
    x = np.float32(u) * np.float32(w) - np.float32(0.5); y = np.float32(v) * np.float32(h) - np.float32(0.5)
    fx, fy = np.floor(x), np.floor(y)
    tx, ty = (x - fx).astype(np.float32), (y - fy).astype(np.float32)
    x0 = np.clip(fx.astype(int), 0, w - 1); x1 = np.clip(fx.astype(int) + 1, 0, w - 1)
    y0 = np.clip(fy.astype(int), 0, h - 1); y1 = np.clip(fy.astype(int) + 1, 0, h - 1)
    c = lambda yy, xx: img[yy, xx, :3].astype(np.float32)
    out = ((1 - tx) * (1 - ty))[:, None] * c(y0, x0) + (tx * (1 - ty))[:, None] * c(y0, x1) + ((1 - tx) * ty)[:, None] * c(y1, x0) + \
        (tx * ty)[:, None] * c(y1, x1)
    return (out / 255.0).astype(np.float32)


def test_learn_an_image_matches_oracle_loop(nrc, oracle_mod, state):
    rng = np.random.default_rng(13)
    yy, xx = np.mgrid[0:48, 0:64]
    img = np.stack([128 + 100 * np.sin(xx / 7.0), 128 + 100 * np.cos(yy / 5.0), (xx * 4) % 256, np.full_like(xx, 255.0)], axis=2)
    img = np.clip(img, 0, 255).astype(np.uint8)
    w32 = he_weights(91)
    state.set_weights(w32)
    batch, steps = 2048, 8
    seeds = [(int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32))) for _ in range(steps)]
    d_img = dev(img)
    for sx, sy in seeds:
        state.image_train_step(d_img, sx, sy, batch=batch, lr=0.01)
    gpu_w = state.download()["optimizer_entries"]["weight"]
    fp_w = w32.copy(); w16 = w32.astype(np.float16).view(np.uint16).copy()
    for sx, sy in seeds:
        uv = oracle_mod.learn_image_uv(sx, sy, batch)
        enc = oracle_mod.encode_oneblob32(uv)
        tgt = _bilinear(img, uv[:, 0], uv[:, 1])
        g = oracle_mod.gradient(w16, enc, tgt, oracle_mod.LOSS_L2, 1.0, oracle_mod.ACC_FP32)
        oracle_mod.sgd(fp_w, g, w16, 0.01, float(batch))
    upd = np.abs(fp_w - w32).max()
    assert upd > 1e-4, "the oracle loop must have moved the weights"
    assert np.abs(gpu_w - fp_w).max() <= 2e-2 * upd
    # inference over the pixel grid -> rgba8 (inference.comp:33-53)
    out = state.image_infer(32).cpu().numpy()
    gx, gy = np.meshgrid((np.arange(32) + 0.5) / 32, (np.arange(32) + 0.5) / 32)
    enc = oracle_mod.encode_oneblob32(np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1).astype(np.float32))
    y = oracle_mod.forward(state.download()["weights"], enc, oracle_mod.ACC_FP32)
    ref = np.rint(np.clip(y, 0, 1) * 255).astype(np.int32).reshape(32, 32, 3)
    assert np.abs(out[..., :3].astype(np.int32) - ref).max() <= 2 and (out[..., 3] == 255).all()


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_inference_properties(nrc, oracle_mod):
    """BASELINE config 3 size (1920x1080 queries). Per-query independence gives two exact, size-independent properties:
    a permutation of the queries permutes the outputs bit for bit, and (no biases, ReLU) doubling the inputs doubles
    the outputs (exactly, except where an activation falls in the fp16 subnormal range). A random subset is also checked against the oracle."""
    n = 1920 * 1080
    g = torch.Generator(device="cuda").manual_seed(5)
    w16 = he_weights(5).astype(np.float16)
    w = dev(w16)
    x = (torch.rand((n, 64), device="cuda", generator=g) * 0.5).half()
    y = nrc.mlp_evaluate_encoded(w, x)
    perm = torch.randperm(n, device="cuda", generator=g)
    yp = nrc.mlp_evaluate_encoded(w, x[perm].contiguous())
    assert torch.equal(yp, y[perm])
    y2 = nrc.mlp_evaluate_encoded(w, (x * 2).contiguous())  # exact except where an activation is an fp16 subnormal
    assert (y2.float() - 2 * y.float()).abs().max().item() <= 1e-3 * y.float().abs().max().item()
    assert (y2.float() == 2 * y.float()).float().mean().item() > 0.99
    idx = torch.randint(0, n, (4096,), device="cuda", generator=g)
    ref = oracle_mod.evaluate(w16, x[idx].cpu().numpy(), oracle_mod.ACC_FP32).astype(np.float32)
    assert out_err(f32(y[idx]), ref) <= 1.0


def test_full_size_training_properties(nrc, oracle_mod):
    """BASELINE config 4 size (4 x 16384 records): the batch gradient is additive over any split of the records
    (checked on the reduced fp32 gradient, tolerance = fp32 reassociation), the count slot is exact, and a record
    subset matches the oracle."""
    st = nrc.NrcState(0, (1920, 1080), seed=2)
    w32 = he_weights(15)
    st.set_weights(w32)
    n = 16384
    rec = random_records(21, 4 * n)
    tgt = np.random.default_rng(21).uniform(0, 1, (4 * n, 3)).astype(np.float32)
    d_rec, d_tgt = dev(rec), dev(tgt)
    grads = []
    for b in range(4):
        st.gradient_unpacked(d_rec[b * n:(b + 1) * n], d_tgt[b * n:(b + 1) * n])
        grads.append(st.download()["gradients"].copy())
        assert grads[-1][nrc.GRAD_COUNT_SLOT] == n
    st.gradient_unpacked(d_rec, d_tgt)  # all 65536 at once (several tiles per CTA -> TMEM accumulation across tiles)
    whole = st.download()["gradients"]
    assert whole[nrc.GRAD_COUNT_SLOT] == 4 * n
    parts = np.sum(np.stack(grads).astype(np.float64), axis=0)
    assert max(layer_rel_err(whole[:nrc.WEIGHT_COUNT], parts[:nrc.WEIGHT_COUNT])) <= 1e-4
    assert abs(whole[nrc.GRAD_LOSS_SLOT] - parts[nrc.GRAD_LOSS_SLOT]) <= 1e-4 * parts[nrc.GRAD_LOSS_SLOT]
    sub = slice(5 * 128, 5 * 128 + 1024)
    st.gradient_unpacked(d_rec[sub], d_tgt[sub])
    gref = oracle_mod.gradient(w32.astype(np.float16), oracle_mod.encode(rec[sub]), tgt[sub], oracle_mod.LOSS_RELATIVE_L2_LUMINANCE, 1.0,
                               oracle_mod.ACC_FP32)
    assert max(layer_rel_err(st.download()["gradients"][:nrc.WEIGHT_COUNT], gref)) <= GRAD_REL_TOL
    st.close()


def test_training_calls_replay_correctly_from_a_cuda_graph(nrc, state):
    """The library launches on the caller's stream and never synchronises, so a frame can be captured into a CUDA graph.
    Nothing that changes from launch to launch (grid-barrier base, optimizer step count) may be baked into the launch
    parameters: three replays of a captured nrc_train_frame must equal three direct calls bit for bit."""
    w32 = he_weights(71)
    n = 3000
    recs = [dev(random_records(80 + b, n)) for b in range(4)]
    tgts = [dev(np.random.default_rng(90 + b).uniform(0, 1, (n, 3)).astype(np.float32)) for b in range(4)]
    state.set_weights(w32)
    for _ in range(3):
        state.train_frame_unpacked(recs, tgts)
    direct = state.download()
    state.set_weights(w32)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    x = torch.rand((5000, 64), device="cuda").half()
    y = torch.zeros((5000, 3), device="cuda", dtype=torch.float16)
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            state.train_frame_unpacked(recs, tgts)
            state.infer_encoded(x, y, clamp=True)  # (launched with programmatic stream serialization: must capture too)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    state.set_weights(w32)  # (capture does not execute; start from the same state as the direct run)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    replayed = state.download()
    assert np.array_equal(direct["weights"].view(np.uint16), replayed["weights"].view(np.uint16))
    assert np.array_equal(direct["optimizer_entries"].view(np.uint32), replayed["optimizer_entries"].view(np.uint32))
    assert direct["optimizer_state"] == replayed["optimizer_state"] and replayed["optimizer_state"]["t"] == 12
    assert torch.equal(y, state.infer_encoded(x, clamp=True))  # the captured inference ran on the weights of the third replay


def test_infer_eval_records_host_equals_device_path(nrc):
    """nrc_infer_eval_records_host (20-byte NRCEvalRecords in host memory -> fp16x3 in host memory, chunked three-stream pipeline)
    must equal nrc_infer_packed on device copies of the same records bit for bit, also for a ragged size and a second call that
    reuses the staging buffers."""
    from util import make_scene
    from vknrc_b200 import synth
    sc = make_scene(33)
    dsc = nrc.DeviceScene(sc.vertices, sc.vertex_indices, sc.texcoords, sc.texcoord_indices, sc.materials, sc.material_ids, sc.transforms, sc.textures)
    st = nrc.NrcState(0, (64, 64), seed=9)
    for n in (70001, 5000):
        ev = synth.eval_records_screen(34, n, 1, sc.material_ids.shape[0], sc.transforms.shape[0])
        h_ev = torch.from_numpy(ev.view(np.uint8).reshape(-1)).pin_memory()
        h_out = torch.empty((n, 3), dtype=torch.float16).pin_memory()
        st.infer_eval_records_host(h_ev, dsc, h_out)
        torch.cuda.synchronize()
        d_ev = h_ev.cuda()
        ref = st.infer_packed(d_ev[4:], dsc, stride_bytes=20, max_count=n).cpu()
        assert torch.equal(h_out, ref)
    st.close()

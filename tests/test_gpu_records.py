"""GPU parity of the reference's own record formats (run with -m gpu): 16-byte PackedNRCInput inside 20-byte NRCEvalRecord
/ 40-byte NRCTrainRecord arrays, unpacked on the fly from scene buffers (UnpackNRCInput, shader/src/NRCRecord.glsl:98-125
over shader/src/Scene.glsl:8-71) - against the CPU restatement in oracle/. The reference ships no fixtures for this step
and its GLSL cannot run here, so this row is pinned by the restatement alone."""
import numpy as np
import pytest

from util import GRAD_REL_TOL, he_weights, layer_rel_err, make_scene, out_err, random_packed_inputs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def nrc():
    import vknrc_b200
    assert torch.cuda.is_available(), "these tests need a GPU"
    vknrc_b200.lib()
    return vknrc_b200


@pytest.fixture(scope="module")
def oracle_mod():
    import oracle
    return oracle


@pytest.fixture()
def state(nrc):
    st = nrc.NrcState(0, (64, 48), seed=3)
    yield st
    st.close()


def upload_scene(nrc, sc):
    return nrc.DeviceScene(sc.vertices, sc.vertex_indices, sc.texcoords, sc.texcoord_indices, sc.materials, sc.material_ids, sc.transforms,
                           sc.textures)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_unpack_matches_oracle(nrc, oracle_mod):
    sc = make_scene(5)
    dsc = upload_scene(nrc, sc)
    for n in (1, 255, 5000):
        pk = random_packed_inputs(6, n, sc)
        got = nrc.unpack_inputs(dev(pk), dsc).cpu().numpy()
        ref = oracle_mod.unpack(sc, pk)
        # integer decodes are exact; gathers + fp32 arithmetic differ by FMA contraction and libm ulps only
        assert np.array_equal(got[:, 3:5], ref[:, 3:5])                      # scattered dir: unorm16 decode, bit-exact
        assert np.array_equal(got[:, 7], ref[:, 7])                          # roughness: a plain gather
        assert np.abs(got[:, :3] - ref[:, :3]).max() <= 4e-6                 # position (|p| < 4)
        assert np.abs(got[:, 5:7] - ref[:, 5:7]).max() <= 2e-6               # spherical normal
        assert np.abs(got[:, 8:] - ref[:, 8:]).max() <= 2e-5                 # diffuse / specular incl. sRGB texture fetches
        # (a 1-ulp difference in uv is scaled by the texture size times the texel contrast; fp16 features resolve 5e-4)
    # strided source: the PackedNRCInput sits at byte 4 of a 20-byte NRCEvalRecord and byte 24 of a 40-byte NRCTrainRecord
    pk = random_packed_inputs(7, 300, sc)
    ev = np.zeros(300, nrc.EVAL_RECORD_DTYPE)
    ev["packed_input"] = np.ascontiguousarray(pk).view(nrc.EVAL_RECORD_DTYPE["packed_input"]).reshape(300)
    d_ev = dev(ev.view(np.uint8).reshape(-1))
    got = nrc.unpack_inputs(d_ev[4:], dsc, stride_bytes=20, n=300).cpu().numpy()
    assert np.abs(got - oracle_mod.unpack(sc, pk)).max() <= 2e-5


def test_prim_table_gather_is_bit_identical(nrc):
    """NrcScene::prim_table (the per-primitive 64-byte rows built by nrc_scene_build_prim_table) only changes where the
    gather reads its values from: the unpacked inputs must match the index-buffer gather bit for bit."""
    sc = make_scene(23)
    args = (sc.vertices, sc.vertex_indices, sc.texcoords, sc.texcoord_indices, sc.materials, sc.material_ids, sc.transforms, sc.textures)
    flat, indexed = nrc.DeviceScene(*args, prim_table=True), nrc.DeviceScene(*args, prim_table=False)
    assert flat.c.prim_table and not indexed.c.prim_table
    pk = dev(random_packed_inputs(29, 20000, sc))
    assert torch.equal(nrc.unpack_inputs(pk, flat), nrc.unpack_inputs(pk, indexed))


def test_infer_packed_matches_oracle(nrc, oracle_mod, state):
    sc = make_scene(11)
    dsc = upload_scene(nrc, sc)
    w32 = he_weights(12)
    state.set_weights(w32)
    n = 3001
    pk = random_packed_inputs(13, n, sc)
    y = state.infer_packed(dev(pk), dsc).float().cpu().numpy()
    ref = oracle_mod.evaluate(w32.astype(np.float16), oracle_mod.encode(oracle_mod.unpack(sc, pk)), oracle_mod.ACC_FP32, clamp=True)
    assert out_err(y, ref.astype(np.float32)) <= 1.0


def test_full_size_record_inference_properties(nrc, oracle_mod):
    """BASELINE config 3 size (1920x1080 queries) through the record input modes, where every CTA runs many tiles per slot and
    the producer warps run several tiles ahead (buffer hand-offs in both directions): queries are independent, so a
    permutation of the records permutes the outputs bit for bit; a random subset is checked against the oracle; the fused
    path equals (bit for bit) the stand-alone unpack followed by the 14-float record path."""
    n = 1920 * 1080
    st = nrc.NrcState(0, (1920, 1080), seed=4)
    w32 = he_weights(41)
    st.set_weights(w32)
    sc = make_scene(17)
    dsc = upload_scene(nrc, sc)
    g = torch.Generator(device="cuda").manual_seed(9)
    pk = dev(random_packed_inputs(19, 8192, sc).view(np.int32))  # (int32 view: torch indexes no uint32 tensors)
    pk = pk[torch.randint(0, 8192, (n,), device="cuda", generator=g)].contiguous()  # n records drawn from 8192 distinct ones
    y = st.infer_packed(pk, dsc)
    perm = torch.randperm(n, device="cuda", generator=g)
    assert torch.equal(st.infer_packed(pk[perm].contiguous(), dsc), y[perm])
    unp = nrc.unpack_inputs(pk, dsc)
    yu = st.infer_unpacked(unp)
    assert torch.equal(yu, y)
    assert torch.equal(st.infer_unpacked(unp[perm].contiguous()), yu[perm])
    # the encode stage on its own at the same size: rows permute with the records, both record forms give the same rows, and
    # the pre-encoded path on them is the fused path, bit for bit
    enc = nrc.encode_inputs(unp)
    assert torch.equal(nrc.encode_inputs(unp[perm].contiguous()), enc[perm])
    assert torch.equal(nrc.encode_packed_inputs(pk, dsc), enc)
    assert torch.equal(st.infer_encoded(enc, clamp=True), y)
    del enc
    idx = torch.randint(0, n, (2048,), device="cuda", generator=g)
    sub = pk[idx].cpu().numpy().view(np.uint32)
    ref = oracle_mod.evaluate(w32.astype(np.float16), oracle_mod.encode(oracle_mod.unpack(sc, sub)), oracle_mod.ACC_FP32, clamp=True)
    assert out_err(y[idx].float().cpu().numpy(), ref.astype(np.float32)) <= 1.0
    # ragged tail + device-resident count smaller than the buffer
    m = n - 12345
    cnt = torch.tensor([m], dtype=torch.int32, device="cuda")
    out = torch.full((n, 3), -1.0, dtype=torch.float16, device="cuda")
    st.infer_unpacked(unp, count=cnt, outputs=out)
    assert torch.equal(out[:m], yu[:m]) and (out[m:] == -1).all()
    st.close()


def test_nrc_infer_eval_records_scatter(nrc, oracle_mod, state):
    """The full nrc_inference.comp pass on NRCEvalRecord[]: device count, invalid / screen / train destinations."""
    sc = make_scene(21)
    dsc = upload_scene(nrc, sc)
    w32 = he_weights(22)
    state.set_weights(w32)
    W, H, n = 64, 48, 2500
    rng = np.random.default_rng(23)
    pk = random_packed_inputs(24, n, sc)
    dst = np.zeros(n, np.uint32)
    perm = rng.permutation(W * H)[:n]  # distinct pixels: the screen RMW has no ordering between duplicates
    kind = rng.integers(0, 10, n)
    tr_cursor = [0, 0, 0, 0]
    for i in range(n):
        if kind[i] == 0:
            dst[i] = 0xFFFFFFFF
        elif kind[i] <= 6:
            dst[i] = oracle_mod.dst_screen(int(perm[i] % W), int(perm[i] // W))
        else:  # disjoint [l, r] ranges per batch (as the path tracer emits them, path_tracer.comp:343-371)
            b = int(rng.integers(0, 4))
            ln = int(rng.integers(1, 5))
            if tr_cursor[b] + ln > 1024:
                dst[i] = 0xFFFFFFFF
                continue
            dst[i] = oracle_mod.dst_train(b, tr_cursor[b], tr_cursor[b] + ln - 1)
            tr_cursor[b] += ln
    ev = np.zeros(n + 50, nrc.EVAL_RECORD_DTYPE)  # 50 records past the device count must be ignored
    ev["dst"][:n] = dst
    ev["dst"][n:] = oracle_mod.dst_screen(0, 0)
    ev["packed_input"][:n] = np.ascontiguousarray(pk).view(nrc.EVAL_RECORD_DTYPE["packed_input"]).reshape(n)
    ev["packed_input"][n:] = ev["packed_input"][0]
    bf = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
    gb = rng.uniform(0, 1, (H, W, 2)).astype(np.float32)
    tr = [np.zeros(1024, nrc.TRAIN_RECORD_DTYPE) for _ in range(4)]
    for t in tr:
        t["bias"], t["factor"] = rng.uniform(0, 1, (1024, 3)), rng.uniform(0, 1, (1024, 3))
    d_bf, d_gb = dev(bf), dev(gb)
    d_tr = [dev(t.view(np.uint8).reshape(-1)) for t in tr]
    count = torch.tensor([n], dtype=torch.int32, device="cuda")
    state.infer(dev(ev.view(np.uint8).reshape(-1)), count, dsc, d_bf, d_gb, W, d_tr, max_count=n + 50)
    g_bf = d_bf.cpu().numpy()
    g_tr = [t.cpu().numpy().view(nrc.TRAIN_RECORD_DTYPE) for t in d_tr]
    # oracle: predictions through the restated unpack + encode + network, then the restated scatter
    pred = oracle_mod.evaluate(w32.astype(np.float16), oracle_mod.encode(oracle_mod.unpack(sc, pk)), oracle_mod.ACC_FP32, clamp=True).astype(np.float32)
    exp_bf = bf.copy()
    exp_tr = [t.copy().view(np.float32).reshape(-1, 10) for t in tr]  # 40-byte records as 10 floats: bias, factor, packed input
    oracle_mod.scatter(pred, dst, exp_bf, gb, W, exp_tr)
    scale = np.abs(pred).max()
    assert np.abs(g_bf - exp_bf).max() <= 1e-2 * scale
    untouched = np.ones(W * H, bool)
    for i in range(n):
        if dst[i] != 0xFFFFFFFF and (dst[i] & 1) == 0:
            untouched[perm[i]] = False
    assert np.array_equal(g_bf.reshape(-1, 4)[untouched], bf.reshape(-1, 4)[untouched])  # bit-exact indexing: nothing else moved
    for b in range(4):
        got = g_tr[b].view(np.float32).reshape(-1, 10)
        assert np.abs(got[:, :6] - exp_tr[b][:, :6]).max() <= 1e-2 * scale
        assert np.array_equal(got[tr_cursor[b]:].view(np.uint32), exp_tr[b][tr_cursor[b]:].view(np.uint32))  # rows past the ranges untouched
        assert np.array_equal(g_tr[b]["packed_input"], tr[b]["packed_input"])


def test_train_records_match_oracle_and_unpacked_path(nrc, oracle_mod, state):
    sc = make_scene(31)
    dsc = upload_scene(nrc, sc)
    w32 = he_weights(32)
    n = 4000
    pk = random_packed_inputs(33, n, sc)
    rng = np.random.default_rng(34)
    rec = np.zeros(n, nrc.TRAIN_RECORD_DTYPE)
    rec["bias"], rec["factor"] = rng.uniform(0, 1, (n, 3)), rng.uniform(0, 1, (n, 3))
    rec["packed_input"] = np.ascontiguousarray(pk).view(nrc.TRAIN_RECORD_DTYPE["packed_input"]).reshape(n)
    d_rec = dev(rec.view(np.uint8).reshape(-1))
    state.set_weights(w32)
    state.gradient(d_rec, dsc)
    g = state.download()["gradients"]
    # The gather itself is checked in test_unpack_matches_oracle. Its fp32 ulp differences (FMA contraction, libm) are
    # amplified 2048x by the top frequency octave, so the NETWORK is compared on the device's own unpacked values.
    unp = nrc.unpack_inputs(d_rec[24:], dsc, stride_bytes=40, n=n).cpu().numpy()
    assert np.abs(unp - oracle_mod.unpack(sc, pk)).max() <= 2e-5
    gref = oracle_mod.gradient(w32.astype(np.float16), oracle_mod.encode(unp), np.ascontiguousarray(rec["bias"]), oracle_mod.LOSS_RELATIVE_L2_LUMINANCE,
                               1.0, oracle_mod.ACC_FP32)
    assert max(layer_rel_err(g[:nrc.WEIGHT_COUNT], gref)) <= GRAD_REL_TOL
    assert g[nrc.GRAD_COUNT_SLOT] == n
    # the same batch through the 14-float entry point (inputs unpacked by the oracle): same network, same optimizer
    state.set_weights(w32)
    state.train_batch(d_rec, dsc, write_use_weights=True)
    a = state.download()
    state.set_weights(w32)
    state.train_batch_unpacked(dev(unp), dev(np.ascontiguousarray(rec["bias"])), write_use_weights=True)
    b = state.download()
    assert np.abs(a["weights"].astype(np.float32) - b["weights"].astype(np.float32)).max() <= 2e-3
    assert a["optimizer_state"] == b["optimizer_state"]


def test_train_frame_records_equals_four_batches(nrc, state):
    sc = make_scene(41)
    dsc = upload_scene(nrc, sc)
    w32 = he_weights(42)
    rng = np.random.default_rng(43)
    nb = 2048
    recs = []
    for b in range(4):
        r = np.zeros(nb, nrc.TRAIN_RECORD_DTYPE)
        r["bias"], r["factor"] = rng.uniform(0, 1, (nb, 3)), rng.uniform(0, 1, (nb, 3))
        r["packed_input"] = np.ascontiguousarray(random_packed_inputs(50 + b, nb, sc)).view(nrc.TRAIN_RECORD_DTYPE["packed_input"]).reshape(nb)
        recs.append(dev(r.view(np.uint8).reshape(-1)))
    res = []
    for frame in (True, False):
        state.set_weights(w32)
        counts = [torch.tensor([c], dtype=torch.int32, device="cuda") for c in (nb, 100, 0, 5000)]
        if frame:
            state.train_frame(recs, dsc, counts, max_count=nb)
        else:
            for b in range(4):
                state.train_batch(recs[b], dsc, count=counts[b], max_count=nb, write_use_weights=(b == 3))
        res.append((state.download(), [int(c.item()) for c in counts]))
    (a, ca), (b, cb) = res
    assert ca == cb == [nb, 100, 0, nb]
    for k in ("weights", "use_weights", "optimizer_entries", "gradients"):
        assert np.array_equal(np.ascontiguousarray(a[k]).view(np.uint8), np.ascontiguousarray(b[k]).view(np.uint8)), k
    assert a["optimizer_state"]["t"] == 3


def test_whole_frame_on_path_structured_records(nrc, oracle_mod):
    """A frame as the reference's render graph runs it (src/rg/NRCRenderGraph.cpp:46-80) on records with the structure its
    path tracer emits (vknrc_b200.synth.frame_records: screen queries for every pixel, train paths with suffix-scanned
    bias / factor appended contiguously to random batches, a tail query per unfinished path whose answer is fed back into
    that path's training targets, one batch overfull): nrc_infer (device count, use_weights) then nrc_train_frame (device
    counts, clamped in place) against the oracle's unpack -> encode -> network -> scatter -> 4 x (gradient, Adam/EMA)."""
    from vknrc_b200 import synth
    sc = make_scene(61)
    dsc = upload_scene(nrc, sc)
    W, H, cap = 192, 96, 1024
    n_prims, n_inst = sc.material_ids.shape[0], sc.transforms.shape[0]
    fr = synth.frame_records(62, W, H, n_prims, n_inst, train_probability=0.25, batch_size=cap)
    assert fr["train_counts"].max() > cap and fr["eval_count"] > W * H  # at least one batch overflows, tail queries exist
    st = nrc.NrcState(0, (W, H), seed=5)
    w32 = he_weights(63)
    st.set_weights(w32)
    rng = np.random.default_rng(64)
    bf = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
    gb = rng.uniform(0, 1, (H, W, 2)).astype(np.float32)
    ev = fr["eval_records"]
    n_ev = fr["eval_count"]
    d_ev = dev(ev.view(np.uint8).reshape(-1))
    d_bf, d_gb = dev(bf), dev(gb)
    d_tr = [dev(t.view(np.uint8).reshape(-1)) for t in fr["train_records"]]
    d_evc = torch.tensor([n_ev], dtype=torch.int32, device="cuda")
    d_trc = [torch.tensor([int(c)], dtype=torch.int32, device="cuda") for c in fr["train_counts"]]
    # ---- the frame
    st.infer(d_ev, d_evc, dsc, d_bf, d_gb, W, d_tr, max_count=n_ev)
    fed = [t.cpu().numpy().view(nrc.TRAIN_RECORD_DTYPE).copy() for t in d_tr]  # train records after the feedback
    st.train_frame(d_tr, dsc, d_trc, max_count=cap)
    got = st.download()
    assert [int(c.item()) for c in d_trc] == [min(int(c), cap) for c in fr["train_counts"]]  # nrc_train_prepare.comp:17-19
    # ---- oracle: inference + scatter (screen composite and feedback into the train targets)
    pk = np.ascontiguousarray(ev["packed_input"]).view(np.uint32).reshape(n_ev, 4)
    pred = oracle_mod.evaluate(w32.astype(np.float16), oracle_mod.encode(oracle_mod.unpack(sc, pk)), oracle_mod.ACC_FP32, clamp=True).astype(np.float32)
    exp_bf = bf.copy()
    exp_tr = [t.copy().view(np.float32).reshape(-1, 10) for t in fr["train_records"]]
    oracle_mod.scatter(pred, np.ascontiguousarray(ev["dst"]), exp_bf, gb, W, exp_tr)
    scale = np.abs(pred).max()
    assert np.abs(d_bf.cpu().numpy() - exp_bf).max() <= 1e-2 * scale
    for b in range(4):
        g = fed[b].view(np.float32).reshape(-1, 10)
        assert np.abs(g[:, :6] - exp_tr[b][:, :6]).max() <= 1e-2 * max(scale, np.abs(exp_tr[b][:, :3]).max())
        assert np.array_equal(g[:, 6:].view(np.uint32), exp_tr[b][:, 6:].view(np.uint32))
    # ---- oracle: the four training batches on the device's own fed-back targets and unpacked inputs (the gather is checked
    # separately; its fp32 ulps are amplified by the top frequency octave), Adam / EMA chained, use_weights from batch 3
    opt = oracle_mod.Optimizer(w32)
    for b in range(4):
        cnt = min(int(fr["train_counts"][b]), cap)
        unp = nrc.unpack_inputs(d_tr[b][24:], dsc, stride_bytes=40, n=cnt).cpu().numpy()
        tgt = np.ascontiguousarray(fed[b]["bias"][:cnt])
        grad = oracle_mod.gradient(opt.weights, oracle_mod.encode(unp), tgt, oracle_mod.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle_mod.ACC_FP32)
        opt.step(grad, cnt, b == 3, False, batch_cap=cap)
    # fp16 weights after four chained steps: the gradients agree to ~1e-5 of scale, Adam's first steps move every weight by
    # about lr regardless of the gradient's size, so a sign flip of a near-zero gradient shows up as 2 lr = 4e-3
    dw = np.abs(got["weights"].view(np.float16).astype(np.float32) - opt.weights.view(np.float16).astype(np.float32))
    assert np.median(dw) <= 1e-4 and (dw > 5e-3).mean() <= 2e-3 and dw.max() <= 2e-2, (np.median(dw), (dw > 5e-3).mean(), dw.max())
    s = got["optimizer_state"]
    assert (s["t"], s["beta1_t"], s["beta2_t"], s["alpha_t"], s["alpha_t_1"]) == (
        opt.state.t, opt.state.beta1_t, opt.state.beta2_t, opt.state.alpha_t, opt.state.alpha_t_1)
    st.close()

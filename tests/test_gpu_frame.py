"""GPU tests of the frame-level semantics (run with -m gpu): the render graph's schedule src/rg/NRCRenderGraph.cpp:46-80 and
:100-113 as nrc_frame_begin / nrc_frame, use_weights publication (Q9), and robustness of the one-launch training frame."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import he_weights, make_scene, random_packed_inputs, random_records

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def nrc():
    import vknrc_b200
    assert torch.cuda.is_available(), "these tests need a GPU"
    vknrc_b200.lib()
    return vknrc_b200


@pytest.fixture()
def state(nrc):
    st = nrc.NrcState(0, (64, 48), seed=3)
    yield st
    st.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def same(a, b, keys=("weights", "use_weights", "optimizer_entries", "gradients")):
    return all(np.array_equal(np.ascontiguousarray(a[k]).view(np.uint8), np.ascontiguousarray(b[k]).view(np.uint8)) for k in keys)


def test_q9_empty_last_batch_does_not_republish_use_weights(nrc, state):
    """use_weights is written only by batch 3's optimizer (NRCRenderGraph.cpp:66-68), and that shader returns early when its
    batch is empty (nrc_optimize.comp:33-34): after a frame whose LAST batch is empty the trained weights have moved three
    times but inference still sees the weights published before the frame."""
    w32 = he_weights(91)
    nb = 4096
    recs = [dev(random_records(30 + b, nb)) for b in range(4)]
    tgts = [dev(np.random.default_rng(40 + b).uniform(0, 1, (nb, 3)).astype(np.float32)) for b in range(4)]
    state.set_weights(w32)
    state.set_use_ema_weights(False)
    before = state.download()
    counts = [torch.tensor([c], dtype=torch.int32, device="cuda") for c in (nb, 100, 777, 0)]
    state.train_frame_unpacked(recs, tgts, counts)
    after = state.download()
    assert after["optimizer_state"]["t"] == 3
    assert not np.array_equal(after["weights"].view(np.uint16), before["weights"].view(np.uint16))
    assert np.array_equal(after["use_weights"].view(np.uint16), before["use_weights"].view(np.uint16))  # not republished
    # ... and the next frame with a non-empty last batch publishes the current weights
    counts = [torch.tensor([c], dtype=torch.int32, device="cuda") for c in (0, 0, 0, 5)]
    state.train_frame_unpacked(recs, tgts, counts)
    later = state.download()
    assert later["optimizer_state"]["t"] == 4
    assert np.array_equal(later["use_weights"].view(np.uint16), later["weights"].view(np.uint16))


def _poison():
    lib = os.path.join(HERE, "helpers", "_build", "libtesthelpers.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(HERE, "helpers")], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(lib)
    L.poison_shared_memory.argtypes = [C.c_void_p]
    assert L.poison_shared_memory(torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()


def test_cta_whose_first_tile_comes_in_a_later_batch(nrc, state):
    """A CTA without a tile in batch 0 never runs the TMA weight load whose out-of-bounds fill pads W_5 to 64 rows; it stages
    the weights itself in the first batch where it has work and must write those padding rows as zeros. Shared memory is
    poisoned with fp16 NaNs first: a stale padding row would turn every dA of such a CTA into NaN (silently zeroed
    gradients). One launch for the frame must equal four single-batch launches bit for bit."""
    w32 = he_weights(93)
    nb = nrc.TRAIN_BATCH_SIZE
    recs = [dev(random_records(50 + b, nb)) for b in range(4)]
    tgts = [dev(np.random.default_rng(60 + b).uniform(0, 1, (nb, 3)).astype(np.float32)) for b in range(4)]
    res = []
    for frame in (True, False):
        state.set_weights(w32)
        counts = [torch.tensor([c], dtype=torch.int32, device="cuda") for c in (100, nb - 384, nb, 16000)]
        _poison()
        if frame:
            state.train_frame_unpacked(recs, tgts, counts)
        else:
            for b in range(4):
                state.train_batch_unpacked(recs[b], tgts[b], count=counts[b], write_use_weights=(b == 3))
        res.append(state.download())
    assert same(res[0], res[1])
    assert np.isfinite(res[0]["gradients"]).all() and np.abs(res[0]["gradients"][:nrc.WEIGHT_COUNT]).max() > 0


def test_nrc_frame_is_inference_then_training_and_replays_from_a_graph(nrc):
    """nrc_frame_begin + nrc_frame = PreExecute's counter reset (NRCRenderGraph.cpp:108-112) + nn_inference_pass +
    4 x nn_train_pass (:46-80), capturable: three replays of the captured pair (with a 'producer' copy that refills the
    records and counts in between, as the path tracer would) equal three direct frames bit for bit."""
    from vknrc_b200 import synth
    sc = make_scene(71)
    dsc = nrc.DeviceScene(sc.vertices, sc.vertex_indices, sc.texcoords, sc.texcoord_indices, sc.materials, sc.material_ids, sc.transforms,
                          sc.textures)
    W, H, cap = 160, 96, nrc.TRAIN_BATCH_SIZE
    fr = synth.frame_records(72, W, H, sc.material_ids.shape[0], sc.transforms.shape[0], train_probability=0.2, batch_size=cap)
    n_ev = int(fr["eval_count"])
    src_ev = dev(fr["eval_records"].view(np.uint8).reshape(-1))
    src_tr = [dev(t.view(np.uint8).reshape(-1)) for t in fr["train_records"]]
    src_cnt = torch.tensor([n_ev] + [int(c) for c in fr["train_counts"]], dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(73)
    src_bf = dev(rng.uniform(0, 1, (H, W, 4)).astype(np.float32))
    d_gb = dev(rng.uniform(0, 1, (H, W, 2)).astype(np.float32))
    d_ev, d_tr, d_bf = torch.empty_like(src_ev), [torch.empty_like(t) for t in src_tr], torch.empty_like(src_bf)
    d_cnt = torch.full((5,), 12345, dtype=torch.int32, device="cuda")
    ev_c, tr_c = d_cnt[0:1], [d_cnt[1 + b:2 + b] for b in range(4)]
    st = nrc.NrcState(0, (W, H), seed=5)
    w32 = he_weights(74)

    def one_frame():
        st.frame_begin(ev_c, tr_c)               # counters <- 0
        d_ev.copy_(src_ev), d_bf.copy_(src_bf)   # the "producer": this frame's records, images and counts
        for a, b in zip(d_tr, src_tr):
            a.copy_(b)
        d_cnt.add_(src_cnt)                      # (appends on top of the zeroed counters, like the path tracer's atomics)
        st.frame(d_ev, ev_c, dsc, d_bf, d_gb, W, d_tr, tr_c, max_eval_count=n_ev)

    st.set_weights(w32)
    for _ in range(3):
        one_frame()
    direct, direct_bf, direct_cnt = st.download(), d_bf.clone(), d_cnt.clone()
    assert [int(c) for c in direct_cnt[1:]] == [min(int(c), cap) for c in fr["train_counts"]] and int(direct_cnt[0]) == n_ev
    # the same frames: inference (separate call) then training (separate call)
    st.set_weights(w32)
    for _ in range(3):
        d_cnt.zero_(), d_ev.copy_(src_ev), d_bf.copy_(src_bf)
        for a, b in zip(d_tr, src_tr):
            a.copy_(b)
        d_cnt.add_(src_cnt)
        st.infer(d_ev, ev_c, dsc, d_bf, d_gb, W, d_tr, max_count=n_ev)
        st.train_frame(d_tr, dsc, tr_c, max_count=cap)
    split = st.download()
    assert same(direct, split) and torch.equal(direct_bf, d_bf)
    # captured once, replayed three times
    st.set_weights(w32)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            one_frame()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    st.set_weights(w32)
    d_cnt.fill_(999)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    replayed = st.download()
    assert same(direct, replayed) and torch.equal(direct_bf, d_bf) and torch.equal(direct_cnt, d_cnt)
    assert replayed["optimizer_state"]["t"] == sum(3 for c in fr["train_counts"] if c > 0)
    st.close()

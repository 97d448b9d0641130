"""CPU checks of the path-structured frame record generator (vknrc_b200.synth.frame_records): the buffers must
have the structure shader/src/path_tracer.comp:254-375 produces, and decode with the oracle's dst codec."""
import numpy as np

import oracle
from vknrc_b200 import synth


def test_frame_records_have_the_path_tracers_structure():
    W, H, cap = 160, 64, 512
    fr = synth.frame_records(5, W, H, n_prims=300, n_instances=3, train_probability=0.2, batch_size=cap)
    ev, n_ev = fr["eval_records"], fr["eval_count"]
    assert n_ev == len(ev) > W * H
    # every pixel has exactly one screen-destined query, in the first W*H records
    dec = [oracle.dst_decode(int(d)) for d in ev["dst"]]
    screen = dec[:W * H]
    assert all(kind == 0 for kind, *_ in screen)
    assert {(x, y) for _, x, y, _ in screen} == {(x, y) for y in range(H) for x in range(W)}
    # tail queries: train-destined, disjoint contiguous ranges inside the batch capacity
    tails = dec[W * H:]
    assert tails and all(kind == 1 for kind, *_ in tails)
    covered = [np.zeros(cap, np.int32) for _ in range(4)]
    for _, b, l, r in tails:
        assert 0 <= b < 4 and 0 <= l <= r < cap and r - l + 1 <= synth.MAX_BOUNCE
        covered[b][l:r + 1] += 1
    assert all(c.max() <= 1 for c in covered)  # a record is fed back by at most one tail query
    # counts run past the capacity; records beyond min(count, cap) stay zero
    for b in range(4):
        cnt = int(fr["train_counts"][b])
        filled = min(cnt, cap)
        rec = fr["train_records"][b]
        assert cnt > 0 and (rec["factor"][:filled] > 0).all() and (rec["factor"][:filled] <= 0.95 + 1e-6).all()
        assert not rec["factor"][filled:].any() and not rec["bias"][filled:].any()
    assert fr["train_counts"].max() > cap  # (this configuration overflows at least one batch)


def test_frame_records_suffix_scan():
    """bias_i = radiance collected from vertex i on, factor_i = throughput from vertex i to the tail (path_tracer.comp:347-350):
    along a path factor is non-increasing towards the head and bias_i = light_i + color_i * bias_{i+1} >= 0."""
    fr = synth.frame_records(9, 96, 32, n_prims=100, n_instances=2, train_probability=0.5, batch_size=4096)
    dec = [oracle.dst_decode(int(d)) for d in fr["eval_records"]["dst"][96 * 32:]]
    assert dec
    for _, b, l, r in dec[:200]:
        f = fr["train_records"][b]["factor"][l:r + 1]
        assert (np.diff(f, axis=0) >= -1e-7).all()  # factor_i = color_i * factor_{i+1} <= factor_{i+1}
        assert (fr["train_records"][b]["bias"][l:r + 1] >= 0).all()

"""Training step latency for small batches (one launch: gradient + reduction + Adam), 14-float records."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vknrc_b200 as nrc
st = nrc.NrcState(0, (64, 64), seed=1)
g = torch.Generator(device="cuda").manual_seed(1)
for n in (128, 1024, 2048, 4096, 8192, 16384):
    rec, tgt = torch.rand((n, 14), device="cuda", generator=g), torch.rand((n, 3), device="cuda", generator=g)
    for _ in range(5):
        st.train_batch_unpacked(rec, tgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        st.train_batch_unpacked(rec, tgt)
    e1.record()
    torch.cuda.synchronize()
    print(f"train_batch {n:6d} records: {e0.elapsed_time(e1) / 50 * 1e3:6.2f} us")

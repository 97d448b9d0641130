"""Developer smoke on a GPU box: every kernel flavour against the CPU oracle, printing error magnitudes."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
import vknrc_b200 as nrc

torch.cuda.init()
dev = "cuda:0"
rng = np.random.default_rng(7)
W32 = (rng.standard_normal(nrc.WEIGHT_COUNT) * np.sqrt(2 / 64)).astype(np.float32)
W16 = W32.astype(np.float16)

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))

def section(name): print(f"--- {name}", flush=True)

section("evaluate_encoded")
for n in (128, 1000, 16384, 100000):
    x = rng.uniform(0, 1, (n, 64)).astype(np.float16)
    y = nrc.mlp_evaluate_encoded(torch.from_numpy(W16).to(dev), torch.from_numpy(x).to(dev))
    torch.cuda.synchronize()
    ref = oracle.evaluate(W16, x, oracle.ACC_FP32).astype(np.float32)
    print(n, "rel err vs oracle fp32acc", rel(y.cpu().numpy().astype(np.float32), ref), flush=True)

section("gradient_encoded")
for n in (128, 1000, 16384, 40000):
    x = rng.uniform(0, 1, (n, 64)).astype(np.float16)
    t = rng.uniform(0, 1, (n, 3)).astype(np.float16)
    dw = torch.zeros(nrc.WEIGHT_COUNT, dtype=torch.float32, device=dev)
    nrc.mlp_gradient_encoded(torch.from_numpy(W16).to(dev), dw, torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev))
    torch.cuda.synchronize()
    ref = oracle.gradient(W16, x, t.astype(np.float32), oracle.LOSS_L2, 1.0, oracle.ACC_FP32)
    g = dw.cpu().numpy()
    print(n, "dW rel err", rel(g, ref), "per layer", [round(rel(g[l*4096:(l+1)*4096], ref[l*4096:(l+1)*4096]), 5) for l in range(6)], flush=True)

section("state: unpacked inference + training step")
st = nrc.NrcState(0, (64, 64), seed=3)
st.set_weights(W32)
n = 5000
rec = np.concatenate([rng.uniform(-3, 3, (n, 3)), rng.uniform(0, 1, (n, 11))], axis=1).astype(np.float32)
y = st.infer_unpacked(torch.from_numpy(rec).to(dev))
torch.cuda.synchronize()
enc = oracle.encode(rec)
ref = oracle.evaluate(W16, enc, oracle.ACC_FP32, clamp=True).astype(np.float32)
print("infer_unpacked rel err", rel(y.cpu().numpy().astype(np.float32), ref), flush=True)

tg = rng.uniform(0, 1, (n, 3)).astype(np.float32)
st.gradient_unpacked(torch.from_numpy(rec).to(dev), torch.from_numpy(tg).to(dev))
d = st.download()
gref, yref = oracle.gradient(W16, enc, tg, oracle.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle.ACC_FP32, want_y=True)
print("gradient_unpacked rel err", rel(d["gradients"][:nrc.WEIGHT_COUNT], gref), "count", d["gradients"][nrc.GRAD_COUNT_SLOT],
      "loss", d["gradients"][nrc.GRAD_LOSS_SLOT] / n, "oracle loss", oracle.loss_value(yref, tg, oracle.LOSS_RELATIVE_L2_LUMINANCE), flush=True)
opt = oracle.Optimizer(W32)
opt.step(d["gradients"][:nrc.WEIGHT_COUNT], n, True, False)
st.adam_step(True)
d2 = st.download()
print("adam: weight bits equal", np.array_equal(d2["weights"].view(np.uint16), opt.weights), "master max abs diff",
      float(np.abs(d2["optimizer_entries"]["weight"] - opt.entries["weight"]).max()), "state", d2["optimizer_state"], flush=True)

section("timing")
n = 1920 * 1080
x = torch.rand((n, 64), device=dev, dtype=torch.float16)
wt = torch.from_numpy(W16).to(dev)
out = torch.empty((n, 3), device=dev, dtype=torch.float16)
for _ in range(3): nrc.mlp_evaluate_encoded(wt, x, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): nrc.mlp_evaluate_encoded(wt, x, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"evaluate_encoded 1080p: {ms*1e3:.1f} us  -> {n/ms/1e6:.2f} Gq/s, {n*41344/ms/1e9:.1f} TFLOP/s", flush=True)
recs = torch.rand((n, 14), device=dev, dtype=torch.float32)
for _ in range(3): st.infer_unpacked(recs, outputs=out)
torch.cuda.synchronize(); e0.record()
for _ in range(10): st.infer_unpacked(recs, outputs=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"infer_unpacked 1080p: {ms*1e3:.1f} us  -> {n/ms/1e6:.2f} Gq/s", flush=True)
nb = 16384
trec = torch.rand((nb, 14), device=dev); ttg = torch.rand((nb, 3), device=dev)
for _ in range(3): st.train_batch_unpacked(trec, ttg)
torch.cuda.synchronize(); e0.record()
for _ in range(20): st.train_batch_unpacked(trec, ttg)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"train_batch 16384: {ms*1e3:.1f} us/step -> {nb/ms/1e3:.2f} Mrec/s", flush=True)

"""Generates tests/golden/nrc_golden_v1.npz IN THIS CONTAINER (needs /root/reference for oracle/_ref).

The reference ships no golden vectors, so the fixtures are made by running the reference's OWN CPU
`Evaluate` / `Train` (test/main.cpp:11-74, compiled unmodified into oracle/_ref) on seeded inputs drawn from the
distributions of the reference's test (weights U(-0.02, 0.02), inputs/targets U(0,1), test/main.cpp:95-101,155-162)
and from its production initialiser (He-normal, src/VkNRCState.cpp:39-44). The oracle's own outputs are stored next
to them so that later edits of oracle/nrc_oracle.c are caught bit-exactly.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle

def main():
    oracle.build()
    assert oracle.ref_available(), "needs /root/reference to build oracle/_ref"
    rng = np.random.default_rng(20240223)
    n = 384  # 3 workgroups of the reference
    out = {}
    w_he = (rng.standard_normal(oracle.WEIGHT_COUNT) * np.sqrt(2.0 / 64.0)).astype(np.float32)
    w_u = rng.uniform(-0.02, 0.02, oracle.WEIGHT_COUNT).astype(np.float32)
    x = rng.uniform(0, 1, (n, 64)).astype(np.float32).astype(np.float16)
    t = rng.uniform(0, 1, (n, 3)).astype(np.float32).astype(np.float16)
    rec = np.concatenate([rng.uniform(-4, 4, (n, 3)), rng.uniform(0, 1, (n, 11))], axis=1).astype(np.float32)
    out["weights_he_fp32"], out["weights_uniform_fp32"] = w_he, w_u
    out["inputs"], out["targets"], out["records14"] = x, t, rec
    for tag, w in (("he", w_he), ("uniform", w_u)):
        w16 = w.astype(np.float16)
        out[f"ref_evaluate_{tag}"] = oracle.ref_evaluate(w16, x)                      # the reference's own code
        out[f"oracle_evaluate_fp32acc_{tag}"] = oracle.evaluate(w16, x, oracle.ACC_FP32)
        out[f"oracle_evaluate_fp16acc_{tag}"] = oracle.evaluate(w16, x, oracle.ACC_FP16_CHUNK16)
    w16 = w_he.astype(np.float16)
    out["ref_train_he"] = oracle.ref_train(w16, x, t)                                 # only layer 5 is meaningful
    out["oracle_dw_l2_fp32acc_he"] = oracle.gradient(w16, x, t.astype(np.float32), oracle.LOSS_L2, 1.0, oracle.ACC_FP32)
    out["oracle_dw_l2_fp16acc_he"] = oracle.gradient(w16, x, t.astype(np.float32), oracle.LOSS_L2, 1.0, oracle.ACC_FP16_CHUNK16)
    out["oracle_encoded_records"] = oracle.encode(rec)
    out["oracle_dw_rel_fp32acc_he"] = oracle.gradient(w16, out["oracle_encoded_records"], t.astype(np.float32),
                                                      oracle.LOSS_RELATIVE_L2_LUMINANCE, 1.0, oracle.ACC_FP32)
    opt = oracle.Optimizer(w_he)
    for step in range(3):
        g = oracle.gradient(opt.weights, out["oracle_encoded_records"], t.astype(np.float32), oracle.LOSS_RELATIVE_L2_LUMINANCE, 1.0,
                            oracle.ACC_FP32)
        opt.step(g, n, True, True)
    out["oracle_adam3_weights"] = opt.weights.view(np.float16)
    out["oracle_adam3_use_weights_ema"] = opt.use_weights.view(np.float16)
    out["oracle_adam3_entries"] = opt.entries.view(np.float32).reshape(-1, 4)
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "nrc_golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes")

if __name__ == "__main__":
    main()

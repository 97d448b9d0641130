"""Development tool: times the training paths of a (variant) build of libnrc_b200.so.
   build:  python tools/lab_train.py build NAME [-DFLAG ...]     -> tools/lab_lib_NAME.so
   run  :  python tools/lab_train.py run                          (on the GPU box: every tools/lab_lib_*.so, one process each)"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bench_one():
    import torch
    import vknrc_b200 as nrc
    dev = "cuda:0"
    st = nrc.NrcState(0, (1920, 1080), seed=1)
    g = torch.Generator(device=dev).manual_seed(5)

    def timed(fn, steps=40, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps * 1e3
    nb = nrc.TRAIN_BATCH_SIZE
    rec = torch.rand((4, nb, 14), device=dev, generator=g)
    tgt = torch.rand((4, nb, 3), device=dev, generator=g)
    recs, tgts = [rec[b] for b in range(4)], [tgt[b] for b in range(4)]
    us_frame = timed(lambda: st.train_frame_unpacked(recs, tgts))
    us_batch = timed(lambda: st.train_batch_unpacked(recs[0], tgts[0], write_use_weights=True))
    big = 1 << 20
    brec = torch.rand((big, 14), device=dev, generator=g)
    btgt = torch.rand((big, 3), device=dev, generator=g)
    us_big = timed(lambda: st.train_batch_unpacked(brec, btgt, write_use_weights=True), steps=10, warm=2)
    x = torch.rand((big, 64), device=dev, generator=g).half()
    t16 = torch.rand((big, 3), device=dev, generator=g).half()
    us_big_enc = timed(lambda: st.gradient_encoded(x, t16), steps=10, warm=2)
    print(f"frame 4x16384: {us_frame:7.2f} us | batch 16384: {us_batch:6.2f} us | 2^20 unpacked+adam: {us_big:7.1f} us "
          f"({big * 115840 / us_big / 1e6:6.1f} TFLOP/s) | 2^20 encoded grad: {us_big_enc:7.1f} us ({big * 115840 / us_big_enc / 1e6:6.1f} TFLOP/s)")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        from vknrc_b200 import build
        out = os.path.join(ROOT, "tools", f"lab_lib_{sys.argv[2]}.so")
        build.build(force=True, extra_flags=tuple(sys.argv[3:]), out=out)
        log = open(out + ".obj/ptxas.log").read().split("\n")
        for i, l in enumerate(log):
            if "nrc_train_kernel" in l and "Compiling" in l:
                print(l.split("'")[1][:60], "|", log[i + 2].strip() if i + 2 < len(log) else "", "|", log[i + 3].strip() if i + 3 < len(log) else "")
        print("built", out)
    elif sys.argv[1] == "run":
        for so in sorted(glob.glob(os.path.join(ROOT, "tools", "lab_lib_*.so"))):
            print("==", os.path.basename(so), flush=True)
            subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, NRC_B200_LIB=so))
    else:
        bench_one()

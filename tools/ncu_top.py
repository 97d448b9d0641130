"""Summarise an .ncu-rep: headline metrics + top stall locations (source page). Usage: ncu_top.py file.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum", "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:100])
    for h, u, v in zip(hdr, units, r):
        if any(h == w or h.endswith("." + w) for w in want):
            print(f"   {h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try: s = int(r[isamp])
    except ValueError: continue
    data.append((s, r))
tot = sum(s for s, _ in data) or 1
agg = {}
for s, r in data:
    for i in stall_cols:
        if r[i] not in ("", "0"): agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
print("total samples", tot, "stall mix:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for s, r in sorted(data, key=lambda x: -x[0])[:topn]:
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{s:7d} {100*s/tot:5.1f}% ex={r[iex]:>9} {r[isrc][:80]:80s} {top}")

"""Summarise an .ncu-rep (read here, no GPU): headline metrics + stall samples per opcode."""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    for w in want:
        if w in hdr:
            print("  %-75s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            v = float(r[i] or 0)
            if v > 0.05:
                print("  stall %-40s %.3f" % (h.split("stalled_")[1].split("_per_issue")[0], v))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ci[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
agg = collections.Counter(); st = collections.defaultdict(collections.Counter)
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    op = re.sub(r"^@!?U?P\d+\s+", "", r[ci["Source"]].strip()).split()[0] if r[ci["Source"]].strip() else "?"
    agg[op] += f(r, "# Samples")
    for k in keys: st[op][k] += f(r, k)
print("total samples", tot)
for op, c in agg.most_common(14):
    print("  %-22s %6.1f%%  %s" % (op, 100 * c / tot, {k[6:]: int(v) for k, v in st[op].most_common(3)}))

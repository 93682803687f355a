"""Summarise an .ncu-rep (raw page + per-source-line instruction counts)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local",
        "lts__t_bytes.sum ", "lts__t_sectors_op_write.sum", "l1tex__t_bytes.sum "]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w.strip()) or w in h for w in want) and "Not Issued" not in h:
        if "stalled" in h and float(v or 0) < 0.15:
            continue
        print(f"{h:95s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
per, tot = [], 0
for r in rows[hi + 1:]:
    if r and r[0].isdigit():
        try:
            n, s = int(r[iE]), int(r[iS])
        except ValueError:
            continue
        per.append((n, s, int(r[0]), r[1].strip()[:100])); tot += n
per.sort(reverse=True)
print("total warp instructions (source-attributed):", tot)
for n, s, l, t in per[:top]:
    print(f"{n / tot * 100:5.1f}% inst {s:7d} smp  L{l}: {t}")

"""Summarise an .ncu-rep here (no GPU): key raw metrics + samples per SASS opcode / per source line."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
for vals in rows[2:]:
    for i, h in enumerate(hdr):
        if h in want or ("pipe" in h and "pct_of_peak_sustained_active" in h and "inst_executed" in h):
            print(f"{h:80s} {units[i]:10s} {vals[i]}")
    print("-" * 60)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
op = collections.Counter(); inst = collections.Counter(); stalls = collections.Counter(); tot = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr) - 5: continue
    s = int(r[ci["Warp Stall Sampling (All Samples)"]]); tot += s
    m = [x for x in r[ci["Source"]].split() if not x.startswith("@")]
    full = ".".join(m[0].split(".")[:3]) if m else "?"
    op[full] += s; inst[full] += int(r[ci["Instructions Executed"]])
    for c in stall_cols: stalls[c] += int(r[ci[c]] or 0)
print("total samples", tot, "total warp inst", sum(inst.values()))
for k, v in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{k:28s} samples {v:7d} {100*v/max(tot,1):5.1f}%  inst {inst[k]}")
print({k: round(100 * v / max(tot, 1), 1) for k, v in stalls.most_common(10)})

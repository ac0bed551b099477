"""Per-kernel table from a set of .ncu-rep files (read here, no GPU): duration, DRAM bytes and throughput, L2 hit rate,
issue utilisation, occupancy, registers.   python tools/ncu_table.py gpurun_out/r02_ncu_*.ncu-rep > profiles/r02_ncu_kernels.md"""
import csv, io, subprocess, sys, os
M = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
     "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram%", "lts__t_sector_hit_rate.pct": "l2hit%",
     "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
     "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
     "smsp__inst_executed.sum": "inst", "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes"}
def num(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    if u in ("gbyte",): v *= 1e9
    elif u in ("mbyte",): v *= 1e6
    elif u in ("kbyte",): v *= 1e3
    elif u in ("ms", "msecond"): v *= 1e3
    elif u in ("ns", "nsecond"): v *= 1e-3
    elif u in ("s", "second"): v *= 1e6
    return v
# DRAM % of peak = achieved GB/s / MEASURED_PEAKS.json hbm_gbs (6548.5); ncu's own dram__throughput metric is not in --set full here
print("| capture | kernel | grid x block | regs | time us | DRAM rd+wr MB | achieved GB/s | DRAM % of peak | L2 hit % | issue % | occupancy % | lanes/32 | Mwarp-inst |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for vals in rows[2:]:
        d = {}
        for k, short in M.items():
            if k in ix and vals[ix[k]] not in ("", "n/a"):
                d[short] = num(vals[ix[k]], units[ix[k]])
        name = vals[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("cs::", "")
        mb = (d.get("rd", 0) + d.get("wr", 0)) / 1e6
        gbs = (d.get("rd", 0) + d.get("wr", 0)) / max(d.get("us", 1), 1e-9) / 1e3
        cap = os.path.basename(rep).replace("r02_ncu_", "").replace(".ncu-rep", "")
        print(f"| {cap} | `{name}` | {int(d.get('grid', 0))} x {int(d.get('block', 0))} | {int(d.get('regs', 0))} | {d.get('us', 0):.1f} | {mb:.1f} | {gbs:.0f} | "
              f"{100 * gbs / 6548.5:.1f} | {d.get('l2hit%', 0):.1f} | {d.get('issue%', 0):.1f} | {d.get('occ%', 0):.1f} | {d.get('lanes', 0):.1f} | {d.get('inst', 0) / 1e6:.1f} |")

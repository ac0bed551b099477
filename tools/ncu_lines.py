"""Per CUDA source line: stall samples and executed warp instructions from an .ncu-rep (needs -lineinfo + --import-source)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", ""): continue
    try: data.append((int(r[4]), int(r[7]), cur_file, int(r[0]), r[1].strip()[:110]))
    except ValueError: pass
ts = sum(d[0] for d in data); ti = sum(d[1] for d in data)
print("samples", ts, "warp inst", ti)
print("--- by samples")
for d in sorted(data, reverse=True)[:top]: print(f"{100*d[0]/ts:5.1f}% s  {100*d[1]/ti:5.1f}% i  {d[2]}:{d[3]:<4d} {d[4]}")
print("--- by instructions")
for d in sorted(data, key=lambda x: -x[1])[:top]: print(f"{100*d[0]/ts:5.1f}% s  {100*d[1]/ti:5.1f}% i  {d[2]}:{d[3]:<4d} {d[4]}")

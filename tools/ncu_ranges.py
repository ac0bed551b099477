"""Aggregate executed warp instructions / stall samples of an .ncu-rep by (file, line range).  usage: ncu_ranges.py rep file:lo-hi:name ..."""
import csv, io, subprocess, sys
rep = sys.argv[1]
specs = []
for a in sys.argv[2:]:
    f, rng, name = a.split(":")
    lo, hi = rng.split("-")
    specs.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = ""; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", ""): continue
    try: data.append((cur, int(r[0]), int(r[4]), int(r[7])))
    except ValueError: pass
ts = sum(d[2] for d in data); ti = sum(d[3] for d in data)
acc = {s[3]: [0, 0] for s in specs}; other = [0, 0]; otherf = {}
for f, ln, s, i in data:
    for (sf, lo, hi, name) in specs:
        if f == sf and lo <= ln <= hi:
            acc[name][0] += s; acc[name][1] += i; break
    else:
        other[0] += s; other[1] += i; otherf[f] = otherf.get(f, 0) + i
print(f"total samples {ts} warp inst {ti}")
for name, (s, i) in acc.items(): print(f"{name:28s} {100*s/ts:5.1f}% samples {100*i/ti:5.1f}% inst  ({i/1e6:.1f} M)")
print(f"{'other':28s} {100*other[0]/ts:5.1f}% samples {100*other[1]/ti:5.1f}% inst", {k: round(v/1e6,1) for k, v in sorted(otherf.items(), key=lambda x: -x[1])[:8]})

import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
src = open('/root/repo/comfystereo_b200/csrc/cs_polylines.cu').read().split('\n')
# phase boundaries from markers in the source
marks = []
for i, l in enumerate(src, 1):
    for tag in ("// ---- A:", "// ---- B:", "// ---- C:", "// ---- D:", "// ---- D2:", "// ---- E:", "__device__", "template <int PER>"):
        if tag in l: marks.append((i, l.strip()[:60]))
cur = ""; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", ""): continue
    try: data.append((cur, int(r[0]), int(r[4]), int(r[7])))
    except ValueError: pass
ts = sum(d[2] for d in data); ti = sum(d[3] for d in data)
import collections
agg = collections.OrderedDict()
for f, ln, s_, i_ in data:
    if f != "cs_polylines.cu": key = f
    else:
        key = "?"
        for m, name in marks:
            if ln >= m: key = f"{m}:{name}"
    a = agg.setdefault(key, [0, 0]); a[0] += s_; a[1] += i_
for k, (s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*s_/ts:5.1f}% samples {100*i_/ti:5.1f}% inst  {k}")

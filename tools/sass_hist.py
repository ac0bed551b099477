"""SASS opcode histogram per kernel of the built library (no GPU needed) -> profiles/r02_sass_opcodes.txt.
Shows what the kernels are made of, and that the Blackwell data-movement instructions are really there
(UBLKCP = cp.async.bulk / TMA, SYNCS = mbarrier)."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "comfystereo_b200/libcomfystereo_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); hist[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
    if m and fn:
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        hist[fn][op.split(".")[0]] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
with open("profiles/r02_sass_opcodes.txt", "w") as f:
    f.write("SASS opcode histogram per kernel (static instruction counts; tools/sass_hist.py)\n\n")
    for (fn, h), nm in zip(hist.items(), names):
        nm = re.sub(r"\(.*", "", nm).replace("cs::", "")
        f.write(f"{nm}  ({sum(h.values())} instructions)\n")
        f.write("    " + "  ".join(f"{o} {c}" for o, c in h.most_common(14)) + "\n")
        extra = [f"{o} {h[o]} ({what})" for o, what in (("UBLKCP", "TMA bulk copy"), ("SYNCS", "mbarrier"), ("LDL", "local loads"), ("STL", "local stores"), ("DFMA", "fp64"), ("DADD", "fp64"), ("DMUL", "fp64"), ("MUFU", "sfu")) if h.get(o)]
        if extra:
            f.write("    notable: " + ", ".join(extra) + "\n")
        f.write("\n")
print(len(hist), "kernels")

import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
def find(o):
    if isinstance(o,dict):
        if "kernels" in o: return o["kernels"]
        for v in o.values():
            r=find(v)
            if r: return r
ks=find(d) or {}
print(round(d["value"],1), {k:round(v["ms_per_step"],3) for k,v in ks.items() if v["ms_per_step"]>0})

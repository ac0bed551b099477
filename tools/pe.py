import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), d["e2e"].get("ms_per_step"))

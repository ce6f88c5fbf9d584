import json,sys
f=sys.argv[1]
l=[x for x in open(f) if x.startswith("{")][-1]
d=json.loads(l); print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["frac"])
for c in d["kernel_classes"]: print(c)

import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], "n_gpus", l["n_gpus"])
print("stages", json.dumps(l["stages"]))
for k,v in l.get("extra_configs",{}).items(): print(k, json.dumps(v)[:900])

import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
sc=d["single_caller"]
print("value %.4g e2e %.4g | single-caller pageable %.4g pinned %.4g | adversarial %.4g" % (d["value"], d["e2e"]["value"], sc["pageable"]["value"], sc["pinned"]["value"], [c["value"] for c in d["secondary"] if c["metric"].startswith("adversarial")][0]))

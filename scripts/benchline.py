import json, sys
for f in sys.argv[1:]:
    try:
        j=[json.loads(l) for l in open(f) if l.startswith("{")][0]
        print(f.split('/')[-1], "value %.0f ms %.2f chain %.2f pool %.2f frac %.3f/%.3f chains %d e2e %.0f tmv %.0f tr %d" % (j["value"], j["ms_per_step"], j["roofline"]["launch_ms"], j["roofline_prepass"]["launch_ms"], j["roofline"]["frac"], j["roofline_prepass"]["frac"], j["config"]["chains_per_gpu"], j["e2e"]["value"], (j.get("e2e_tmvrnorm") or {}).get("value", 0), j["counters"]["transitions"]))
    except Exception as e:
        print(f, "ERR", e)

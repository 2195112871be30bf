import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms/step", d["ms_per_step"], "MDE/s", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "sgm", d["roofline"]["sgm_stage"]["ms"], "sgm frac", d["roofline"]["sgm_stage"]["frac"])
    for k in d["kernels"]:
        print("   %-20s x%-3g %8.4f ms  %8s GB/s  frac %s" % (k["kernel"], k["launches_per_step"], k["avg_ms"], k["achieved_gbs"], k["frac"]))

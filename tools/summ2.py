import json, sys
for line in sys.stdin:
    if not line.startswith("{"): continue
    d = json.loads(line); r = d["roofline"]
    if "trees" in r:
        print("value=%.3e ms/step=%.1f trees=%d expanded/tree=%.0f relaxed/tree=%.0f improved/tree=%.0f us/tree=%.1f plan_ms=%.0f waves/tree=%.0f far_visits/tree=%.0f" % (
            d["value"], d["ms_per_step"], r["trees"], r["nodes_expanded"]/max(r["trees"],1), r["edges_relaxed"]/max(r["trees"],1),
            r["improved"]/max(r["trees"],1), 1e3*r["us_per_launch"]*r["launches"]/max(r["trees"],1)/1e3, d["phases"]["plan_ms"], r.get("waves",0)/max(r["trees"],1), r.get("far_visits",0)/max(r["trees"],1)))

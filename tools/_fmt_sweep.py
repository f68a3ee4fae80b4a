import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if "error" in d:
        print(d["config"], "ERROR", d["error"][:120])
        continue
    print(d["config"], round(d["us_per_search"],1), round(d["wall_ms"]), int(d["expanded_per_search"]), int(d["relaxed_per_search"]), int(d["waves_per_search"]), d["costs_equal_first_config"], d["answered"])

"""Print the key fields of bench.py JSON lines read from stdin (one per line)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print("%s n=%d value=%.3e ms/step=%.2f e2e=%.3e adv_us=%.2f frac=%.4f moving=%.3f routes/s=%.3e launches=%s reroutes/s=%s" % (
        d["config"].get("workload"), d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"],
        r.get("us_per_launch", 0), r.get("frac", 0), d.get("moving_fraction", 0), d.get("routes_per_s", 0),
        d.get("gpu_launches"), d.get("reroutes_per_s")), d.get("phases", ""))

"""Timing probe of the tick phases on the default bench workload: runs blocks of ticks with parts of
phase A switched off (engine option debug_mask; results are then NOT reference semantics) to see
what each part costs.  GPU only:  python tools/tick_probe.py [--agents N]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from drift_b200.engine import Engine, ROUTER_BATCHED  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="regional1m")
    ap.add_argument("--agents", type=int, default=0)
    ap.add_argument("--masks", default="0,1,2,3,4,6,7")
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    n = args.agents or w["agents"]
    lg, start, model, rng, _kind = bench.build_workload(args.workload, n, 0, 0)
    eng = Engine(lg, n, route_arena_links=max(1 << 24, n * 1536))
    _, dst = bench.make_chains(model, rng, lg, start, 4 + 2 * 3)
    dst = dst.reshape(n, -1)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    eng.set_landmarks(8)
    eng.set_demand(0, bench.HOURLY, start, chain_capacity=4, router=ROUTER_BATCHED)
    eng.set_option("route_window", 300)
    eng.push_trips(dst[:, :4])
    t = bench.T_START
    for b in range(3):
        eng.push_trips(np.ascontiguousarray(dst[:, 4 + 2 * b:6 + 2 * b]))
        eng.run(300, 1.0, t)
        t += 300
    eng.sync()
    print("moving", eng.stats()["moving_agents"])
    for m in [int(x) for x in args.masks.split(",")]:
        eng.set_option("debug_mask", m)
        eng.phase_times(reset=True)
        eng.run(299, 1.0, t)  # stays inside the planning window
        eng.sync()
        ph = eng.phase_times(reset=True)
        print("mask", m, {k: round(v, 2) for k, v in ph.items() if isinstance(v, float)}, "moving", eng.stats()["moving_agents"])
    eng.close()


if __name__ == "__main__":
    main()

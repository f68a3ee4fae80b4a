"""Whole-horizon runs of the BASELINE.json configs (24 h city / 48 h regional; lib config.py:163-165 allows 1-48 h):
device-driven demand in blocks of 300 one-second ticks through ``drift_b200.block_runner.BlockSimulation`` --
every block pushes the s-t model's next trips, runs, drains the departure / arrival logs and closes the trip records
(optionally streamed to the reference's CSV format).  Prints one JSON line: ticks, trips, arena compactions, in-tick
fallback routes, wall time split into device and host parts.

    python tools/long_horizon.py --workload brussels100k --hours 24 [--csv trips.csv]
    python tools/long_horizon.py --workload regional1m --hours 48
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="brussels100k", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--hours", type=float, default=24.0)
    ap.add_argument("--agents", type=int, default=0)
    ap.add_argument("--ticks-per-block", type=int, default=300)
    ap.add_argument("--csv", default=None, help="stream the trip records to this CSV (reference format)")
    ap.add_argument("--json", default=None, help="stream the trip records to this JSON file (reference format)")
    ap.add_argument("--reroute", action="store_true", help="congestion-aware reroute at every block (extension)")
    ap.add_argument("--no-labels", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    from drift_b200 import build as _build

    _build.build()
    from drift_b200 import hierarchy
    from drift_b200._capi import DriftError
    from drift_b200.block_runner import BlockSimulation

    w = bench.WORKLOADS[args.workload]
    n = args.agents or w["agents"]
    lg, start, model, rng, kind = bench.build_workload(args.workload, n, args.seed, 0)
    hier = None
    t_h = 0.0
    if kind == "road" and not args.no_labels and not args.reroute:
        t0 = time.perf_counter()
        try:
            hier, _ = hierarchy.build_cached(lg, time_limit_s=300.0)
        except (DriftError, ValueError):
            hier = None
        t_h = time.perf_counter() - t0
    tpb = args.ticks_per_block
    sim = BlockSimulation(lg, start, bench.HOURLY, seed=args.seed, csv_path=args.csv, json_path=args.json,
                          ticks_per_block=tpb,
                          reroute=args.reroute, record_routes=False,
                          route_arena_links=max(1 << 24, n * w.get("arena_per_agent", 2560)),  # as bench.py sizes it
                          hierarchy=hier)
    blocks = int(round(args.hours * 3600 / tpb))
    # the s-t model's trips for the whole horizon, generated in slabs of 64 blocks (the reference asks its selector
    # once per departure; a trip chain is just the sequence of targets, drift_b200/demand.py)
    slab = 64
    dep = arr = 0
    t_dev = t_host = t_gen = 0.0
    peak_moving = 0
    prev = start
    t_all = time.perf_counter()
    for b0 in range(0, blocks, slab):
        nb = min(slab, blocks - b0)
        t0 = time.perf_counter()
        _, dst = bench.make_chains(model, rng, lg, prev, 2 * nb)
        dst = dst.reshape(n, nb, 2)
        prev = dst[:, -1, -1].copy()
        t_gen += time.perf_counter() - t0
        for b in range(nb):
            t0 = time.perf_counter()
            sim.engine.push_trips(np.ascontiguousarray(dst[:, b, :]))
            if sim.reroute and sim.engine.tick > 0:
                sim.engine.update_travel_times()
                sim.reroutes += sim.engine.reroute_moving(sim.router)
                sim.engine.invalidate_prefetch(sim.t, sim.delta, tpb)
            sim.engine.run(tpb, sim.delta, sim.t)
            sim.engine.sync()
            t_dev += time.perf_counter() - t0
            t0 = time.perf_counter()
            t_block0 = sim.t
            for _ in range(tpb):
                sim.t += sim.delta
            d = sim.engine.drain_departures_routes()
            a = sim.engine.drain_arrivals()
            sim._account(d, a, sim.engine.tick - tpb, t_block0)
            dep += len(d[0])
            arr += len(a[0])
            t_host += time.perf_counter() - t0
        st = sim.engine.stats()
        peak_moving = max(peak_moving, st["moving_agents"])
        sys.stderr.write("[long_horizon] block %d/%d  t=%.1f h  departures=%d arrivals=%d moving=%d compactions=%d\n" % (
            b0 + nb, blocks, sim.t / 3600.0, dep, arr, st["moving_agents"], st["arena_compactions"]))
    wall = time.perf_counter() - t_all
    st = sim.engine.stats()
    rows = sim.finish()
    out = dict(workload=args.workload, agents=n, nodes=lg.V, links=lg.L, demand=model, hours=args.hours,
               ticks=blocks * tpb, tick_seconds=sim.delta, departures=dep, completed_trips=arr, trip_rows=rows,
               reroutes=sim.reroutes, peak_moving_agents=int(peak_moving), arena_compactions=st["arena_compactions"],
               fallback_routes=st["fallback_routes"], tie_fallbacks=st["tie_fallbacks"],
               router="hub labels" if hier is not None else "trees + bounded searches",
               hierarchy_seconds=t_h, wall_seconds=wall, device_seconds=t_dev, host_accounting_seconds=t_host,
               demand_generation_seconds=t_gen, agent_steps_per_s_wall=n * blocks * tpb / wall,
               agent_steps_per_s_device=n * blocks * tpb / max(t_dev, 1e-9), csv=args.csv)
    sim.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

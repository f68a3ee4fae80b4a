"""All-moving worst case of the advance kernel: N agents, every one of them on
a link, ~5 % crossing per tick.  Reports the per-launch CUDA-event time of k_advance (standalone
kernel) and the phase times of the persistent tick kernel, as GB/s of algorithmic bytes
(33 B per moving agent-step) against the measured HBM peak.

    python tools/advance_roofline.py [--agents 16000000]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(agents=16_000_000, ticks=40, grid=64, seed=0):
    from drift_b200.engine import Engine, ROUTER_BATCHED
    from drift_b200.synth import grid_graph

    lg = grid_graph(grid, grid, seed=seed, len_lo=250.0, len_hi=350.0, speed_kph=50.0)
    rng = np.random.default_rng(seed)
    start = rng.integers(0, lg.V, size=agents).astype(np.int32)
    cand = rng.integers(0, lg.V, size=(agents, 2)).astype(np.int32)
    hourly = np.zeros(24)
    hourly[6] = 900.0  # 06:00-07:00: p = 900 * 4 / 3600 = 1 per one-second tick: everybody departs at tick 1
    eng = Engine(lg, agents, route_arena_links=agents * 2 * 2 * grid)
    eng.set_demand(seed, hourly, start, chain_capacity=2, router=ROUTER_BATCHED)
    eng.set_option("route_window", 10 ** 6)
    eng.push_trips(cand)
    eng.run(1, 1.0, 0.0)  # plan + depart
    eng.drain_departures()
    st = eng.stats()
    moving = st["moving_agents"]
    # persistent kernel phases
    eng.phase_times(reset=True)
    eng.run(ticks, 1.0, 3600.0)  # 07:00: nobody departs any more
    ph = eng.phase_times(reset=True)
    ev = eng.stats()["events_last_tick"]
    # standalone kernel, CUDA events on the engine stream
    eng.set_profiling(True)
    eng.run(ticks, 1.0, 3600.0 + ticks)
    kt = eng.kernel_times()
    eng.set_profiling(False)
    st2 = eng.stats()
    eng.close()
    us = 1e3 * kt["advance_ms"] / max(kt["advance_launches"], 1)
    mv = 0.5 * (moving + st2["moving_agents"])
    alg = mv * 33.0 + (agents - mv) * 1.0
    return dict(agents=agents, moving_agents=mv, events_per_tick=ev, links=lg.L,
                k_advance_us=us, k_advance_gbs=alg / (us * 1e-6) / 1e9,
                persistent_advance_us=ph["advance_us_per_tick"], persistent_resolve_us=ph["resolve_us_per_tick"],
                persistent_advance_gbs=alg / (ph["advance_us_per_tick"] * 1e-6) / 1e9,
                algorithmic_bytes=alg)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=16_000_000)
    a = ap.parse_args()
    r = measure(a.agents)
    peak = 6549.1
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    r["frac_k_advance"] = r["k_advance_gbs"] / peak
    r["frac_persistent_advance"] = r["persistent_advance_gbs"] / peak
    print(json.dumps(r))

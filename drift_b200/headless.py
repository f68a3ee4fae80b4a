"""Headless driver: runs the drop-in ``SimulationThread`` without the PyQt GUI and writes the trip
CSV / JSON, replacing the start/stop plumbing of lib/simulation_controller.py:256-345.

    python -m drift_b200.headless --graph g.graphml --agents 300 --mode random --hours 1 --seed 0

``--graph`` is read with the reference's own GraphLoader when the DRIFT tree is importable
(``--drift-root``), or from a lowered ``.npz`` fixture (tests/golden format) otherwise.
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import time

import numpy as np


def run_headless(graph, agents, mode="random", hours=1, accel=10, seed=None, settings=None, st_selector=None,
                 out_dir=".", on_tick=None, device=0, lowered=None, fast_selectors=False):
    """Returns (thread, wall_seconds).  One ``update_simulation_step`` per tick, no GUI pacing."""
    from .simulation_thread import SimulationThread

    if seed is not None:
        random.seed(seed)
        np.random.seed(seed)
    cwd = os.getcwd()
    os.makedirs(out_dir, exist_ok=True)
    os.chdir(out_dir)  # the data logger writes ./data/simulation_trips_<time>.csv like the reference
    try:
        th = SimulationThread(graph, agents, mode, duration_hours=hours, device=device, lowered=lowered)
        th.time_acceleration = accel
        th.fast_selectors = bool(fast_selectors)
        if settings is not None:
            th.apply_settings(settings)
        if st_selector is not None:
            th.st_selector = st_selector
        th.running = True
        if not th.initialize_simulation():
            raise RuntimeError("initialize_simulation failed: %r" % getattr(th, "_init_error", None))
        th.update_frequency = 10 ** 9  # the reference's own knob to skip the per-tick viz snapshot
        t0 = time.perf_counter()
        while th.update_simulation_step():
            if on_tick is not None:
                on_tick(th)
        wall = time.perf_counter() - t0
        th.running = True
        th.finish_simulation()
        return th, wall
    finally:
        os.chdir(cwd)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--graph", required=True)
    ap.add_argument("--agents", type=int, default=300)
    ap.add_argument("--mode", default="random", choices=["random", "activity", "zone", "hub", "gravity"])
    ap.add_argument("--hours", type=float, default=1)
    ap.add_argument("--accel", type=float, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=".")
    ap.add_argument("--json", action="store_true", help="also write the JSON export")
    ap.add_argument("--fast-selectors", action="store_true",
                    help="vectorised gravity / hub / zone models (drift_b200/fast_selection.py: same trips for the same seed)")
    ap.add_argument("--drift-root", default=os.environ.get("DRIFT_ROOT"), help="path of the DRIFT checkout")
    args = ap.parse_args(argv)
    if not args.drift_root:
        raise SystemExit("--drift-root (or $DRIFT_ROOT) is needed to load graph files with the reference's GraphLoader")
    sys.path.insert(0, args.drift_root)
    from lib.graph_loader import GraphLoader

    random.seed(args.seed)
    np.random.seed(args.seed)
    G = GraphLoader().load_graph(args.graph)
    th, wall = run_headless(G, args.agents, args.mode, hours=args.hours, accel=args.accel, seed=args.seed,
                            out_dir=args.out, fast_selectors=args.fast_selectors)
    if args.json:
        th.data_logger.save_to_json()
    n = args.agents * th.step
    print("ticks=%d agent-steps=%d wall=%.2fs rate=%.3g agent-steps/s trips=%d" % (
        th.step, n, wall, n / wall, len(th.data_logger.trips_data)))
    th.close()


if __name__ == "__main__":
    main()

"""BlockSimulation: streamed trip CSV in the reference's format, consistent with the device logs."""
import csv

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_block_runner_csv(tmp_path):
    from drift_b200 import demand
    from drift_b200.block_runner import BlockSimulation
    from drift_b200.data_logger import FIELDNAMES
    from drift_b200.synth import road_graph

    lg = road_graph(3000, seed=2)
    N = 5000
    rng = np.random.default_rng(1)
    start = rng.integers(0, lg.V, size=N).astype(np.int32)
    hourly = np.full(24, 0.9)
    path = tmp_path / "trips.csv"
    sim = BlockSimulation(lg, start, hourly, seed=3, csv_path=str(path), ticks_per_block=200, record_routes=True)
    n_dep = n_arr = 0
    for b in range(4):
        _, cand = demand.random_chains(rng, lg.V, start, 2)
        d, a = sim.run_block(cand.reshape(N, 2))
        n_dep += d
        n_arr += a
    st = sim.status()
    assert st["total_agents"] == N and 0 < st["moving_agents"] < N
    rows_total = sim.finish()
    sim.close()
    with open(path, newline="", encoding="utf-8") as fh:
        rows = list(csv.DictReader(fh))
    assert list(rows[0].keys()) == FIELDNAMES
    assert len(rows) == rows_total
    # every agent has exactly one first-trip row (origin == destination == its start node, 06:00:00)
    first = [r for r in rows if r["start_time"] == "06:00:00" and r["origin_node"] == r["destination_node"]]
    assert len(first) == N
    assert n_arr > 500 and len(rows) >= n_arr
    ids = [int(r["trip_id"][1:]) for r in rows]
    assert len(set(ids)) == len(ids) and min(ids) == 1
    done = [r for r in rows if float(r["trip_distance_km"]) > 0]
    assert sum(len(eval(r["route_taken"])) > 2 for r in done) > len(done) // 2
    for r in done[:200]:
        route = eval(r["route_taken"])
        assert route[-1] == int(r["destination_node"]) or r in first
        assert float(r["trip_duration_min"]) > 0 and float(r["average_speed_kmh"]) > 0
        if len(route) > 2:  # [origin, destination] stands in when the route was overwritten before export
            for a, b in zip(route[:-1], route[1:]):  # consecutive nodes are joined by a link
                assert lg.hop_entry(a, b) >= 0

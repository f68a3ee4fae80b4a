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
    jpath = tmp_path / "trips.json"
    sim = BlockSimulation(lg, start, hourly, seed=3, csv_path=str(path), ticks_per_block=200, record_routes=True,
                          json_path=str(jpath))
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
    # get_statistics() from running sums == the reference's expression over the rows (lib/data_logger.py:154-171)
    stats = sim.get_statistics()
    km = sum(float(r["trip_distance_km"]) for r in rows)
    minutes = sum(float(r["trip_duration_min"]) for r in rows)
    assert stats["total_trips"] == len(rows) and stats["total_distance_km"] == km
    assert stats["total_duration_hours"] == minutes / 60
    assert stats["average_speed_kmh"] == sum(float(r["average_speed_kmh"]) for r in rows) / len(rows)
    # the streamed JSON export (lib/data_logger.py:260-306) carries the same trips, routes as lists
    import json
    js = json.load(open(jpath, encoding="utf-8"))
    assert set(js) == {"metadata", "trips"} and js["metadata"]["total_trips"] == rows_total == len(js["trips"])
    for r, j in zip(rows[:300] + rows[-300:], js["trips"][:300] + js["trips"][-300:]):
        assert list(j.keys()) == FIELDNAMES
        assert j["trip_id"] == r["trip_id"] and j["route_taken"] == eval(r["route_taken"])
        assert j["trip_distance_km"] == float(r["trip_distance_km"])
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


@pytest.mark.parametrize("labels", [False, True])
def test_block_runner_rows_equal_port(tmp_path, labels):
    """The streamed CSV against the oracle port replaying the device's departure log: same rows in the same
    order -- trip ids, stamps, origin / destination, route_taken (incl. the reference's first-trip and
    falsy-label quirks: node 0 is skipped), fp64 distance / duration / speed -- for agents that arrive within the
    block they left in, depart twice in a block, are still under way at the end, or never leave."""
    from drift_b200 import demand, hierarchy
    from drift_b200.block_runner import BlockSimulation
    from drift_b200.synth import road_graph
    from oracle.port import PortSim, plain_from_arrays

    lg = road_graph(1500, seed=6)
    N = 4000
    rng = np.random.default_rng(4)
    start = rng.integers(0, lg.V, size=N).astype(np.int32)
    start[:40] = 0  # the falsy label
    hourly = {h: 0.9 for h in range(24)}
    path = tmp_path / "trips.csv"
    sim = BlockSimulation(lg, start, np.full(24, 0.9), seed=5, csv_path=str(path), ticks_per_block=150,
                          record_routes=True, keep_log=True, route_arena_links=N * 256,
                          hierarchy=hierarchy.build(lg) if labels else None)
    sim.engine.set_option("arena_compact_frac", 0.05)  # compact at every window: finished routes are dropped
    blocks = 6
    for b in range(blocks):
        _, cand = demand.random_chains(rng, lg.V, start, 2)
        cand = cand.reshape(N, 2)
        cand[:20, 0] = 0
        sim.run_block(cand)
    assert sim.engine.stats()["arena_compactions"] > 0  # finished routes were dropped in between
    sim.finish()
    sim.close()
    with open(path, newline="", encoding="utf-8") as fh:
        rows = list(csv.DictReader(fh))
    pg = plain_from_arrays(lg.nodes, lg.link_u, lg.link_v, lg.link_len, lg.link_v0, lg.link_t0, lg.link_cap,
                           lg.fwd_rowptr, lg.fwd_col, lg.fwd_w, lg.fwd_link, lg.fwd_minlen, lg.bwd_rowptr,
                           lg.bwd_col, lg.bwd_w)
    trace = {}
    for tk, a, s, d, st in sim.departure_log:
        trace.setdefault(tk, {})[a] = (s, d, "route")
    port = PortSim(pg, N, hourly=hourly, accel=10, hours=1, departure_trace={}, init_nodes=start)

    class _Trace(dict):
        def __missing__(self, k):
            return {}

    port.trace = _Trace(trace)
    for _ in range(blocks * 150):
        port.tick()
    port.finish()
    assert len(rows) == len(port.trips) and len(rows) > N
    twice = 0
    seen = {}
    for r, w in zip(rows, port.trips):
        assert r["trip_id"] == w["trip_id"]
        assert int(r["agent_id"][1:]) - 1 == w["agent_index"]
        for k in ("start_time", "end_time", "route_taken"):
            assert r[k] == w[k], (k, r, w)
        assert r["origin_node"] == str(w["origin_node"]) and r["destination_node"] == str(w["destination_node"])
        for k in ("trip_distance_km", "trip_duration_min", "average_speed_kmh"):
            assert float(r[k]) == w[k], (k, r, w)  # csv writes repr(float): exact round trip
        seen[w["agent_index"]] = seen.get(w["agent_index"], 0) + 1
    twice = sum(1 for v in seen.values() if v >= 3)
    assert twice > 10  # several trips per agent were recorded

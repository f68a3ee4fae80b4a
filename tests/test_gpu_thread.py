"""The drop-in SimulationThread, seeded like the reference, against the recorded reference runs:
same MT19937 stream, same (s, t) draws, NetworkX-identical routes, bit-equal volumes, trip records
equal row by row (agent_id aside, which is id(obj) in the reference)."""
import csv
import os
import random

import numpy as np
import pytest

from oracle import golden

pytestmark = pytest.mark.gpu
LIVE = [n for n in golden.case_names() if "_random_" in n or "congested" in n]
# recorded with the reference's own hub / zone / gravity selector objects; replayed here with the vectorised
# restatements of drift_b200/fast_selection.py (the reference package does not exist on the GPU box)
LIVE_FAST = [n for n in golden.case_names() if any(m in n for m in ("_hub_", "_zone_", "_gravity_", "_activity_"))]


def _settings(case):
    st = case.params.get("settings")
    if not st:
        return None
    st = dict(st)
    if "hourly_probabilities" in st:
        st["hourly_probabilities"] = {int(k): v for k, v in st["hourly_probabilities"].items()}
    return st


@pytest.mark.parametrize("name", LIVE + LIVE_FAST)
def test_thread_matches_reference_run(name, tmp_path):
    from drift_b200.headless import run_headless
    from drift_b200.simulation_thread import AgentView

    case = golden.GoldenCase(name)
    G = case.nx_graph()
    occ_ref = case.occupancy()
    deps_ref = {}
    for tick, agent, s, t, path in case.departures():
        deps_ref.setdefault(tick, []).append((agent, s, t, path))
    seen = {"max_tick": 0}
    logs = []

    def on_tick(th):
        k = th.step
        assert th.simulation_time == case.times[k - 1]
        assert np.array_equal(th.engine.occupancy(), occ_ref[k - 1]), "link volumes differ at tick %d" % k
        seen["max_tick"] = k
        if case.params.get("max_ticks") and k >= case.params["max_ticks"]:
            th.running = False

    AgentView._next_id = 1
    th, wall = run_headless(G, case.N, case.params["mode"], hours=case.hours, accel=case.accel,
                            seed=case.params["seed"], settings=_settings(case), out_dir=str(tmp_path), on_tick=on_tick,
                            fast_selectors=name in LIVE_FAST)
    try:
        assert seen["max_tick"] == case.ticks
        assert [a.source for a in th.agents][:0] == []
        want = case.trips
        got = th.data_logger.trips_data
        id2idx = {"A%06d" % id(a): a.index for a in th.agents}
        if case.finish_error:
            # the reference's finish_simulation dies with a TypeError half way through the flush
            # (SURVEY.md 0-9a); the drop-in guards it, so only the rows up to the crash compare
            assert len(got) > len(want)
            got = got[:len(want)]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            g = dict(g)
            g["agent_index"] = id2idx[g.pop("agent_id")]
            assert g == w
        stats = th.data_logger.get_statistics()
        for k, v in (case.stats or {}).items():
            assert stats[k] == v
        assert np.array_equal(np.array([a.trip_count for a in th.agents]), case.z["trip_counts"])
        # the CSV on disk has the reference's header and one row per trip
        files = os.listdir(tmp_path / "data")
        assert len(files) == 1 and files[0].startswith("simulation_trips_") and files[0].endswith(".csv")
        with open(tmp_path / "data" / files[0], newline="", encoding="utf-8") as fh:
            rows = list(csv.reader(fh))
        assert rows[0] == ["trip_id", "agent_id", "start_time", "end_time", "origin_node", "destination_node",
                           "route_taken", "trip_distance_km", "trip_duration_min", "average_speed_kmh"]
        assert len(rows) == len(th.data_logger.trips_data) + 1
        assert rows[1][0] == want[0]["trip_id"] and rows[1][6] == want[0]["route_taken"]
        assert rows[1][8] == repr(want[0]["trip_duration_min"])
    finally:
        th.close()


def test_thread_signals_and_status(tmp_path):
    from drift_b200.simulation_thread import SimulationThread

    case = golden.GoldenCase("g40_random_s1")
    G = case.nx_graph()
    random.seed(3)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        th = SimulationThread(G, 120, "random", duration_hours=1)
        th.time_acceleration = 100  # 10 s ticks as in the shipped CSVs
        th.apply_settings(dict(bpr_alpha=0.15, bpr_beta=4.0, hourly_probabilities={h: 0.9 for h in range(24)}))
        got = dict(log=[], trips=[], status=[], updates=[], finished=[], ready=[])
        th.log_message.connect(got["log"].append)
        th.trip_completed.connect(got["trips"].append)
        th.status_updated.connect(got["status"].append)
        th.agents_updated.connect(got["updates"].append)
        th.simulation_finished.connect(got["finished"].append)
        th.st_selector_ready.connect(lambda s, m: got["ready"].append((s, m)))
        th.running = True
        assert th.initialize_simulation()
        th.update_frequency = 50
        n = 0
        while th.update_simulation_step():
            n += 1
        assert n == 360 and th.step == 360
        th.running = True
        th.finish_simulation()
        assert got["ready"] and got["ready"][0][1] == "random"
        assert len(got["status"]) == 3 and set(got["status"][0]) == {
            "simulation_time", "moving_agents", "total_agents", "network_utilization", "active_agent_types"}
        assert got["status"][-1]["total_agents"] == 120
        assert len(got["updates"]) == 7 and len(got["updates"][0]) == 120
        u = got["updates"][-1][0]
        for f in ("id", "position", "state", "agent_type", "trip_count", "speed", "source", "target", "path", "path_index"):
            assert hasattr(u, f)
        assert len(got["trips"]) > 10 and set(got["trips"][0]) >= {"trip_id", "agent_id", "start_node", "end_node", "path_nodes"}
        assert len(got["finished"]) == 1 and "csv_file" in got["finished"][0] and got["finished"][0]["total_trips"] >= 120
        st = th.traffic_manager.get_network_statistics()
        assert st["total_edges"] == case.L and st["total_network_capacity"] == int(case.capacities.sum())
        moving = [a for a in th.agents if a.state == "moving"]
        assert st["total_agents_on_roads"] == len(moving)
        if moving:
            a = moving[0]
            assert a.current_edge_key in th.traffic_manager.edge_capacities
            assert a in th.traffic_manager.edge_agents[a.current_edge_key]
            assert 0 <= a.distance_on_edge < a.edge_length and a.speed > 0
        th.data_logger.save_to_json("t.json")
        import json
        js = json.load(open(tmp_path / "data" / "t.json"))
        assert js["metadata"]["total_trips"] == len(th.data_logger.trips_data) and isinstance(js["trips"][0]["route_taken"], list)
        th.close()
    finally:
        os.chdir(cwd)

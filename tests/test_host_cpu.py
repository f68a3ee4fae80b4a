"""CPU-only checks of the host side: lowering, BPR tables, C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import golden
from oracle.port import bpr_factor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = golden.case_names()


@pytest.mark.parametrize("name", CASES)
def test_lowering_matches_fixture(name):
    from drift_b200.graph import lower_graph

    case = golden.GoldenCase(name)
    lg = lower_graph(case.nx_graph())
    pg = case.plain_graph()
    assert lg.V == case.V and lg.L == case.L
    assert np.array_equal(lg.link_u, case.link_u) and np.array_equal(lg.link_v, case.link_v)
    assert np.array_equal(lg.link_len, case.link_len) and np.array_equal(lg.link_v0, case.link_v0)
    assert np.array_equal(lg.link_t0, case.link_tt)
    assert np.array_equal(lg.link_cap, case.capacities)  # recorded from the reference's TrafficManager
    for u in range(lg.V):
        a, b = lg.fwd_rowptr[u], lg.fwd_rowptr[u + 1]
        want = pg.succ[u]
        assert lg.fwd_col[a:b].tolist() == [r[0] for r in want]
        assert lg.fwd_w[a:b].tolist() == [r[1] for r in want]
        assert lg.fwd_link[a:b].tolist() == [r[2] for r in want]
        assert lg.fwd_minlen[a:b].tolist() == [r[3] for r in want]
        a, b = lg.bwd_rowptr[u], lg.bwd_rowptr[u + 1]
        assert lg.bwd_col[a:b].tolist() == [r[0] for r in pg.pred[u]]
        assert lg.bwd_w[a:b].tolist() == [r[1] for r in pg.pred[u]]
        for e in range(a, b):
            l = lg.bwd_link[e]
            assert lg.link_u[l] == lg.bwd_col[e] and lg.link_v[l] == u


def test_from_links_matches_lower_graph():
    from drift_b200.graph import from_links, lower_graph, to_networkx
    from drift_b200.synth import grid_graph

    lg = grid_graph(7, 5, seed=3)
    lg2 = lower_graph(to_networkx(lg))
    for f in ("link_u", "link_v", "link_len", "link_v0", "link_t0", "link_cap", "fwd_rowptr", "fwd_col", "fwd_w",
              "fwd_link", "bwd_rowptr", "bwd_col", "bwd_w", "bwd_link"):
        assert np.array_equal(getattr(lg, f), getattr(lg2, f)), f


@pytest.mark.parametrize("alpha,beta", [(0.15, 4.0), (0.8, 2.5), (0.01, 1.0), (2.0, 8.0)])
def test_bpr_tables_equal_reference_expression(alpha, beta):
    from drift_b200.bpr import build_tables

    caps = np.array([1, 4, 12, 24, 160, 532, 532, 1500, 12])
    table, off, cls = build_tables(caps, alpha, beta)
    assert len(off) - 1 == len(np.unique(caps))
    for l, cap in enumerate(caps.tolist()):
        c = cls[l]
        row = table[off[c]:off[c + 1]]
        assert row[0] == 1.0
        n_check = min(len(row), 4000)
        for count in range(n_check):
            assert row[count] == bpr_factor(count, cap, alpha, beta)
        if alpha >= 0.15:
            assert row[-1] == 0.1
            # everything past the table is saturated too
            for count in (len(row), len(row) + 7, 100 * len(row)):
                assert bpr_factor(count, cap, alpha, beta) == 0.1


def test_capi_exports_every_declared_symbol():
    from drift_b200 import _capi, build

    build.build()
    header = open(os.path.join(ROOT, "include", "drift_b200.h")).read()
    declared = set(re.findall(r"\b(drift_[a-z0-9_]+)\s*\(", header))
    declared -= {"drift_engine", "drift_graph_desc", "drift_stats_out"}
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    lib = _capi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.drift_abi_version() == 1


def test_capi_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from drift_b200 import _capi
    from drift_b200.engine import Engine
    from drift_b200.synth import grid_graph

    with pytest.raises(_capi.DriftError) as ei:
        Engine(grid_graph(3, 3, seed=0), 8)
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)


def test_graph_cache_round_trip(tmp_path):
    from drift_b200 import graph_cache
    from drift_b200.graph import lower_graph

    case = golden.GoldenCase("multi_random_s0")
    G = case.nx_graph()
    lg = lower_graph(G)
    a = graph_cache.lower_cached(G, cache_dir=str(tmp_path))
    assert len(os.listdir(tmp_path)) == 1
    b = graph_cache.lower_cached(G, cache_dir=str(tmp_path))  # served from disk
    for x in (a, b):
        assert x.nodes == lg.nodes and x.link_key == lg.link_key
        for f in graph_cache._ARRAYS:
            assert np.array_equal(getattr(x, f), getattr(lg, f)), f
    G2 = case.nx_graph()
    G2[G2.nodes.__iter__().__next__()]  # same graph object content -> same digest
    assert graph_cache.graph_digest(G2) == graph_cache.graph_digest(G)
    u, v, k = next(iter(G2.edges(keys=True)))
    G2[u][v][k]["length"] += 1.0
    assert graph_cache.graph_digest(G2) != graph_cache.graph_digest(G)


def test_departure_threshold_is_the_bernoulli_test():
    """The tick kernel compares the 53 random bits with ceil(p * 2^53) instead of converting them to a
    double and testing u < p (lib/agent.py:180): the two are the same predicate for every p, including
    the values right next to a threshold."""
    import math

    from oracle.philox import uniform53

    rng = np.random.default_rng(5)
    agents = np.arange(20000, dtype=np.uint32)
    u = uniform53(99, agents, 1234)
    bits = np.round(u * 9007199254740992.0).astype(np.uint64)
    assert np.array_equal(bits.astype(np.float64) / 9007199254740992.0, u)  # u = bits * 2^-53 exactly
    ps = list(rng.random(20) * 1e-3) + [0.35 * 4 / 3600, 0.9 * 4 * 5 / 3600, 0.0, 1.0, 1.5, 2.0 ** -53, 3 * 2.0 ** -54]
    ps += [float(x) for x in u[:40]] + [float(np.nextafter(x, 1.0)) for x in u[:40]]  # on / just above a draw
    for p in ps:
        if not p > 0.0:
            thr = 0
        else:
            s = math.ceil(p * 9007199254740992.0)
            thr = min(s, 2 ** 64 - 1)
        assert np.array_equal(bits < np.uint64(thr) if thr < 2 ** 64 else np.ones_like(bits, bool), u < p), p


def test_paired_philox_words():
    """One Philox block serves the two agents of a pair: agents 2k and 2k+1 share a counter and take
    different output words, so their draws differ and do not depend on the neighbour's presence."""
    from oracle.philox import uniform53

    a = uniform53(7, np.array([10, 11, 12, 13], dtype=np.uint32), 5)
    assert len(set(a.tolist())) == 4
    assert uniform53(7, np.array([11], dtype=np.uint32), 5)[0] == a[1]
    assert uniform53(7, np.array([10], dtype=np.uint32), 6)[0] != a[0]
    assert np.all((a >= 0) & (a < 1))


@pytest.mark.parametrize("n,threshold", [(10, 0.3), (500, 0.01), (5000, 0.0004), (70000, 0.0004), (3000, 0.0), (200, 1.0)])
def test_bulk_departure_draws_equal_the_python_loop(n, threshold):
    """SimulationThread._draw_departures (MT19937 state handed to numpy, bulk uniforms, rewind to the first hit) ==
    the reference's loop of random.random() per waiting agent (lib/agent.py:180) with the selector's own draws in
    between: same departures, same generator state afterwards."""
    import random
    import types

    from drift_b200.simulation_thread import SimulationThread

    def journeys(log):
        def start(agent):  # a selector that consumes a varying number of draws, like random.choice's rejection loop
            k = 1 + (agent % 3)
            log.append((agent, [random.random() for _ in range(k)], random.randrange(1000)))
        return start

    waiting = sorted(random.Random(5).sample(range(2 * n), n))
    for seed in (1, 2):
        want_log, got_log = [], []
        random.seed(seed)
        want = []
        start = journeys(want_log)
        for i in waiting:
            if random.random() < threshold:
                want.append(i)
                start(i)
        state_want = random.getstate()
        random.seed(seed)
        fake = types.SimpleNamespace(_waiting=waiting, agents=list(range(2 * n)), _start_new_journey=journeys(got_log))
        got = SimulationThread._draw_departures(fake, threshold)
        assert got == want and got_log == want_log
        assert random.getstate() == state_want


def test_streamed_json_export_equals_the_logger_export(tmp_path):
    """TripJsonStream (block_runner's trip writer) against SimulationDataLogger.export_to_json, which follows
    lib/data_logger.py:260-306: same keys, same trips, same conversions; also the empty file."""
    import json

    from drift_b200.data_logger import FIELDNAMES, SimulationDataLogger, TripJsonStream

    rows = []
    for i in range(7):
        rows.append({"trip_id": "T%06d" % (i + 1), "agent_id": "A%06d" % (i % 3 + 1), "start_time": "06:00:0%d" % i,
                     "end_time": "06:10:0%d" % i, "origin_node": i, "destination_node": "n%d" % i,
                     "route_taken": str([i, "n%d" % i]) if i != 4 else "not a list(",
                     "trip_distance_km": 0.5 * i, "trip_duration_min": 1.5 * i,
                     "average_speed_kmh": 20.0 if i else 0.0, "extra": "dropped"})
    del rows[5]["trip_distance_km"]  # missing number -> 0.0, like trip.get(..., 0.0)
    lg = SimulationDataLogger(output_dir=str(tmp_path))
    lg.trips_data = [dict(r) for r in rows]
    lg.export_to_json(str(tmp_path / "a.json"))
    st = TripJsonStream(str(tmp_path / "b.json"), lg.simulation_start_time)
    st.write(rows[:3])
    st.write(rows[3:])
    st.close()
    a = json.load(open(tmp_path / "a.json", encoding="utf-8"))
    b = json.load(open(tmp_path / "b.json", encoding="utf-8"))
    assert a["trips"] == b["trips"] and [list(t.keys()) for t in b["trips"]] == [FIELDNAMES] * 7
    assert b["trips"][4]["route_taken"] == [] and b["trips"][5]["trip_distance_km"] == 0.0
    assert {k: v for k, v in a["metadata"].items() if k != "export_timestamp"} == \
           {k: v for k, v in b["metadata"].items() if k != "export_timestamp"}
    e = TripJsonStream(str(tmp_path / "c.json"), lg.simulation_start_time)
    e.close()
    c = json.load(open(tmp_path / "c.json", encoding="utf-8"))
    assert c["trips"] == [] and c["metadata"]["total_trips"] == 0


def test_running_trip_statistics_equal_the_list_based_ones():
    """TripStatistics (streamed runs) against get_statistics over the kept list (lib/data_logger.py:154-171),
    exact equality: the running sums end where the builtin sum does (compensated from Python 3.12 on)."""
    import random

    from drift_b200.data_logger import SimulationDataLogger, TripStatistics

    rnd = random.Random(5)
    lg = SimulationDataLogger(output_dir=".")
    st = TripStatistics()
    assert st.result() == {} == lg.get_statistics()
    for i in range(20000):
        km = rnd.random() * 10 ** rnd.randint(-3, 3) if i % 7 else 0.0
        minutes = rnd.random() * 10 ** rnd.randint(-2, 2)
        kmh = (km / minutes) * 60 if minutes > 0 else 0.0
        lg.trips_data.append({"trip_distance_km": km, "trip_duration_min": minutes, "average_speed_kmh": kmh})
        st.add(km, minutes, kmh)
        if i in (0, 1, 17, 4999, 19999):
            assert st.result() == lg.get_statistics()


def test_landmark_tables_are_lower_bounds():
    """engine.landmark_tables: the distance rows of the picked landmarks, and the ALT bounds the bounded searches
    derive from them (d(s,v) >= from[k][v] - from[k][s] and >= to[k][s] - to[k][v]) never exceed the distance."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    from drift_b200.engine import landmark_tables
    from drift_b200.synth import road_graph

    lg = road_graph(1500, seed=4)
    picks, frm, to = landmark_tables(lg, K=6, seed=1)
    assert len(picks) == 6 == len(set(picks)) and frm.shape == to.shape == (6, lg.V)
    W = csr_matrix((lg.fwd_w, lg.fwd_col, lg.fwd_rowptr), shape=(lg.V, lg.V))
    d_from = dijkstra(W, directed=True, indices=picks)
    assert np.array_equal(frm, d_from) and np.all(frm[np.arange(6), picks] == 0.0) and np.all(to[np.arange(6), picks] == 0.0)
    rng = np.random.default_rng(0)
    src = rng.integers(0, lg.V, size=40)
    d = dijkstra(W, directed=True, indices=src)  # d[i][v] = d(src_i, v)
    for i, s in enumerate(src):
        b0 = frm - frm[:, s:s + 1]          # [K][V]
        b1 = to[:, s:s + 1] - to
        h = np.maximum(np.nanmax(np.where(np.isfinite(b0), b0, -np.inf), axis=0),
                       np.nanmax(np.where(np.isfinite(b1), b1, -np.inf), axis=0))
        ok = np.isfinite(d[i])
        assert np.all(h[ok] <= d[i][ok] * (1 + 1e-12) + 1e-9)
        # what the kernel stores: the bound as a float rounded DOWN stays a lower bound
        hf = np.nextafter(np.maximum(h[ok], 0).astype(np.float32), np.float32(-np.inf)).astype(np.float64)
        assert np.all(np.minimum(hf, np.maximum(h[ok], 0)) <= d[i][ok] * (1 + 1e-12) + 1e-9)

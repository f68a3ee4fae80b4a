"""Size-independent properties at a scale the Python oracle cannot reach (SURVEY.md 8d 'parity checks
reported next to throughput'): conservation of agents and link volumes, volume recount, route
validity and optimality against an independent Dijkstra (scipy), and run-to-run determinism."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HOURLY = np.array([0.05, 0.02, 0.01, 0.01, 0.02, 0.05, 0.15, 0.25, 0.35, 0.20, 0.15, 0.18,
                   0.22, 0.18, 0.16, 0.20, 0.25, 0.30, 0.35, 0.25, 0.15, 0.12, 0.08, 0.06])


def _run(lg, N, seed, ticks, model="activity", blocks=3, options=()):
    from drift_b200 import demand
    from drift_b200.engine import Engine, ROUTER_BATCHED

    rng = np.random.default_rng(seed)
    start = rng.integers(0, lg.V, size=N).astype(np.int32)
    _, dst = demand.MODELS[model](rng, lg, start, 4 + 2 * blocks)
    dst = dst.reshape(N, -1)
    eng = Engine(lg, N, route_arena_links=N * 600)
    eng.set_demand(seed, HOURLY * 3, start, chain_capacity=4, router=ROUTER_BATCHED)
    eng.set_option("route_window", ticks)
    for name, value in options:
        eng.set_option(name, value)
    eng.push_trips(dst[:, :4])
    t = 2 * 3600.0
    out = []
    for b in range(blocks):
        eng.push_trips(dst[:, 4 + 2 * b: 6 + 2 * b])
        eng.run(ticks, 1.0, t)
        t += ticks
        occ = eng.occupancy()
        a = eng.agents(fields=("state", "dist", "speed", "elen", "cur_link", "route_idx", "trip_count"))
        dep = eng.drain_departures()
        arr = eng.drain_arrivals()
        out.append((occ, a, dep, arr, eng.recount_occupancy(), eng.stats()))
    return eng, out


def test_conservation_and_route_optimality():
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra
    from drift_b200.engine import ROUTE_OK, ROUTE_TRIVIAL
    from drift_b200.synth import road_graph

    lg = road_graph(60_000, seed=3)
    N = 200_000
    eng, out = _run(lg, N, seed=11, ticks=150)
    try:
        n_dep_ok = n_arr = 0
        for occ, a, dep, arr, recount, st in out:
            moving = a["state"] == 1
            assert np.array_equal(occ, recount), "volume recount differs from the event-maintained volumes"
            assert occ.sum() == moving.sum() == st["moving_agents"] == st["total_agents_on_roads"]
            assert (occ >= 0).all()
            assert np.all(a["cur_link"][moving] >= 0) and np.all(a["cur_link"][~moving] == -1)
            assert np.all(a["dist"][moving] < a["elen"][moving]) and np.all(a["dist"][moving] >= 0)
            assert np.all(a["speed"][moving] > 0)
            agent, tick, src, dst, status = dep
            n_dep_ok += int((status == ROUTE_OK).sum())
            n_arr += len(arr[0])
            assert set(np.unique(status).tolist()) <= {0, 1, 2}
            assert np.all(src[status == ROUTE_TRIVIAL] == dst[status == ROUTE_TRIVIAL])
            assert n_dep_ok - n_arr == moving.sum(), "departures - arrivals != agents on the road"
        assert n_dep_ok > 20_000 and n_arr > 1_000
        # routes of agents still under way: valid walks, optimal cost (independent Dijkstra, 1e-9 relative)
        occ, a, dep, arr, recount, st = out[-1]
        W = csr_matrix((lg.fwd_w, lg.fwd_col, lg.fwd_rowptr), shape=(lg.V, lg.V))
        movers = np.nonzero(a["state"] == 1)[0]
        rng = np.random.default_rng(0)
        sample = rng.choice(movers, size=40, replace=False)
        for i in sample.tolist():
            links = eng.agent_route(i)
            assert np.all(lg.link_v[links[:-1]] == lg.link_u[links[1:]]), "route is not a walk"
            assert links[a["route_idx"][i]] == a["cur_link"][i]
            s, t = int(lg.link_u[links[0]]), int(lg.link_v[links[-1]])
            cost = 0.0
            for l in links:
                cost += lg.link_t0[l]
            best = dijkstra(W, directed=True, indices=s)[t]
            assert abs(cost - best) <= 1e-9 * best
    finally:
        eng.close()


def test_run_to_run_determinism():
    from drift_b200.synth import road_graph

    lg = road_graph(20_000, seed=5)
    outs = []
    for _ in range(2):
        eng, out = _run(lg, 60_000, seed=7, ticks=100, model="hub", blocks=2)
        eng.close()
        outs.append(out)
    for (o1, a1, d1, r1, _, _), (o2, a2, d2, r2, _, _) in zip(*outs):
        assert np.array_equal(o1, o2)
        for k in a1:
            assert np.array_equal(a1[k], a2[k]), k  # bit-identical fp64 state
        assert sorted(zip(*[x.tolist() for x in d1])) == sorted(zip(*[x.tolist() for x in d2]))
        assert sorted(zip(*[x.tolist() for x in r1])) == sorted(zip(*[x.tolist() for x in r2]))


def test_route_arena_compaction_is_transparent():
    """A tiny arena forces several compactions; results equal a run with a large arena."""
    from drift_b200 import demand
    from drift_b200.engine import Engine, ROUTER_BATCHED
    from drift_b200.synth import road_graph

    lg = road_graph(4000, seed=9)
    N = 6000
    outs = []
    for arena in (1 << 24, 1_200_000):
        rng = np.random.default_rng(2)
        start = rng.integers(0, lg.V, size=N).astype(np.int32)
        eng = Engine(lg, N, route_arena_links=arena)
        eng.set_demand(5, np.full(24, 0.9), start, chain_capacity=4, router=ROUTER_BATCHED)
        eng.set_option("route_window", 40)
        eng.set_option("arena_compact_frac", 0.5)
        t = 0.0
        snap = []
        for b in range(12):
            _, cand = demand.random_chains(rng, lg.V, start, 2)
            eng.push_trips(cand.reshape(N, 2))
            eng.run(40, 1.0, t)
            t += 40
            snap.append((eng.occupancy().copy(), eng.agents(fields=("state", "dist", "speed", "route_idx"))))
            eng.drain_departures(), eng.drain_arrivals()
        st = eng.stats()
        eng.close()
        outs.append((snap, st))
    assert outs[0][1]["arena_compactions"] == 0 and outs[1][1]["arena_compactions"] >= 2
    for (o1, a1), (o2, a2) in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(o1, o2)
        for k in a1:
            assert np.array_equal(a1[k], a2[k])


def test_one_wave_tiles_are_bit_identical():
    """950 k agents = 928 plain tiles of 1 024: more than the 6 x 148 resident blocks of the persistent tick kernel can
    take in one wave, so the engine widens the tiles (1 024 agents four per thread + 64 one per thread = 874 tiles).
    Same results, bit for bit, as plain tiles run in two waves (one_wave_tiles = 0)."""
    from drift_b200.synth import road_graph

    lg = road_graph(20_000, seed=5)
    outs = []
    for opts in ((), (("one_wave_tiles", 0),)):
        eng, out = _run(lg, 950_000, seed=11, ticks=60, model="hub", blocks=2, options=opts)
        eng.close()
        outs.append(out)
    for (o1, a1, d1, r1, h1, s1), (o2, a2, d2, r2, h2, s2) in zip(*outs):
        assert np.array_equal(o1, o2) and np.array_equal(h1, o1) and s1["moving_agents"] > 10_000
        for k in a1:
            assert np.array_equal(a1[k], a2[k]), k
        assert sorted(zip(*[x.tolist() for x in d1])) == sorted(zip(*[x.tolist() for x in d2]))
        assert sorted(zip(*[x.tolist() for x in r1])) == sorted(zip(*[x.tolist() for x in r2]))

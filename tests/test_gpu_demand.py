"""Device-driven demand (drift_run): departures decided, routed and injected on the device.

The oracle port replays the device's departure log (agent, tick, src, dst) with its own
NetworkX-exact router and sequential dynamics; link volumes, arrivals and fp64 state must match
bit for bit, and the departure decisions themselves must equal the Philox law restated in numpy.
"""
import numpy as np
import pytest

from oracle import golden
from oracle.philox import uniform53
from oracle.port import DEFAULT_HOURLY, PortSim, travel_probability

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,router,window,persistent,lookahead", [
    ("g40_congested", 0, 25, 1, 0), ("g100_random_s0", 0, 300, 1, 0), ("multi_random_s0", 0, 7, 1, 0),
    ("g40_congested", 0, 1, 1, 0), ("g40_congested", 0, 25, 0, 0), ("multi_random_s0", 0, 300, 0, 0),
    # lookahead planning: routes only for the departures the coming window's draws allow
    ("g40_congested", 0, 25, 1, 1), ("g100_random_s0", 0, 50, 1, 1), ("multi_random_s0", 0, 7, 0, 1),
    # the batched router (trees / bounded searches + tie detector -> NetworkX-exact fallback) on the bundled,
    # tie-heavy graphs: the same link sequences as the reference's router, hence the same dynamics
    ("g40_congested", 1, 25, 1, 0), ("g100_random_s0", 1, 300, 1, 0), ("multi_random_s0", 1, 7, 1, 1),
    ("musicnet_random_s0", 1, 50, 1, 0)])
def test_device_demand_matches_port(name, router, window, persistent, lookahead):
    from drift_b200.engine import Engine
    from drift_b200.graph import lower_graph

    case = golden.GoldenCase(name)
    lg = lower_graph(case.nx_graph())
    pg = case.plain_graph()
    rng = np.random.default_rng(7)
    N, V = 3000, lg.V
    start = rng.integers(0, V, size=N).astype(np.int32)
    per_agent = 6
    cands = rng.integers(0, V, size=(N, per_agent)).astype(np.int32)
    hourly = {h: 0.9 for h in range(24)}
    h24 = np.array([hourly[h] for h in range(24)])
    seed, delta, blocks, ticks_per_block = 1234567, 1.0, 12, 50
    with Engine(lg, N, alpha=0.15, beta=4.0) as eng:
        eng.set_demand(seed, h24, start, chain_capacity=per_agent, router=router)
        eng.set_option("route_window", window)
        eng.set_option("persistent", persistent)
        eng.set_option("plan_lookahead", lookahead)
        eng.push_trips(cands)
        occs, t = [], 0.0
        dep_log = []
        arr_log = []
        for b in range(blocks):
            eng.run(ticks_per_block, delta, t)
            for _ in range(ticks_per_block):
                t += delta
            occs.append(eng.occupancy().copy())
            a, tk, s, d, st = eng.drain_departures()
            dep_log.extend(zip(tk.tolist(), a.tolist(), s.tolist(), d.tolist(), st.tolist()))
            aa, at = eng.drain_arrivals()
            arr_log.extend(zip(at.tolist(), aa.tolist()))
        final = eng.agents(fields=("state", "dist", "speed", "route_idx", "trip_count", "cur_node"))
    assert len(dep_log) > 500
    # port replay of the device's departures
    trace = {}
    for tk, a, s, d, st in dep_log:
        trace.setdefault(tk, {})[a] = (s, d, "route")
    sim = PortSim(pg, N, hourly=hourly, accel=10, hours=1, departure_trace={}, init_nodes=start)

    class _Trace(dict):
        def __missing__(self, k):
            return {}

    sim.trace = _Trace(trace)
    waiting_before = {}
    for k in range(1, blocks * ticks_per_block + 1):
        waiting = np.array([a.idx for a in sim.agents if a.state == 0 and a.trip_count < per_agent], dtype=np.uint32)
        # Philox law: exactly the waiting agents with u < p depart
        p = travel_probability(sim.t + delta, hourly) * delta / 3600
        u = uniform53(seed, waiting, k)
        want = set(waiting[u < p].tolist())
        got = set(trace.get(k, {}).keys())
        assert got == want, "departure set differs at tick %d" % k
        sim.tick()
        if k % ticks_per_block == 0:
            assert np.array_equal(np.asarray(sim.occ, dtype=np.int32), occs[k // ticks_per_block - 1]), k
    assert sorted(arr_log) == sim.arrivals
    st, d, sp, pi = sim.snapshot()
    assert np.array_equal(final["state"], st)
    mv = st == 1
    assert np.array_equal(final["dist"][mv], d[mv]) and np.array_equal(final["speed"][mv], sp[mv])
    assert np.array_equal(final["route_idx"][mv], pi[mv])
    assert np.array_equal(final["trip_count"], np.array([a.trip_count for a in sim.agents]))
    # statuses in the log agree with the port's router
    port_status = {(tk, a): (2 if path is None else (1 if len(path) == 1 else 0)) for tk, a, s, t_, path in sim.departures}
    for tk, a, s, d_, stt in dep_log:
        assert port_status[(tk, a)] == stt

"""CUDA tick (through the C ABI) against the fixtures recorded from the live reference.

The recorded departures (agent, tick, NetworkX path) are replayed; everything downstream of a
departure -- crossing ticks, per-tick link volumes, fp64 entry speeds and positions, arrival
ticks -- must be bit-equal to the reference.
"""
import numpy as np
import pytest

from oracle import golden

pytestmark = pytest.mark.gpu
CASES = golden.case_names()


def _replay(case, use_device_router, persistent=True):
    from drift_b200.engine import Engine, ROUTE_OK, ROUTE_TRIVIAL, ROUTE_NOPATH
    from drift_b200.graph import lower_graph

    lg = lower_graph(case.nx_graph())
    delta = 0.1 * case.accel
    trace = case.departure_trace()
    occ_ref = case.occupancy()
    snaps = case.snapshots()
    with Engine(lg, case.N, alpha=case.alpha, beta=case.beta) as eng:
        eng.enable_entry_trace(1 << 18)
        eng.set_option("persistent", 1 if persistent else 0)
        arr, ent = [], []

        def drain():
            agent, tick = eng.drain_arrivals()
            arr.extend(zip(tick.tolist(), agent.tolist()))
            et, ea, el, es = eng.drain_entries()
            ent.extend(zip(et.tolist(), ea.tolist(), el.tolist(), es.tolist()))

        for k in range(1, case.ticks + 1):
            deps = trace.get(k)
            if deps:
                agents = sorted(deps)
                if use_device_router:
                    src = [deps[a][0] for a in agents]
                    dst = [deps[a][1] for a in agents]
                    status, nl, links = eng.route(src, dst)
                    for a, st, ln in zip(agents, status, links):
                        path = deps[a][2]
                        if path is None:
                            assert st == ROUTE_NOPATH
                        elif len(path) == 1:
                            assert st == ROUTE_TRIVIAL
                        else:
                            assert st == ROUTE_OK
                            assert lg.links_to_nodes(path[0], ln) == path, "route differs from NetworkX"
                    eng.commit_departures(agents)
                else:
                    routes = [lg.nodes_to_links(deps[a][2]) if deps[a][2] is not None else [] for a in agents]
                    eng.submit_departures(agents, routes)
            eng.step(1, delta)
            occ = eng.occupancy()
            assert np.array_equal(occ, occ_ref[k - 1]), "link volumes differ at tick %d" % k
            if k in snaps:
                a = eng.agents()
                st, d, sp, pi = snaps[k]
                assert np.array_equal(a["state"], st)
                mv = st == 1
                assert np.array_equal(a["dist"][mv], d[mv])
                assert np.array_equal(a["speed"][mv], sp[mv])
                assert np.array_equal(a["route_idx"][mv], pi[mv])
                assert np.array_equal(eng.recount_occupancy(), occ)
                drain()  # the rings hold max(N, 4096) records
        drain()
        assert sorted(arr) == case.arrivals()
        want = sorted((t, a, l, s) for t, a, l, s, _ in case.entries())
        assert sorted(ent) == want
        s = eng.stats()
        assert s["arrivals_total"] == len(case.arrivals())
        assert s["moving_agents"] == int(case.z["final_state"].sum())
        assert s["total_agents_on_roads"] == int(occ_ref[-1].sum())


@pytest.mark.parametrize("name", CASES)
def test_replay_host_routes(name):
    _replay(golden.GoldenCase(name), use_device_router=False)


@pytest.mark.parametrize("name", [n for n in CASES if n in (
    "g100_random_s0", "g40_congested_bpr", "chpn_random_nopath", "multi_random_s0", "g100_activity_s0")])
def test_replay_device_routes(name):
    _replay(golden.GoldenCase(name), use_device_router=True)


@pytest.mark.parametrize("name", ["g40_congested", "multi_random_s0", "g100_random_accel7"])
def test_replay_separate_kernels(name):
    """Same tick as three separate kernels (the path the multi-rank exchange and the profiler use)."""
    _replay(golden.GoldenCase(name), use_device_router=False, persistent=False)

"""Hub-label router (contraction hierarchy + labels built on the device) against live NetworkX / scipy,
against the tree / bounded-search router, and inside a device-demand run against the oracle port."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _graph(kind):
    from drift_b200.graph import from_links
    from drift_b200.synth import rgg_graph, road_graph

    if kind == "road":
        return road_graph(3000, seed=5)
    if kind == "rgg":
        return rgg_graph(2500, seed=2)
    # one-way ring with chords, a source-only and a sink-only node: unreachable pairs exist
    rng = np.random.default_rng(11)
    n = 400
    u = list(range(n))
    v = [(i + 1) % n for i in range(n)]
    for _ in range(300):
        a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
        if a != b and (a, b) not in set(zip(u, v)):
            u.append(a)
            v.append(b)
    u += [n, 7]          # node n only leaves, node n + 1 is only entered
    v += [3, n + 1]
    pairs = sorted(set(zip(u, v)))
    u = np.array([p[0] for p in pairs])
    v = np.array([p[1] for p in pairs])
    length = rng.uniform(40.0, 400.0, size=len(u))
    speed = rng.choice([30.0, 50.0, 70.0], size=len(u))
    return from_links(n + 2, u, v, length, speed)


@pytest.mark.parametrize("kind", ["road", "rgg", "oneway"])
def test_label_router_equals_networkx(kind):
    """Paths (node for node), statuses and fp64 costs of the label router == nx.shortest_path on graphs with
    continuous travel times (unique shortest paths), incl. trivial, invalid and unreachable pairs."""
    import networkx as nx

    from drift_b200 import hierarchy
    from drift_b200.engine import Engine, ROUTE_NOPATH, ROUTE_OK, ROUTE_TRIVIAL, ROUTER_BATCHED
    from drift_b200.graph import to_networkx

    lg = _graph(kind)
    G = to_networkx(lg)
    hier = hierarchy.build(lg)
    rng = np.random.default_rng(3)
    Q = 3000
    src = rng.integers(0, lg.V, size=Q).astype(np.int32)
    dst = rng.integers(0, lg.V, size=Q).astype(np.int32)
    dst[:20] = src[:20]          # trivial
    src[20:25] = -1              # node missing (lib/agent.py:87-91)
    dst[25:30] = lg.V + 5
    with Engine(lg, 8, route_arena_links=1 << 23) as eng:
        eng.set_hierarchy(hier)
        ls = eng.label_stats()
        assert ls["label_entries"] > 2 * lg.V
        status, nl, links = eng.route(src, dst, router=ROUTER_BATCHED)
        costs = eng.route_costs(Q)
        # the same batch through the tree / bounded-search router
        eng.set_option("use_labels", 0)
        status_t, nl_t, links_t = eng.route(src, dst, router=ROUTER_BATCHED)
        costs_t = eng.route_costs(Q)
    n_ok = n_nopath = 0
    for q in range(Q):
        s, t = int(src[q]), int(dst[q])
        if s < 0 or t < 0 or s >= lg.V or t >= lg.V:
            assert status[q] == ROUTE_NOPATH
            continue
        if s == t:
            assert status[q] == ROUTE_TRIVIAL and nl[q] == 0
            continue
        try:
            want = nx.shortest_path(G, s, t, weight="travel_time")
        except nx.NetworkXNoPath:
            assert status[q] == ROUTE_NOPATH, (s, t)
            n_nopath += 1
            continue
        assert status[q] == ROUTE_OK, (s, t)
        assert lg.links_to_nodes(s, links[q]) == want, (s, t)
        c = 0.0
        for l in links[q]:
            c += lg.link_t0[l]
        assert costs[q] == c
        n_ok += 1
    assert n_ok > Q // 2
    if kind == "oneway":
        assert n_nopath > 0
    assert np.array_equal(status, status_t) and np.array_equal(nl, nl_t) and np.array_equal(costs, costs_t)
    for a, b in zip(links, links_t):
        assert np.array_equal(a, b)


def test_label_router_large_batch_against_scipy():
    """A 50 k-node road graph, 200 k queries: every cost equals scipy's Dijkstra distance within 1e-9 relative
    and every route is a contiguous walk from source to target."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    from drift_b200 import hierarchy
    from drift_b200.engine import Engine, ROUTE_OK, ROUTER_BATCHED
    from drift_b200.synth import road_graph

    lg = road_graph(50_000, seed=1)
    hier = hierarchy.build(lg)
    rng = np.random.default_rng(5)
    Q = 200_000
    sources = rng.integers(0, lg.V, size=16)
    src = sources[rng.integers(0, 16, size=Q)].astype(np.int32)
    dst = rng.integers(0, lg.V, size=Q).astype(np.int32)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    Q = len(src)
    with Engine(lg, 8, route_arena_links=1 << 27) as eng:
        eng.set_hierarchy(hier)
        status, nl, links = eng.route(src, dst, router=ROUTER_BATCHED)
        costs = eng.route_costs(Q)
    assert (status == ROUTE_OK).all()
    W = csr_matrix((lg.fwd_w, lg.fwd_col, lg.fwd_rowptr), shape=(lg.V, lg.V))
    D = dijkstra(W, directed=True, indices=sources)
    row = {int(s): k for k, s in enumerate(sources)}
    want = np.array([D[row[int(s)], int(t)] for s, t in zip(src, dst)])
    assert np.all(np.abs(costs - want) <= 1e-9 * np.maximum(want, 1.0))
    for q in rng.integers(0, Q, size=2000):
        ln = links[q]
        assert lg.link_u[ln[0]] == src[q] and lg.link_v[ln[-1]] == dst[q]
        assert np.array_equal(lg.link_v[ln[:-1]], lg.link_u[ln[1:]])


@pytest.mark.parametrize("lookahead", [0, 1])
def test_device_demand_with_labels_equals_trees_and_port(lookahead):
    """Device-driven demand routed by hub labels: identical, bit for bit, to the same run routed by trees /
    bounded searches, and to the oracle port replaying the device's departure log."""
    from drift_b200 import hierarchy
    from drift_b200.engine import Engine, ROUTER_BATCHED
    from drift_b200.synth import road_graph
    from oracle.port import PortSim, plain_from_arrays

    lg = road_graph(2000, seed=9)
    hier = hierarchy.build(lg)
    rng = np.random.default_rng(7)
    N, V = 6000, lg.V
    start = rng.integers(0, V, size=N).astype(np.int32)
    cands = rng.integers(0, V, size=(N, 6)).astype(np.int32)
    hourly = {h: 0.9 for h in range(24)}
    h24 = np.full(24, 0.9)
    blocks, tpb = 6, 50

    def run(use_labels):
        with Engine(lg, N, route_arena_links=N * 400) as eng:
            if use_labels:
                eng.set_hierarchy(hier)
            eng.set_demand(99, h24, start, chain_capacity=6, router=ROUTER_BATCHED)
            eng.set_option("route_window", 25)
            eng.set_option("plan_lookahead", lookahead)
            eng.push_trips(cands)
            occs, deps, arrs, t = [], [], [], 0.0
            for _ in range(blocks):
                eng.run(tpb, 1.0, t)
                t += tpb
                occs.append(eng.occupancy().copy())
                a, tk, s, d, st = eng.drain_departures()
                deps.extend(zip(tk.tolist(), a.tolist(), s.tolist(), d.tolist(), st.tolist()))
                aa, at = eng.drain_arrivals()
                arrs.extend(zip(at.tolist(), aa.tolist()))
            final = eng.agents(fields=("state", "dist", "speed", "route_idx", "trip_count"))
            if use_labels:
                assert eng.label_stats()["queries"] > 0
            assert eng.stats()["fallback_routes"] == 0
        return occs, sorted(deps), sorted(arrs), final

    occ_l, dep_l, arr_l, fin_l = run(True)
    occ_t, dep_t, arr_t, fin_t = run(False)
    assert dep_l == dep_t and arr_l == arr_t and len(arr_l) > 200
    for a, b in zip(occ_l, occ_t):
        assert np.array_equal(a, b)
    for k in fin_l:
        assert np.array_equal(fin_l[k], fin_t[k]), k
    # oracle port (NetworkX-order router, sequential dynamics) on the device's departure log
    pg = plain_from_arrays(lg.nodes, lg.link_u, lg.link_v, lg.link_len, lg.link_v0, lg.link_t0, lg.link_cap,
                           lg.fwd_rowptr, lg.fwd_col, lg.fwd_w, lg.fwd_link, lg.fwd_minlen, lg.bwd_rowptr,
                           lg.bwd_col, lg.bwd_w)
    trace = {}
    for tk, a, s, d, st in dep_l:
        trace.setdefault(tk, {})[a] = (s, d, "route")
    sim = PortSim(pg, N, hourly=hourly, accel=10, hours=1, departure_trace={}, init_nodes=start)

    class _Trace(dict):
        def __missing__(self, k):
            return {}

    sim.trace = _Trace(trace)
    for k in range(1, blocks * tpb + 1):
        sim.tick()
        if k % tpb == 0:
            assert np.array_equal(np.asarray(sim.occ, dtype=np.int32), occ_l[k // tpb - 1]), k
    assert arr_l == sim.arrivals
    st, d, sp, pi = sim.snapshot()
    mv = st == 1
    assert np.array_equal(fin_l["state"], st)
    assert np.array_equal(fin_l["dist"][mv], d[mv]) and np.array_equal(fin_l["speed"][mv], sp[mv])
    assert np.array_equal(fin_l["route_idx"][mv], pi[mv])

"""Device router (K5a, NetworkX-exact) against recorded reference paths and live NetworkX."""
import random

import numpy as np
import pytest

from oracle import golden

pytestmark = pytest.mark.gpu
CASES = golden.case_names()


@pytest.mark.parametrize("name", CASES)
def test_recorded_paths(name):
    """Every path the reference took in the fixture, as one batch."""
    from drift_b200.engine import Engine, ROUTE_OK, ROUTE_TRIVIAL, ROUTE_NOPATH
    from drift_b200.graph import lower_graph

    case = golden.GoldenCase(name)
    lg = lower_graph(case.nx_graph())
    deps = case.departures()
    src = [d[2] for d in deps]
    dst = [d[3] for d in deps]
    with Engine(lg, 8, route_arena_links=1 << 22) as eng:
        status, nl, links = eng.route(src, dst)
        costs = eng.route_costs(len(src))
    for d, st, ln, c in zip(deps, status, links, costs):
        path = d[4]
        if path is None:
            assert st == ROUTE_NOPATH
        elif len(path) == 1:
            assert st == ROUTE_TRIVIAL
        else:
            assert st == ROUTE_OK
            assert lg.links_to_nodes(path[0], ln) == path
            want = 0.0
            for l in ln:
                want += lg.link_t0[l]
            assert c == want


def _tie_graph(kind, seed):
    import networkx as nx

    rng = random.Random(seed)
    if kind == "grid":  # uniform weights: ~98 % of queries have several shortest paths
        H = nx.grid_2d_graph(14, 14)
        G = nx.MultiDiGraph()
        nodes = list(H.nodes)
        rng.shuffle(nodes)
        G.add_nodes_from(nodes)
        edges = [(a, b) for a, b in H.edges] + [(b, a) for a, b in H.edges]
        rng.shuffle(edges)
        for a, b in edges:
            G.add_edge(a, b, travel_time=120.0, length=1000.0, speed_kph=30.0)
    elif kind == "lattice_weights":  # metre lengths / {30,50,70}: exact fp ties do occur
        G = nx.MultiDiGraph()
        n = 150
        G.add_nodes_from(range(n))
        for _ in range(600):
            a, b = rng.randrange(n), rng.randrange(n)
            ln = float(rng.choice([50, 100, 150, 200]))
            sp = float(rng.choice([30, 50, 70]))
            G.add_edge(a, b, length=ln, speed_kph=sp, travel_time=(ln / 1000.0) / sp * 3600)
    else:  # sparse digraph with unreachable pairs, parallel edges and self-loops
        G = nx.MultiDiGraph()
        n = 120
        G.add_nodes_from(range(n))
        for _ in range(260):
            a, b = rng.randrange(n), rng.randrange(n)
            G.add_edge(a, b, length=100.0, speed_kph=30.0, travel_time=float(rng.choice([1, 2, 3])))
        for _ in range(40):
            a, b, _k = rng.choice(list(G.edges(keys=True)))
            G.add_edge(a, b, length=80.0, speed_kph=30.0, travel_time=float(rng.choice([1, 2, 3])))
    return G


@pytest.mark.parametrize("kind", ["grid", "lattice_weights", "sparse"])
def test_live_networkx_tie_breaking(kind):
    """nx.shortest_path (3.6.1, installed on the box) vs the device router on tie-heavy graphs."""
    import networkx as nx
    from drift_b200.engine import Engine, ROUTE_OK, ROUTE_TRIVIAL, ROUTE_NOPATH
    from drift_b200.graph import lower_graph

    G = _tie_graph(kind, seed=5)
    lg = lower_graph(G)
    rng = random.Random(9)
    nodes = list(G.nodes)
    Q = 1500
    pairs = [(rng.randrange(len(nodes)), rng.randrange(len(nodes))) for _ in range(Q)]
    with Engine(lg, 8, route_arena_links=1 << 22) as eng:
        status, nl, links = eng.route([p[0] for p in pairs], [p[1] for p in pairs])
    n_ok = 0
    for (s, t), st, ln in zip(pairs, status, links):
        try:
            want = nx.shortest_path(G, nodes[s], nodes[t], weight="travel_time")
        except nx.NetworkXNoPath:
            assert st == ROUTE_NOPATH
            continue
        if s == t:
            assert st == ROUTE_TRIVIAL and want == [nodes[s]]
            continue
        assert st == ROUTE_OK
        got = [nodes[i] for i in lg.links_to_nodes(s, ln)]
        assert got == want
        n_ok += 1
    assert n_ok > Q // 3


def _unique_path_graph(seed, n=400, m=1400):
    """Continuous random weights: shortest paths are unique."""
    import networkx as nx

    rng = random.Random(seed)
    G = nx.MultiDiGraph()
    G.add_nodes_from(range(n))
    for i in range(n):
        G.add_edge(i, (i + 1) % n, length=rng.uniform(50, 300), speed_kph=50.0)
    for _ in range(m):
        a, b = rng.randrange(n), rng.randrange(n)
        if a != b and not G.has_edge(a, b):
            G.add_edge(a, b, length=rng.uniform(50, 300), speed_kph=rng.choice([30.0, 50.0, 70.0]))
    for u, v, k, d in G.edges(keys=True, data=True):
        d["travel_time"] = (d["length"] / 1000.0) / d["speed_kph"] * 3600
    for i in range(5):  # a few nodes nobody can reach / leave
        G.add_node(n + i)
    G.add_edge(n, 0, length=100.0, speed_kph=30.0, travel_time=12.0)
    G.add_edge(1, n + 1, length=100.0, speed_kph=30.0, travel_time=12.0)
    return G


def test_batched_router_unique_paths_match_networkx():
    import networkx as nx
    from drift_b200.engine import Engine, ROUTE_OK, ROUTE_TRIVIAL, ROUTE_NOPATH, ROUTER_BATCHED, ROUTER_NX_EXACT
    from drift_b200.graph import lower_graph

    G = _unique_path_graph(3)
    lg = lower_graph(G)
    rng = random.Random(4)
    nodes = list(G.nodes)
    Q = 3000
    dests = [rng.randrange(len(nodes)) for _ in range(60)]
    pairs = [(rng.randrange(len(nodes)), rng.choice(dests)) for _ in range(Q)]
    src, dst = [p[0] for p in pairs], [p[1] for p in pairs]
    with Engine(lg, 8, route_arena_links=1 << 22) as eng:
        st_b, nl_b, links_b = eng.route(src, dst, router=ROUTER_BATCHED)
        cost_b = eng.route_costs(Q)
        st_x, nl_x, links_x = eng.route(src, dst, router=ROUTER_NX_EXACT)
        cost_x = eng.route_costs(Q)
        # second batch hits the tree cache
        st_c, nl_c, links_c = eng.route(src[:500], dst[:500], router=ROUTER_BATCHED)
    assert np.array_equal(st_b, st_x)
    n_ok = 0
    for q, ((s, t), st) in enumerate(zip(pairs, st_b)):
        try:
            want = nx.shortest_path(G, nodes[s], nodes[t], weight="travel_time")
        except nx.NetworkXNoPath:
            assert st == ROUTE_NOPATH
            continue
        if s == t:
            assert st == ROUTE_TRIVIAL
            continue
        assert st == ROUTE_OK
        assert [nodes[i] for i in lg.links_to_nodes(s, links_b[q])] == want  # unique shortest path
        assert links_b[q].tolist() == links_x[q].tolist()
        assert abs(cost_b[q] - cost_x[q]) <= 1e-9 * abs(cost_x[q])  # tolerance stated by north_star
        if q < 500:
            assert links_c[q].tolist() == links_b[q].tolist()
        n_ok += 1
    assert n_ok > Q // 2


@pytest.mark.parametrize("bounded", [0, 8])
@pytest.mark.parametrize("name", ["g100_random_s0", "musicnet_random_s0", "multi_random_s0"])
def test_batched_router_equals_networkx_on_tie_graphs(name, bounded):
    """Bundled graphs carry one travel_time value: most queries have several shortest paths and
    the reference's answer depends on NetworkX's search order.  The batched router flags every query with a
    second (nearly) tight hop and re-runs it through the NetworkX-exact router, so its link sequences equal
    live nx.shortest_path -- from cached trees (bounded=0) and from bounded searches (bounded=8) alike."""
    import networkx as nx
    from drift_b200.engine import Engine, ROUTE_NOPATH, ROUTE_OK, ROUTE_TRIVIAL, ROUTER_BATCHED, ROUTER_NX_EXACT
    from drift_b200.graph import lower_graph

    case = golden.GoldenCase(name)
    G = case.nx_graph()
    lg = lower_graph(G)
    nodes = list(G.nodes)
    rng = random.Random(1)
    Q = 800
    src = [rng.randrange(lg.V) for _ in range(Q)]
    dst = [rng.randrange(lg.V) for _ in range(Q)]
    with Engine(lg, 8, route_arena_links=1 << 22) as eng:
        eng.set_option("bounded_max", bounded)
        st_b, nl_b, links_b = eng.route(src, dst, router=ROUTER_BATCHED)
        cost_b = eng.route_costs(Q)
        ties = eng.stats()["tie_fallbacks"]
        st_x, nl_x, links_x = eng.route(src, dst, router=ROUTER_NX_EXACT)
        cost_x = eng.route_costs(Q)
        # detector off: still optimal walks, but not necessarily NetworkX's
        eng.set_option("detect_ties", 0)
        st_o, nl_o, links_o = eng.route(src, dst, router=ROUTER_BATCHED)
        cost_o = eng.route_costs(Q)
    assert ties > (Q // 10 if name != "multi_random_s0" else 0)  # these graphs do tie (the sparse multigraph less often)
    assert np.array_equal(st_b, st_x) and np.array_equal(nl_b, nl_x) and np.array_equal(cost_b, cost_x)
    assert np.array_equal(st_o, st_x)
    ok = st_x == ROUTE_OK
    assert np.all(np.abs(cost_o[ok] - cost_x[ok]) <= 1e-9 * np.abs(cost_x[ok]))
    for q in range(Q):
        s, t = src[q], dst[q]
        if s == t:
            assert st_b[q] == ROUTE_TRIVIAL
            continue
        try:
            want = nx.shortest_path(G, nodes[s], nodes[t], weight="travel_time")
        except nx.NetworkXNoPath:
            assert st_b[q] == ROUTE_NOPATH
            continue
        assert [nodes[k] for k in lg.links_to_nodes(s, links_b[q])] == want, (s, t)
        walk = lg.links_to_nodes(s, links_o[q])
        assert walk[0] == s and walk[-1] == t and all(lg.link_u[l] == a for l, a in zip(links_o[q], walk[:-1]))


def test_tree_cache_eviction():
    """A cache that holds 8 trees serves batches with many more destinations over time."""
    import networkx as nx
    from drift_b200.engine import Engine, ROUTE_OK, ROUTER_BATCHED
    from drift_b200.graph import lower_graph

    G = _unique_path_graph(11, n=200, m=600)
    lg = lower_graph(G)
    nodes = list(G.nodes)
    rng = random.Random(2)
    with Engine(lg, 8, route_arena_links=1 << 22) as eng:
        eng.set_option("tree_cache_bytes", 8 * 8 * lg.V)  # 8 slots of 8 B per node
        for rnd in range(12):
            dests = [rng.randrange(200) for _ in range(6)] + [0, 1]  # 0 and 1 stay hot, others churn
            pairs = [(rng.randrange(200), rng.choice(dests)) for _ in range(300)]
            st, nl, links = eng.route([p[0] for p in pairs], [p[1] for p in pairs], router=ROUTER_BATCHED)
            for (s, t), stt, ln in zip(pairs, st, links):
                if s == t:
                    continue
                want = nx.shortest_path(G, nodes[s], nodes[t], weight="travel_time")
                assert stt == ROUTE_OK and [nodes[i] for i in lg.links_to_nodes(s, ln)] == want
        from drift_b200._capi import DriftError
        eng.set_option("bounded_max", 0)  # full trees only: every destination of a batch needs a slot
        with pytest.raises(DriftError):  # 20 distinct destinations cannot fit 8 slots in one batch
            eng.route(list(range(20, 40)), list(range(100, 120)), router=ROUTER_BATCHED)


@pytest.mark.parametrize("bounded_max,landmarks", [(8, 0), (0, 0), (8, 6), (1, 3)])
def test_batched_router_cold_destinations(bounded_max, landmarks):
    """Random (s, t) pairs: most destinations are wanted by a handful of queries, so the builder
    stops as soon as their sources are settled and writes the routes itself (bounded search)."""
    import networkx as nx
    from drift_b200.engine import Engine, ROUTE_OK, ROUTE_TRIVIAL, ROUTE_NOPATH, ROUTER_BATCHED

    from drift_b200.graph import lower_graph

    G = _unique_path_graph(21, n=600, m=1500)
    lg = lower_graph(G)
    nodes = list(G.nodes)
    rng = random.Random(8)
    Q = 2500
    pairs = [(rng.randrange(len(nodes)), rng.randrange(len(nodes))) for _ in range(Q)]
    pairs += [(7, 7), (len(nodes) - 4, 3), (3, len(nodes) - 4)]  # trivial, and the unreachable / dead-end extras
    with Engine(lg, 8, route_arena_links=1 << 22) as eng:
        eng.set_option("bounded_max", bounded_max)
        if landmarks:
            assert len(eng.set_landmarks(landmarks)) == landmarks  # goal-directed (ALT) bounded searches
        for rnd in range(2):
            st, nl, links = eng.route([p[0] for p in pairs], [p[1] for p in pairs], router=ROUTER_BATCHED)
            costs = eng.route_costs(len(pairs))
            for q, ((s, t), stt) in enumerate(zip(pairs, st)):
                if s == t:
                    assert stt == ROUTE_TRIVIAL
                    continue
                try:
                    want = nx.shortest_path(G, nodes[s], nodes[t], weight="travel_time")
                except nx.NetworkXNoPath:
                    assert stt == ROUTE_NOPATH, (q, s, t)
                    continue
                assert stt == ROUTE_OK, (q, s, t)
                assert [nodes[i] for i in lg.links_to_nodes(s, links[q])] == want
                c = 0.0
                for l in links[q]:
                    c += lg.link_t0[l]
                assert costs[q] == c


@pytest.mark.parametrize("sides", [1, 2])
def test_batched_router_hot_sources(sides):
    """Return legs leave a few places (depots, hubs) for scattered targets: the batched router then
    builds a forward tree out of the source and walks it back from each destination.  Same routes
    as NetworkX on a unique-path graph, incl. mixed batches and the cache hits of a second batch."""
    import networkx as nx
    from drift_b200.engine import Engine, ROUTE_OK, ROUTE_TRIVIAL, ROUTE_NOPATH, ROUTER_BATCHED
    from drift_b200.graph import lower_graph

    G = _unique_path_graph(31, n=500, m=1400)
    lg = lower_graph(G)
    nodes = list(G.nodes)
    rng = random.Random(5)
    depots = [3, 77, 402]
    pools = [10, 11, 250]
    pairs = [(rng.choice(depots), rng.randrange(500)) for _ in range(900)]       # hot source, cold target
    pairs += [(rng.randrange(500), rng.choice(pools)) for _ in range(900)]       # cold source, hot target
    pairs += [(rng.choice(depots), rng.choice(pools)) for _ in range(100)]       # both hot
    pairs += [(rng.randrange(500), rng.randrange(500)) for _ in range(300)]      # both cold
    pairs += [(3, 3), (len(nodes) - 4, 3), (3, len(nodes) - 4)]
    rng.shuffle(pairs)
    with Engine(lg, 8, route_arena_links=1 << 23) as eng:
        eng.set_option("tree_sides", sides)
        eng.set_option("bounded_max", 8)
        for rnd in range(2):
            st, nl, links = eng.route([p[0] for p in pairs], [p[1] for p in pairs], router=ROUTER_BATCHED)
            costs = eng.route_costs(len(pairs))
            for q, ((s, t), stt) in enumerate(zip(pairs, st)):
                if s == t:
                    assert stt == ROUTE_TRIVIAL
                    continue
                try:
                    want = nx.shortest_path(G, nodes[s], nodes[t], weight="travel_time")
                except nx.NetworkXNoPath:
                    assert stt == ROUTE_NOPATH, (q, s, t)
                    continue
                assert stt == ROUTE_OK, (q, s, t)
                assert [nodes[i] for i in lg.links_to_nodes(s, links[q])] == want
                c = 0.0
                for l in links[q]:
                    c += lg.link_t0[l]
                assert costs[q] == c

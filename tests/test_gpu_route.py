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

"""TEST INFRASTRUCTURE ONLY -- records (source, target) sequences of the reference's s-t selectors.

    python oracle/make_golden_selectors.py          # needs /root/reference (build container only)

For every case the unmodified reference class (lib/st_selection.py) is built on a deterministic
graph, ``random.seed(seed)`` is applied, and a population of agents asks it for trips the way
lib/agent.py:47-53 does (first call with current_location None, then from the previous target).
The node *indices* (position in ``graph.nodes``) of the answers go to
``tests/golden/selectors.npz``; ``tests/test_fast_selection.py`` replays the same questions against
``drift_b200.fast_selection``.
"""
import os
import random
import sys

import networkx as nx
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("DRIFT_REFERENCE", "/root/reference")

CASES = {
    # name: (graph kind, model, constructor kwargs, seed, agents, calls)
    "hub_small": ("geo300", "hub", {}, 11, 40, 3000),
    "hub_large": ("geo2500", "hub", {"hub_trip_probability": 0.7, "hub_percentage": 0.05}, 12, 60, 3000),
    "zone_small": ("geo300", "zone", {}, 13, 40, 3000),
    "zone_sparse": ("line40", "zone", {"intra_zone_probability": 0.5}, 14, 10, 2000),
    "gravity_hops": ("geo300", "gravity", {}, 15, 40, 2000),
    "gravity_hops_beta": ("digeo200", "gravity", {"alpha": 2.0, "beta": 1.5}, 16, 30, 2000),
    "gravity_euclid": ("geo2500", "gravity", {"beta": 2.0}, 17, 60, 2500),
    "gravity_euclid_beta": ("geo2500", "gravity", {"alpha": 0.5, "beta": 2.5}, 18, 60, 1500),
    # the selector is built AFTER seeding: its pools are drawn with random.sample (lib/st_selection.py:732-900)
    "activity_small": ("geo300", "activity", {}, 19, 40, 3000),
    "activity_large": ("geo2500", "activity", {"agent_type_distributions": {"commuter": 0.25, "leisure": 0.25,
                                                                           "business": 0.25, "delivery": 0.25}},
                       20, 60, 3000),
    "activity_tiny": ("line40", "activity", {}, 21, 12, 1500),
}
SEED_BEFORE_BUILD = {"activity"}  # models whose constructor consumes the global generator


def make_graph(kind):
    """Deterministic test graphs with a 'pos' attribute (what lib/graph_loader.py guarantees)."""
    rng = np.random.default_rng(abs(hash(kind)) % 1000 if False else sum(map(ord, kind)))
    if kind == "line40":  # most zones of the 3 x 3 grid are empty or hold one node
        g = nx.MultiDiGraph()
        for i in range(40):
            g.add_node("n%d" % i, pos=(float(i), float(i % 2) * 0.01))
        for i in range(39):
            g.add_edge("n%d" % i, "n%d" % (i + 1))
            g.add_edge("n%d" % (i + 1), "n%d" % i)
        return g
    n = {"geo300": 300, "digeo200": 200, "geo2500": 2500}[kind]
    pts = rng.random((n, 2)) * 1000.0
    g = nx.MultiDiGraph()
    for i in range(n):
        g.add_node(i, pos=(float(pts[i, 0]), float(pts[i, 1])))
    from scipy.spatial import cKDTree

    _, nbr = cKDTree(pts).query(pts, k=4)
    for i in range(n):
        for j in nbr[i, 1:]:
            g.add_edge(i, int(j))
            if kind != "digeo200" or (i + int(j)) % 3:  # digeo200: some links are one-way -> unreachable pairs
                g.add_edge(int(j), i)
    return g


def ask(selector, n_agents, calls):
    """The questions lib/agent.py:47-53 asks, round-robin over the agents."""
    types = ["commuter", "delivery", "leisure", "business", "random"]
    where = [None] * n_agents
    trips = [0] * n_agents
    out = []
    for c in range(calls):
        a = c % n_agents
        s, t = selector.get_source_target(types[a % 5], where[a], trips[a])
        where[a] = t
        trips[a] += 1
        out.append((s, t))
    return out


def reference_selector(model, graph, kwargs):
    sys.path.insert(0, REF)
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    from lib import st_selection as ref

    cls = {"hub": ref.HubAndSpokeSelection, "zone": ref.ZoneBasedSelection, "gravity": ref.GravitySelection,
           "activity": ref.ActivityBasedSelection}[model]
    return cls(graph, **kwargs)


def main():
    out = {}
    for name, (kind, model, kwargs, seed, agents, calls) in CASES.items():
        g = make_graph(kind)
        if model in SEED_BEFORE_BUILD:
            random.seed(seed)
        sel = reference_selector(model, g, kwargs)
        if model not in SEED_BEFORE_BUILD:
            random.seed(seed)
        seq = ask(sel, agents, calls)
        idx = {n: i for i, n in enumerate(g.nodes())}
        out[name] = np.array([(idx[s], idx[t]) for s, t in seq], dtype=np.int32)
        print(name, len(seq), out[name][:3].tolist())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden_selectors", "selectors.npz"), **out)


if __name__ == "__main__":
    main()

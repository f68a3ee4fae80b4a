"""Contraction hierarchy (host side of the hub-label router, drift_hier_build: no device involved): distances through
the hierarchy equal Dijkstra's, shortcuts unpack to contiguous link walks of the stated length, the arrays survive the
disk cache, tie-prone graphs are refused."""
import heapq

import numpy as np
import pytest


def _climb(a, side, s):
    rp, ot, w = a[side + "_rowptr"], a[side + "_other"], a[side + "_w"]
    d = {s: 0.0}
    pq = [(0.0, s)]
    while pq:
        dd, x = heapq.heappop(pq)
        if dd > d[x]:
            continue
        for e in range(rp[x], rp[x + 1]):
            y, nd = int(ot[e]), dd + w[e]
            if nd < d.get(y, np.inf):
                d[y] = nd
                heapq.heappush(pq, (nd, y))
    return d


def _unpack(a, up, arc, out):
    stack = [(up, arc)]
    while stack:
        u, i = stack.pop()
        side = "up" if u else "down"
        x, y = int(a[side + "_a"][i]), int(a[side + "_b"][i])
        if y < 0:
            out.append(x)
        else:
            stack.append((True, y))
            stack.append((False, x))


@pytest.mark.parametrize("kind", ["road", "rgg"])
def test_hierarchy_distances_and_unpacking(kind, tmp_path):
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    from drift_b200 import hierarchy
    from drift_b200.synth import rgg_graph, road_graph

    lg = road_graph(3000, seed=3) if kind == "road" else rgg_graph(1500, seed=4)
    h = hierarchy.build(lg)
    a = h.arrays
    info = h.info()
    assert info["nodes"] == lg.V and info["levels"] == h.max_level + 1 and info["mean_label"] > 1
    # arcs lead to strictly higher levels on both sides
    for side in ("up", "down"):
        rp, ot = a[side + "_rowptr"], a[side + "_other"]
        tail = np.repeat(np.arange(lg.V), np.diff(rp))
        assert np.all(a["level"][ot] > a["level"][tail])
    # d(s, t) = min over common hubs of the two climbs == Dijkstra
    W = csr_matrix((lg.fwd_w, lg.fwd_col, lg.fwd_rowptr), shape=(lg.V, lg.V))
    rng = np.random.default_rng(0)
    src = rng.integers(0, lg.V, size=12)
    D = dijkstra(W, directed=True, indices=src)
    for k, s in enumerate(src.tolist()):
        f = _climb(a, "up", s)
        for t in rng.integers(0, lg.V, size=8).tolist():
            b = _climb(a, "down", t)
            best = min((f[x] + b[x] for x in f if x in b), default=np.inf)
            assert abs(best - D[k, t]) <= 1e-9 * max(1.0, D[k, t])
    # every arc unpacks to `hops` contiguous links from its tail to its head, of the arc's weight
    for side, up in (("up", True), ("down", False)):
        rp, ot = a[side + "_rowptr"], a[side + "_other"]
        tail = np.repeat(np.arange(lg.V), np.diff(rp))
        for i in rng.integers(0, len(ot), size=300).tolist():
            links = []
            _unpack(a, up, i, links)
            u, v = (int(tail[i]), int(ot[i])) if up else (int(ot[i]), int(tail[i]))
            assert len(links) == a[side + "_hops"][i]
            assert lg.link_u[links[0]] == u and lg.link_v[links[-1]] == v
            assert np.array_equal(lg.link_v[links[:-1]], lg.link_u[links[1:]])
            assert abs(float(np.sum(lg.link_t0[links])) - a[side + "_w"][i]) <= 1e-9 * a[side + "_w"][i]
    # disk cache round trip
    h2, cached = hierarchy.build_cached(lg, cache_dir=str(tmp_path))
    h3, cached3 = hierarchy.build_cached(lg, cache_dir=str(tmp_path))
    assert not cached and cached3 and h3.max_level == h.max_level
    for k in a:
        assert np.array_equal(a[k], h2.arrays[k]) and np.array_equal(a[k], h3.arrays[k]), k


def test_hierarchy_refuses_tie_prone_graphs():
    from drift_b200 import hierarchy
    from oracle import golden
    from drift_b200.graph import lower_graph

    lg = lower_graph(golden.GoldenCase("g100_random_s0").nx_graph())  # every link: travel_time 120.0
    with pytest.raises(ValueError):
        hierarchy.build(lg)

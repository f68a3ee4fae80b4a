"""Synthetic road networks for the large BASELINE.json configs.

The real city networks of the reference (example-data/json/*.json) are not shipped, so the
Brussels-scale / regional / grid / random-geometric networks are synthesised as link arrays and
lowered with ``drift_b200.graph.from_links`` (no NetworkX object is built at scale).  Lengths are
continuous on purpose: with continuous travel times shortest paths are unique, so link sequences
do not depend on NetworkX's tie-breaking.
"""
from __future__ import annotations

import numpy as np

from .graph import LoweredGraph, from_links


def grid_graph(nx_cols, ny_rows, seed=0, len_lo=50.0, len_hi=300.0, speed_kph=50.0, lanes=1.0,
               morton=False) -> LoweredGraph:
    """4-neighbour grid, both directions (config 4: 2000 x 2000 -> 4 M nodes / 15.992 M links).
    ``morton``: number the nodes along a Z-order curve instead of row by row (same graph, better
    locality of the per-node arrays)."""
    rng = np.random.default_rng(seed)
    W, H = int(nx_cols), int(ny_rows)
    V = W * H
    node = np.arange(V, dtype=np.int64)
    col, row = node % W, node // W
    cand_u, cand_v = [], []
    # per node, neighbours in a fixed order: left, right, up, down
    for dc, dr in ((-1, 0), (1, 0), (0, -1), (0, 1)):
        c2, r2 = col + dc, row + dr
        ok = (c2 >= 0) & (c2 < W) & (r2 >= 0) & (r2 < H)
        cand_u.append(np.where(ok, node, -1))
        cand_v.append(np.where(ok, r2 * W + c2, -1))
    u = np.stack(cand_u, axis=1).reshape(-1)
    v = np.stack(cand_v, axis=1).reshape(-1)
    keep = u >= 0
    u, v = u[keep], v[keep]
    L = len(u)
    length = rng.uniform(len_lo, len_hi, size=L)
    pos = np.stack([col, row], axis=1).astype(np.float64)
    if morton:
        rank = np.empty(V, dtype=np.int64)
        rank[np.argsort(_morton(col, row), kind="stable")] = np.arange(V)
        u, v = rank[u], rank[v]
        order = np.lexsort((v, u))
        u, v, length = u[order], v[order], length[order]
        p2 = np.empty_like(pos)
        p2[rank] = pos
        pos = p2
    return from_links(V, u, v, length, np.full(L, float(speed_kph)), np.full(L, float(lanes)), pos=pos)


def _morton(x, y):
    """Z-order code of 16-bit lattice coordinates: neighbouring nodes get nearby numbers, so the
    per-node arrays the router and the tick kernels gather from are touched a sector at a time."""
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
        return v
    return spread(x) | (spread(y) << np.uint64(1))


def road_graph(num_nodes, seed=0, mean_degree=2.6, morton=True) -> LoweredGraph:
    """Planar random road network (configs 2/3): nodes on a jittered lattice, lattice edges kept at
    random to reach the requested mean out-degree, both directions, largest connected component
    kept.  Lengths LogNormal(ln 90 m, 0.6) clipped to [10, 2000] (continuous => unique shortest
    paths); speed classes {30,50,70,90} km/h w.p. {.5,.3,.15,.05}; lanes {1,2,3}."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    rng = np.random.default_rng(seed)
    W = int(np.ceil(np.sqrt(num_nodes * 1.08)))
    H = int(np.ceil(num_nodes * 1.08 / W))
    V0 = W * H
    node = np.arange(V0, dtype=np.int64)
    col, row = node % W, node // W
    p_keep = min(1.0, mean_degree / 4.0 * 1.04)
    und = []
    for dc, dr in ((1, 0), (0, 1)):
        ok = (col + dc < W) & (row + dr < H) & (rng.random(V0) < p_keep)
        und.append(np.stack([node[ok], ((row + dr) * W + col + dc)[ok]], axis=1))
    und = np.concatenate(und)
    adj = coo_matrix((np.ones(len(und), dtype=np.int8), (und[:, 0], und[:, 1])), shape=(V0, V0))
    _, lab = connected_components(adj, directed=False)
    big = np.argmax(np.bincount(lab))
    keep = lab == big
    if morton:  # number the kept nodes along a Z-order curve instead of row by row
        kept = np.nonzero(keep)[0]
        rank = np.empty(len(kept), dtype=np.int64)
        rank[np.argsort(_morton(col[kept], row[kept]), kind="stable")] = np.arange(len(kept))
        new_id = np.full(V0, -1, dtype=np.int64)
        new_id[kept] = rank
    else:
        new_id = np.cumsum(keep) - 1
    und = und[keep[und[:, 0]] & keep[und[:, 1]]]
    und = np.stack([new_id[und[:, 0]], new_id[und[:, 1]]], axis=1)
    V = int(keep.sum())
    u = np.concatenate([und[:, 0], und[:, 1]])
    v = np.concatenate([und[:, 1], und[:, 0]])
    order = np.lexsort((v, u))
    u, v = u[order], v[order]
    L = len(u)
    length = np.clip(rng.lognormal(np.log(90.0), 0.6, size=L), 10.0, 2000.0)
    speed = rng.choice([30.0, 50.0, 70.0, 90.0], size=L, p=[0.5, 0.3, 0.15, 0.05])
    lanes = rng.choice([1.0, 2.0, 3.0], size=L, p=[0.6, 0.3, 0.1])
    pos = np.empty((V, 2))
    kept = np.nonzero(keep)[0]
    pos[new_id[kept], 0] = col[kept] + rng.uniform(-0.3, 0.3, V)
    pos[new_id[kept], 1] = row[kept] + rng.uniform(-0.3, 0.3, V)
    return from_links(V, u, v, length, speed, lanes, pos=pos)


def rgg_graph(num_nodes, seed=0, mean_degree=6.0, mean_length_m=120.0, speed_kph=50.0, morton=True) -> LoweredGraph:
    """Random geometric graph (config 5): points uniform in the unit square, an edge in both
    directions between points closer than the radius that gives ``mean_degree`` neighbours, largest
    connected component kept, length = Euclidean distance scaled to ``mean_length_m`` on average
    (continuous => unique shortest paths), one lane at ``speed_kph``."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    n0 = int(num_nodes * 1.03)
    pts = rng.random((n0, 2))
    r = np.sqrt(mean_degree / (np.pi * n0))
    pairs = cKDTree(pts).query_pairs(r, output_type="ndarray")
    adj = coo_matrix((np.ones(len(pairs), dtype=np.int8), (pairs[:, 0], pairs[:, 1])), shape=(n0, n0))
    _, lab = connected_components(adj, directed=False)
    keep = lab == np.argmax(np.bincount(lab))
    kept = np.nonzero(keep)[0]
    V = len(kept)
    new_id = np.full(n0, -1, dtype=np.int64)
    if morton:
        q = np.minimum((pts[kept] * 65535.0).astype(np.int64), 65535)
        new_id[kept[np.argsort(_morton(q[:, 0], q[:, 1]), kind="stable")]] = np.arange(V)
    else:
        new_id[kept] = np.arange(V)
    pairs = pairs[keep[pairs[:, 0]] & keep[pairs[:, 1]]]
    a, b = new_id[pairs[:, 0]], new_id[pairs[:, 1]]
    d = np.hypot(*(pts[pairs[:, 0]] - pts[pairs[:, 1]]).T)
    u = np.concatenate([a, b])
    v = np.concatenate([b, a])
    length = np.concatenate([d, d])
    order = np.lexsort((v, u))
    u, v, length = u[order], v[order], length[order]
    length = np.maximum(length * (mean_length_m / length.mean()), 1.0)
    L = len(u)
    pos = np.empty((V, 2))
    pos[new_id[kept]] = pts[kept]
    return from_links(V, u, v, length, np.full(L, float(speed_kph)), np.ones(L), pos=pos)

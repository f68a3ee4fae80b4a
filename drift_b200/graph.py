"""Graph lowering: loader-processed ``nx.MultiDiGraph`` -> device-ready CSR + link arrays.

The reference keeps the road network as a NetworkX dict-of-dict-of-dict and walks it from
Python on every link entry and every route (lib/agent.py:94,118-151; lib/managers/
traffic_manager.py:33-61).  Here it is lowered once:

* nodes are numbered in ``list(G.nodes)`` order, links in ``G.edges(keys=True)`` order (one link
  per (u, v, key) edge: the unit of occupancy);
* the forward CSR follows ``G._succ[u]`` insertion order and the backward CSR ``G._pred[v]``
  insertion order -- NetworkX's bidirectional Dijkstra scans neighbours in exactly that order
  and its tie-breaking depends on it;
* parallel (u, v, *) edges collapse to one CSR entry with the minimum ``travel_time``
  (nx weighted.py:76-77) that remembers the link agents actually occupy on that hop: the first
  key with minimal ``travel_time`` (lib/agent.py:122-130), and the minimum ``length`` used by
  the trip-distance column (lib/data_logger.py:113-114).

Attribute defaults are the reference's (config.py:184-185): 30 km/h, 1000 m, 1 lane.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional

import numpy as np

DEFAULT_SPEED_KPH = 30.0
DEFAULT_EDGE_LENGTH = 1000.0
VEHICLE_SPACE_M = 7.5


def capacities(lanes, speed_kph, length_m):
    """Vectorised lib/managers/traffic_manager.py:50-61 (all inputs float64 arrays)."""
    lanes = np.asarray(lanes, dtype=np.float64)
    speed_kph = np.asarray(speed_kph, dtype=np.float64)
    length_m = np.asarray(length_m, dtype=np.float64)
    per_lane = np.where(speed_kph >= 50, 2000.0, 1500.0)
    flow = np.trunc(lanes * per_lane).astype(np.int64)  # int() truncates toward zero
    storage = np.maximum(1, np.trunc(length_m / VEHICLE_SPACE_M).astype(np.int64)) * 4
    return np.minimum(flow, storage).astype(np.int32)


@dataclass
class LoweredGraph:
    nodes: List[Any]
    link_u: np.ndarray
    link_v: np.ndarray
    link_key: Optional[List[Any]]
    link_len: np.ndarray
    link_v0: np.ndarray
    link_t0: np.ndarray
    link_cap: np.ndarray
    fwd_rowptr: np.ndarray
    fwd_col: np.ndarray
    fwd_w: np.ndarray
    fwd_link: np.ndarray
    fwd_minlen: np.ndarray
    bwd_rowptr: np.ndarray
    bwd_col: np.ndarray
    bwd_w: np.ndarray
    bwd_link: np.ndarray
    pos: Optional[np.ndarray] = None
    _node_index: Optional[dict] = field(default=None, repr=False)

    @property
    def V(self):
        return len(self.fwd_rowptr) - 1

    @property
    def L(self):
        return len(self.link_u)

    @property
    def node_index(self):
        if self._node_index is None:
            self._node_index = {n: i for i, n in enumerate(self.nodes)}
        return self._node_index

    def link_tuple(self, l):
        """(u, v, key) label of link ``l`` (TrafficManager.edge_agents key)."""
        k = self.link_key[l] if self.link_key is not None else 0
        return (self.nodes[self.link_u[l]], self.nodes[self.link_v[l]], k)

    def hop_entry(self, u, v):
        """CSR entry index of the (u, v) pair, or -1."""
        a, b = self.fwd_rowptr[u], self.fwd_rowptr[u + 1]
        hit = np.nonzero(self.fwd_col[a:b] == v)[0]
        return int(a + hit[0]) if len(hit) else -1

    def links_to_nodes(self, src, links):
        """Node path [src, ...] of a link sequence."""
        return [int(src)] + self.link_v[np.asarray(links, dtype=np.int64)].tolist() if len(links) else [int(src)]

    def nodes_to_links(self, path):
        out = []
        for a, b in zip(path[:-1], path[1:]):
            e = self.hop_entry(a, b)
            if e < 0:
                raise ValueError("no edge %r -> %r" % (a, b))
            out.append(int(self.fwd_link[e]))
        return out

    def path_length_m(self, path_nodes):
        """data_logger.py:103-120: left-to-right sum of the min 'length' per hop."""
        total = 0.0
        for a, b in zip(path_nodes[:-1], path_nodes[1:]):
            e = self.hop_entry(a, b)
            if e >= 0:
                total += float(self.fwd_minlen[e])
        return total


def _f(x, default):
    return float(x) if x is not None else default


def lower_graph(G) -> LoweredGraph:
    """Lower a loader-processed ``nx.MultiDiGraph`` (lib/graph_loader.py:47-134 output)."""
    nodes = list(G.nodes)
    index = {n: i for i, n in enumerate(nodes)}
    V = len(nodes)
    L = G.number_of_edges()
    link_u = np.empty(L, dtype=np.int32)
    link_v = np.empty(L, dtype=np.int32)
    link_key = [None] * L
    link_len = np.empty(L)
    link_v0 = np.empty(L)
    link_t0 = np.empty(L)
    lanes = np.empty(L)
    cap_speed = np.empty(L)
    cap_len = np.empty(L)
    fr, fc, fw, fl, fm = [0], [], [], [], []
    pair_w = {}
    pair_link = {}
    l = 0
    for u in nodes:  # G.edges(keys=True) order == nodes x _succ order x key order
        ui = index[u]
        for v, parallel in G._succ[u].items():
            vi = index[v]
            best_l, best_tt, wmin, lmin = -1, None, None, None
            for key, attr in parallel.items():
                link_u[l], link_v[l], link_key[l] = ui, vi, key
                try:  # lib/agent.py:145-151: both fall back together
                    ln = float(attr.get("length", DEFAULT_EDGE_LENGTH))
                    sp = float(attr.get("speed_kph", DEFAULT_SPEED_KPH))
                except (TypeError, ValueError):
                    ln, sp = DEFAULT_EDGE_LENGTH, DEFAULT_SPEED_KPH
                link_len[l], link_v0[l] = ln, sp
                try:  # traffic_manager.py:40-48: all three fall back together
                    lanes[l] = float(attr.get("lanes", 1))
                    cap_speed[l] = float(attr.get("speed_kph", DEFAULT_SPEED_KPH))
                    cap_len[l] = float(attr.get("length", DEFAULT_EDGE_LENGTH))
                except (TypeError, ValueError):
                    lanes[l], cap_speed[l], cap_len[l] = 1.0, DEFAULT_SPEED_KPH, DEFAULT_EDGE_LENGTH
                tt_route = attr.get("travel_time", 1)               # nx weight: missing -> 1
                tt_pick = attr.get("travel_time", float("inf"))     # agent.py:123: missing -> inf
                link_t0[l] = float(tt_route)
                if wmin is None or tt_route < wmin:
                    wmin = tt_route
                if best_tt is None or tt_pick < best_tt:
                    best_l, best_tt = l, tt_pick
                raw_len = attr.get("length", float("inf"))          # data_logger.py:113
                if lmin is None or raw_len < lmin[0]:
                    lmin = (raw_len, attr.get("length", DEFAULT_EDGE_LENGTH))
                l += 1
            fc.append(vi)
            fw.append(float(wmin))
            fl.append(best_l)
            fm.append(float(lmin[1]))
            pair_w[(ui, vi)] = float(wmin)
            pair_link[(ui, vi)] = best_l
        fr.append(len(fc))
    assert l == L
    br, bc, bw, bl = [0], [], [], []
    for v in nodes:
        vi = index[v]
        for u in G._pred[v]:
            ui = index[u]
            bc.append(ui)
            bw.append(pair_w[(ui, vi)])
            bl.append(pair_link[(ui, vi)])
        br.append(len(bc))
    pos = None
    try:
        pos = np.array([G.nodes[n]["pos"] for n in nodes], dtype=np.float64)
    except (KeyError, TypeError, ValueError):
        pos = None
    return LoweredGraph(
        nodes=nodes, link_u=link_u, link_v=link_v, link_key=link_key, link_len=link_len, link_v0=link_v0,
        link_t0=link_t0, link_cap=capacities(lanes, cap_speed, cap_len),
        fwd_rowptr=np.array(fr, dtype=np.int32), fwd_col=np.array(fc, dtype=np.int32),
        fwd_w=np.array(fw, dtype=np.float64), fwd_link=np.array(fl, dtype=np.int32),
        fwd_minlen=np.array(fm, dtype=np.float64),
        bwd_rowptr=np.array(br, dtype=np.int32), bwd_col=np.array(bc, dtype=np.int32),
        bwd_w=np.array(bw, dtype=np.float64), bwd_link=np.array(bl, dtype=np.int32),
        pos=pos, _node_index=index)


def from_links(num_nodes, link_u, link_v, length_m, speed_kph, lanes=None, nodes=None, pos=None,
               pred_order=None) -> LoweredGraph:
    """Build a lowered graph straight from link arrays (synthetic networks; no NetworkX object).

    Links must be grouped by tail node in ascending order (the order ``G.edges`` would give for
    nodes added 0..V-1); (u, v) pairs must be unique (no parallel edges).  travel_time follows
    lib/graph_loader.py:759-766: ``(length/1000) / speed_kph * 3600``.  The backward CSR lists the
    predecessors of v in link order (what ``G._pred`` holds when edges are added in link order)
    unless ``pred_order`` (rowptr, col) is given.
    """
    link_u = np.ascontiguousarray(link_u, dtype=np.int32)
    link_v = np.ascontiguousarray(link_v, dtype=np.int32)
    L = len(link_u)
    V = int(num_nodes)
    if L and np.any(np.diff(link_u) < 0):
        raise ValueError("links must be grouped by tail node in ascending order")
    length_m = np.ascontiguousarray(length_m, dtype=np.float64)
    speed_kph = np.ascontiguousarray(speed_kph, dtype=np.float64)
    lanes = np.ones(L) if lanes is None else np.ascontiguousarray(lanes, dtype=np.float64)
    t0 = ((length_m / 1000.0) / speed_kph) * 3600
    fr = np.zeros(V + 1, dtype=np.int64)
    np.add.at(fr, link_u.astype(np.int64) + 1, 1)
    fr = np.cumsum(fr).astype(np.int32)
    ids = np.arange(L, dtype=np.int32)
    if pred_order is None:
        order = np.argsort(link_v, kind="stable")
        br = np.zeros(V + 1, dtype=np.int64)
        np.add.at(br, link_v.astype(np.int64) + 1, 1)
        br = np.cumsum(br).astype(np.int32)
        bc = link_u[order]
        bl = ids[order]
        bw = t0[order]
    else:
        raise NotImplementedError("explicit predecessor order")
    return LoweredGraph(
        nodes=list(range(V)) if nodes is None else nodes, link_u=link_u, link_v=link_v, link_key=None,
        link_len=length_m, link_v0=speed_kph, link_t0=t0, link_cap=capacities(lanes, speed_kph, length_m),
        fwd_rowptr=fr, fwd_col=link_v.copy(), fwd_w=t0.copy(), fwd_link=ids, fwd_minlen=length_m.copy(),
        bwd_rowptr=br, bwd_col=np.ascontiguousarray(bc), bwd_w=np.ascontiguousarray(bw),
        bwd_link=np.ascontiguousarray(bl), pos=pos)


def to_networkx(lg: LoweredGraph):
    """Inverse of ``from_links`` for small graphs: the ``nx.MultiDiGraph`` the reference would hold
    (used to hand a synthetic network to the reference / to live NetworkX routing in tests)."""
    import networkx as nx

    G = nx.MultiDiGraph()
    for i, n in enumerate(lg.nodes):
        if lg.pos is not None:
            G.add_node(n, pos=(float(lg.pos[i, 0]), float(lg.pos[i, 1])))
        else:
            G.add_node(n)
    for l in range(lg.L):
        G.add_edge(lg.nodes[lg.link_u[l]], lg.nodes[lg.link_v[l]],
                   key=(lg.link_key[l] if lg.link_key is not None else 0),
                   length=float(lg.link_len[l]), speed_kph=float(lg.link_v0[l]),
                   travel_time=float(lg.link_t0[l]))
    return G

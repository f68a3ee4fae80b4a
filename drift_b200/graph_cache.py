"""Persisted lowering of ``graph_loader`` output.

Walking a 1 M-link ``nx.MultiDiGraph`` in Python to build the CSR takes tens of seconds; the lowered
arrays are therefore cached as ``.npz`` keyed by a digest of what the lowering depends on: node order,
``_succ`` / ``_pred`` insertion order and the per-edge ``length`` / ``speed_kph`` / ``travel_time`` /
``lanes`` attributes.
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from .graph import LoweredGraph, lower_graph

_ARRAYS = ("link_u", "link_v", "link_len", "link_v0", "link_t0", "link_cap", "fwd_rowptr", "fwd_col", "fwd_w",
           "fwd_link", "fwd_minlen", "bwd_rowptr", "bwd_col", "bwd_w", "bwd_link")


def graph_digest(G) -> str:
    h = hashlib.sha256()
    h.update(repr(list(G.nodes)).encode())
    for u in G.nodes:
        for v, par in G._succ[u].items():
            for k, a in par.items():
                h.update(repr((u, v, k, a.get("length"), a.get("speed_kph"), a.get("travel_time"), a.get("lanes"))).encode())
    for v in G.nodes:
        h.update(repr(list(G._pred[v])).encode())
    return h.hexdigest()[:32]


def save(lg: LoweredGraph, path):
    data = {k: getattr(lg, k) for k in _ARRAYS}
    data["nodes_json"] = json.dumps(lg.nodes)
    data["link_key_json"] = json.dumps(lg.link_key)
    if lg.pos is not None:
        data["pos"] = lg.pos
    np.savez_compressed(path, **data)


def load(path) -> LoweredGraph:
    z = np.load(path, allow_pickle=False)
    return LoweredGraph(nodes=json.loads(str(z["nodes_json"])), link_key=json.loads(str(z["link_key_json"])),
                        pos=z["pos"] if "pos" in z.files else None, **{k: z[k] for k in _ARRAYS})


def lower_cached(G, cache_dir=".drift_b200_cache") -> LoweredGraph:
    """``lower_graph(G)`` through the on-disk cache.  Node labels must be JSON-serialisable."""
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, graph_digest(G) + ".npz")
    if os.path.exists(path):
        return load(path)
    lg = lower_graph(G)
    try:
        save(lg, path)
    except TypeError:
        pass  # exotic labels: fall back to lowering every time
    return lg

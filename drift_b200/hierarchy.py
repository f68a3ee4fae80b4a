"""Contraction hierarchy of a lowered graph (host side) -- input of the hub-label router.

The reference computes every route with ``nx.shortest_path(G, s, t, weight='travel_time')``
(lib/agent.py:94) on travel times that never change (lib/graph_loader.py:759-766).  ``build`` preprocesses
those static travel times once (``drift_hier_build``: C++, no device needed); ``Engine.set_hierarchy`` uploads
the result and builds the hub labels on the device, after which ``ROUTER_BATCHED`` answers queries by label
lookups.  ``save`` / ``load`` keep a hierarchy next to the graph cache (``drift_b200.graph_cache``) so that
the ranks of one box, or repeated runs, do not redo the contraction.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os

import numpy as np

from . import _capi

_ARRAYS = ("level", "up_rowptr", "up_other", "up_w", "up_a", "up_b", "up_hops",
           "down_rowptr", "down_other", "down_w", "down_a", "down_b", "down_hops")


class Hierarchy:
    """Plain numpy arrays of ``drift_hier_desc`` (include/drift_b200.h) + build statistics."""

    def __init__(self, arrays, max_level, max_label, mean_label, shortcuts=0, build_seconds=0.0):
        self.arrays = {k: np.ascontiguousarray(arrays[k]) for k in _ARRAYS}
        self.V = len(self.arrays["level"])
        self.max_level, self.max_label, self.mean_label = int(max_level), int(max_label), float(mean_label)
        self.shortcuts, self.build_seconds = int(shortcuts), float(build_seconds)

    def desc(self):
        a = self.arrays
        d = _capi.HierDesc(V=self.V, max_level=self.max_level, max_label=self.max_label, pad0=0,
                           mean_label=self.mean_label, n_up=len(a["up_other"]), n_down=len(a["down_other"]))
        for k in _ARRAYS:
            ct = C.c_double if a[k].dtype == np.float64 else C.c_int32
            setattr(d, k, a[k].ctypes.data_as(C.POINTER(ct)))
        return d

    def info(self):
        a = self.arrays
        return dict(nodes=self.V, up_arcs=len(a["up_other"]), down_arcs=len(a["down_other"]), shortcuts=self.shortcuts,
                    levels=self.max_level + 1, max_label_sampled=self.max_label, mean_label=self.mean_label,
                    build_seconds=self.build_seconds)

    def save(self, path):
        np.savez(path, meta=np.array([self.max_level, self.max_label, self.shortcuts], dtype=np.int64),
                 fmeta=np.array([self.mean_label, self.build_seconds]), **self.arrays)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        m, f = z["meta"], z["fmeta"]
        return cls({k: z[k] for k in _ARRAYS}, m[0], m[1], f[0], shortcuts=m[2], build_seconds=f[1])


def graph_digest(lg):
    h = hashlib.sha256()
    for a in (lg.fwd_rowptr, lg.fwd_col, lg.fwd_w, lg.fwd_link):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:24]


def build(lg, witness_settled=100, time_limit_s=0.0):
    """Contract the graph.  Raises ``_capi.DriftError`` (DRIFT_ESTATE) when ``time_limit_s`` is hit."""
    lib = _capi.load()
    if len(lg.fwd_w) >= 64 and len(np.unique(lg.fwd_w)) < 0.5 * len(lg.fwd_w):
        # e.g. the bundled graphs: one travel_time value for every link -> most queries tie and the
        # reference's answer depends on NetworkX's search order, which only the exact router reproduces
        raise ValueError("tie-prone graph (few distinct travel times): the hub-label router needs unique shortest "
                         "paths; use ROUTER_BATCHED without a hierarchy (tie detector + exact fallback) or ROUTER_NX_EXACT")
    rp = np.ascontiguousarray(lg.fwd_rowptr, dtype=np.int32)
    col = np.ascontiguousarray(lg.fwd_col, dtype=np.int32)
    w = np.ascontiguousarray(lg.fwd_w, dtype=np.float64)
    lk = np.ascontiguousarray(lg.fwd_link, dtype=np.int32)
    h = _capi.HIER()
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    _capi.check(lib.drift_hier_build(lg.V, len(col), p(rp, C.c_int32), p(col, C.c_int32), p(w, C.c_double),
                                     p(lk, C.c_int32), int(witness_settled), float(time_limit_s), C.byref(h)))
    try:
        n_up, n_down, sc = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        ml, mx = C.c_int32(0), C.c_int32(0)
        mean, secs = C.c_double(0), C.c_double(0)
        _capi.check(lib.drift_hier_info(h, C.byref(n_up), C.byref(n_down), C.byref(sc), C.byref(ml), C.byref(mx),
                                        C.byref(mean), C.byref(secs)))
        V = lg.V
        arrays = dict(level=np.zeros(V, np.int32))
        for side, n in (("up", n_up.value), ("down", n_down.value)):
            arrays[side + "_rowptr"] = np.zeros(V + 1, np.int32)
            arrays[side + "_other"] = np.zeros(n, np.int32)
            arrays[side + "_w"] = np.zeros(n, np.float64)
            arrays[side + "_a"] = np.zeros(n, np.int32)
            arrays[side + "_b"] = np.zeros(n, np.int32)
            arrays[side + "_hops"] = np.zeros(n, np.int32)
        out = Hierarchy(arrays, ml.value, mx.value, mean.value, shortcuts=sc.value, build_seconds=secs.value)
        d = out.desc()
        _capi.check(lib.drift_hier_export(h, C.byref(d)))
        return out
    finally:
        lib.drift_hier_free(h)


def build_cached(lg, cache_dir=None, **kw):
    """``build`` behind a file cache keyed by the graph's CSR (safe against concurrent writers: temp file +
    rename).  Returns (hierarchy, from_cache)."""
    cache_dir = cache_dir or os.path.join(os.environ.get("TMPDIR", "/tmp"), "drift_b200_cache")
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, "hier_%s_w%d.npz" % (graph_digest(lg), int(kw.get("witness_settled", 100))))
    if os.path.exists(path):
        try:
            return Hierarchy.load(path), True
        except Exception:
            pass
    h = build(lg, **kw)
    tmp = "%s.%d.tmp.npz" % (path, os.getpid())
    h.save(tmp)
    os.replace(tmp, path)
    return h, False

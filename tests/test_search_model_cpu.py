"""A sequential model of the goal-directed bounded search of ``k_build_trees`` (drift_b200/csrc/route_kernels.cuh):
reverse search from the destination, near / far queues with the bucket rule of the kernel, ALT potentials taken from
the landmark tables and -- as the kernel caches them in its 16-byte search record -- stored as floats ROUNDED DOWN.
The claim the kernel relies on: admissible potentials are enough for its stopping rule ("near queue empty and the
source's label <= T"), so the label of the source is the exact shortest distance, the same fp64 sum a plain Dijkstra
from the destination accumulates.  The model is checked against scipy's Dijkstra with ``==``.

Test infrastructure only (no product code runs through it); the GPU parity tests of the kernel itself are
tests/test_gpu_route.py and tests/test_gpu_scale.py."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import dijkstra

from drift_b200.engine import landmark_tables
from drift_b200.synth import grid_graph, road_graph


def _pot64(frm, to, src, v):
    """The kernel's ``pot``: maximum over all landmark bounds on d(src, v), 0 if none is positive and finite."""
    h = 0.0
    for k in range(frm.shape[0]):
        b0 = frm[k][v] - frm[k][src]
        b1 = to[k][src] - to[k][v]
        if b0 > h and b0 < 1e300:
            h = b0
        if b1 > h and b1 < 1e300:
            h = b1
    return h


def _round_down_f32(x):
    f = np.float32(x)
    if float(f) > x:
        f = np.nextafter(f, np.float32(-np.inf))
    return float(f)


def bounded_search(lg, frm, to, src, dest, delta, rounded=True):
    """Label of ``src`` (= d(src -> dest)) and the number of expanded nodes; mirrors the kernel's control flow: a wave
    reads the labels of its near nodes at its start, improvements go to the next wave when label + potential <= T and
    to the far queue (one entry per node) otherwise; an empty near queue ends the search once the source's label is
    <= T, else T moves to the smallest far key + jdelta and the far queue is split."""
    INF = float("inf")
    dist = {dest: 0.0}
    pot = {}

    def p(v):
        if v not in pot:
            h = _pot64(frm, to, src, v)
            pot[v] = _round_down_f32(h) if rounded else h
        return pot[v]

    jdelta = delta * 0.25
    T_old, T = -1.0, _pot64(frm, to, src, dest) + jdelta
    near, far, infar = [dest], [], set()
    expanded = 0
    while True:
        if not near:
            if not far:
                break
            if dist.get(src, INF) <= T:
                break
            minfar = min(dist[u] + p(u) for u in far)
            T_old, T = T, max(minfar, T) + jdelta
            keep = []
            for u in far:
                d = dist[u] + p(u)
                if d <= T_old:
                    infar.discard(u)
                elif d <= T:
                    infar.discard(u)
                    near.append(u)
                else:
                    keep.append(u)
            far = keep
            continue
        snapshot = [(v, dist[v]) for v in near]
        nxt, stamped = [], set()
        for v, dv in snapshot:
            expanded += 1
            for e in range(lg.bwd_rowptr[v], lg.bwd_rowptr[v + 1]):
                u = int(lg.bwd_col[e])
                cand = dv + float(lg.bwd_w[e])
                if cand < dist.get(u, INF):
                    dist[u] = cand
                    if cand + p(u) <= T:
                        if u not in stamped:
                            stamped.add(u)
                            nxt.append(u)
                    elif u not in infar:
                        infar.add(u)
                        far.append(u)
        near = nxt
    return dist.get(src, INF), expanded


@pytest.mark.parametrize("kind", ["road", "grid"])
def test_rounded_down_potentials_keep_the_bounded_search_exact(kind):
    lg = road_graph(2500, seed=3) if kind == "road" else grid_graph(45, 45, seed=3, morton=True)
    picks, frm, to = landmark_tables(lg, K=4, seed=0)
    WT = csr_matrix((lg.bwd_w, lg.bwd_col, lg.bwd_rowptr), shape=(lg.V, lg.V))
    delta = float(np.mean(lg.fwd_w)) * 2.0
    rng = np.random.default_rng(7)
    pairs = rng.integers(0, lg.V, size=(25, 2))
    total_goal = total_plain = 0
    for s, t in pairs:
        s, t = int(s), int(t)
        if s == t:
            continue
        # labels grow from the destination: dist[u] = dist[v] + w(u, v), the sums scipy forms on the transposed graph
        ref = dijkstra(WT, directed=True, indices=t)
        got, n_goal = bounded_search(lg, frm, to, s, t, delta, rounded=True)
        assert got == ref[s], (kind, s, t, got, ref[s])
        got64, _ = bounded_search(lg, frm, to, s, t, delta, rounded=False)
        assert got64 == ref[s]
        zero = np.zeros_like(frm)
        _, n_plain = bounded_search(lg, zero, zero, s, t, delta, rounded=True)  # no goal direction
        total_goal += n_goal
        total_plain += n_plain
    assert total_goal < total_plain  # the potentials do prune (what they are there for)

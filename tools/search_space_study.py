"""How many nodes must a point-to-point search settle?  CPU study (heapq) that sizes the next step of the
bounded search on the synthetic road graph: unidirectional Dijkstra / ALT (what k_build_trees does for a cold
pair today, minus the bucket slack) against bidirectional Dijkstra and bidirectional ALT with consistent
(averaged) potentials.  Prints settled nodes and relaxed edges per query.

    python tools/search_space_study.py [--nodes 50000] [--pairs 150] [--landmarks 8]
"""
import argparse
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def landmarks(lg, K, seed=0):
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    V = lg.V
    W = csr_matrix((lg.fwd_w, lg.fwd_col, lg.fwd_rowptr), shape=(V, V))
    WT = csr_matrix((lg.bwd_w, lg.bwd_col, lg.bwd_rowptr), shape=(V, V))
    rng = np.random.default_rng(seed)
    d0 = dijkstra(W, directed=True, indices=int(rng.integers(0, V)))
    picks, mind = [], None
    cur = int(np.argmax(np.where(np.isfinite(d0), d0, -1)))
    frm, to = [], []
    for _ in range(K):
        picks.append(cur)
        f = dijkstra(W, directed=True, indices=cur)
        t = dijkstra(WT, directed=True, indices=cur)
        frm.append(f)
        to.append(t)
        m = np.minimum(f, t)
        mind = m if mind is None else np.minimum(mind, m)
        cur = int(np.argmax(np.where(np.isfinite(mind), mind, -1)))
    return np.array(frm), np.array(to)


def pot_to_target(frm, to, t):
    """pi(v) <= d(v, t):  max_k max(to[k][v] - to[k][t], frm[k][t] - frm[k][v])"""
    return lambda v: max(0.0, float(np.max(np.maximum(to[:, v] - to[:, t], frm[:, t] - frm[:, v]))))


def pot_to_target_best2(frm, to, s, t):
    """What the GPU kernel used until now: only the two bounds that are largest for the pair (s, t) itself."""
    b = np.concatenate([to[:, s] - to[:, t], frm[:, t] - frm[:, s]])
    pick = np.argsort(-b)[:2]
    K = frm.shape[0]

    def pot(v):
        best = 0.0
        for i in pick:
            x = to[i, v] - to[i, t] if i < K else frm[i - K, t] - frm[i - K, v]
            best = max(best, float(x))
        return best
    return pot


def pot_from_source(frm, to, s):
    """pi(v) <= d(s, v)"""
    return lambda v: max(0.0, float(np.max(np.maximum(frm[:, v] - frm[:, s], to[:, s] - to[:, v]))))


def uni(lg, s, t, pot=None):
    dist = {s: 0.0}
    done = set()
    pq = [(pot(s) if pot else 0.0, s)]
    relaxed = 0
    while pq:
        _, u = heapq.heappop(pq)
        if u in done:
            continue
        done.add(u)
        if u == t:
            break
        for e in range(lg.fwd_rowptr[u], lg.fwd_rowptr[u + 1]):
            v = int(lg.fwd_col[e])
            relaxed += 1
            nd = dist[u] + lg.fwd_w[e]
            if nd < dist.get(v, np.inf):
                dist[v] = nd
                heapq.heappush(pq, (nd + (pot(v) if pot else 0.0), v))
    return dist.get(t, np.inf), len(done), relaxed


def bidir(lg, s, t, pf=None, pb=None):
    """Bidirectional search; with potentials pf (to target) / pb (from source) the consistent pair
    (pf - pb) / 2, (pb - pf) / 2 is used (Ikeda et al.): both searches see the same reduced costs and the
    plain stopping rule (sum of the two top keys >= best path found) stays valid."""
    if pf:
        hf = lambda v: 0.5 * (pf(v) - pb(v))
        hb = lambda v: 0.5 * (pb(v) - pf(v))
    else:
        hf = hb = lambda v: 0.0
    dist = [{s: 0.0}, {t: 0.0}]
    done = [set(), set()]
    pq = [[(hf(s), s)], [(hb(t), t)]]
    best = np.inf
    relaxed = 0
    rowptr = [lg.fwd_rowptr, lg.bwd_rowptr]
    col = [lg.fwd_col, lg.bwd_col]
    w = [lg.fwd_w, lg.bwd_w]
    h = [hf, hb]
    while pq[0] and pq[1]:
        if pq[0][0][0] + pq[1][0][0] >= best:
            break
        d = 0 if pq[0][0][0] <= pq[1][0][0] else 1
        _, u = heapq.heappop(pq[d])
        if u in done[d]:
            continue
        done[d].add(u)
        for e in range(rowptr[d][u], rowptr[d][u + 1]):
            v = int(col[d][e])
            relaxed += 1
            nd = dist[d][u] + w[d][e]
            if nd < dist[d].get(v, np.inf):
                dist[d][v] = nd
                heapq.heappush(pq[d], (nd + h[d](v), v))
                if v in dist[1 - d]:
                    best = min(best, nd + dist[1 - d][v])
    return best, len(done[0]) + len(done[1]), relaxed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=50000)
    ap.add_argument("--pairs", type=int, default=150)
    ap.add_argument("--landmarks", type=int, default=8)
    ap.add_argument("--graph", default="road", choices=["road", "grid"])
    a = ap.parse_args()
    from drift_b200.synth import grid_graph, road_graph

    lg = road_graph(a.nodes, seed=0) if a.graph == "road" else grid_graph(int(a.nodes ** 0.5), int(a.nodes ** 0.5), morton=True)
    frm, to = landmarks(lg, a.landmarks)
    rng = np.random.default_rng(1)
    rows = {k: [] for k in ("dijkstra", "alt_best2", "alt", "bidir", "bidir_alt")}
    for _ in range(a.pairs):
        s, t = int(rng.integers(0, lg.V)), int(rng.integers(0, lg.V))
        if s == t:
            continue
        d0, n0, r0 = uni(lg, s, t)
        if not np.isfinite(d0):
            continue
        d1, n1, r1 = uni(lg, s, t, pot_to_target(frm, to, t))
        d4, n4, r4 = uni(lg, s, t, pot_to_target_best2(frm, to, s, t))
        assert abs(d4 - d0) <= 1e-9 * d0
        rows["alt_best2"].append((n4, r4))
        d2, n2, r2 = bidir(lg, s, t)
        d3, n3, r3 = bidir(lg, s, t, pot_to_target(frm, to, t), pot_from_source(frm, to, s))
        assert abs(d1 - d0) <= 1e-9 * d0 and abs(d2 - d0) <= 1e-9 * d0 and abs(d3 - d0) <= 1e-9 * d0, (d0, d1, d2, d3)
        rows["dijkstra"].append((n0, r0))
        rows["alt"].append((n1, r1))
        rows["bidir"].append((n2, r2))
        rows["bidir_alt"].append((n3, r3))
    print("graph %s: %d nodes, %d links, %d landmarks, %d pairs" % (a.graph, lg.V, lg.L, a.landmarks, len(rows["dijkstra"])))
    base = np.mean([x[0] for x in rows["dijkstra"]])
    for k, v in rows.items():
        n = np.mean([x[0] for x in v])
        print("%-10s settled %9.0f  relaxed %9.0f  (%.2fx fewer settled than Dijkstra)" % (k, n, np.mean([x[1] for x in v]), base / n))


if __name__ == "__main__":
    main()

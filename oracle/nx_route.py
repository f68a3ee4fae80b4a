"""TEST INFRASTRUCTURE ONLY -- restatement of NetworkX 3.6.1 ``bidirectional_dijkstra``.

DRIFT routes with ``nx.shortest_path(G, s, t, weight='travel_time')``
(reference lib/agent.py:94), which NetworkX 3.6.1 dispatches to
``bidirectional_dijkstra`` (networkx/algorithms/shortest_paths/generic.py:172-175,
weighted.py:2310-2459; NetworkX is a third-party dependency, ``requirements.txt:2``
``networkx>=2.6.0``, unpinned; 3.6.1 is what this image ships).  The published
algorithm is restated here over index arrays; the tie-breaking rules that make
the returned path unique are spelled out in SURVEY.md Appendix B:

  * one push counter shared by both directions orders equal-distance heap entries;
  * directions alternate strictly, forward first; a stale pop still uses up the turn;
  * neighbours are scanned in ``G._succ`` / ``G._pred`` insertion order;
  * relaxation and the meeting-point update both use strict ``<``;
  * the search stops when a node is finalised in both directions and the path goes
    through the recorded meeting node, not through the node just popped.

Pinned against live ``nx.shortest_path`` (tests/test_oracle_route.py) on tie-heavy
graphs, parallel edges, self-loops and unreachable pairs.
"""
from heapq import heappop, heappush


def bidirectional_route(pg, s, t):
    """Node-index path from s to t, or None when t is unreachable (NetworkXNoPath)."""
    if s == t:
        return [s]  # weighted.py:2393-2394
    adj = (pg.succ, pg.pred)
    final = ({}, {})
    seen = ({s: 0}, {t: 0})
    parent = ({s: -1}, {t: -1})
    heaps = ([(0, 0, s)], [(0, 1, t)])
    pushes = 2
    best = None
    meet = -1
    side = 1
    while heaps[0] and heaps[1]:
        side = 1 - side
        dist, _, v = heappop(heaps[side])
        mine = final[side]
        if v in mine:
            continue
        mine[v] = dist
        if v in final[1 - side]:
            fwd = []
            x = meet
            while x != -1:
                fwd.append(x)
                x = parent[0][x]
            fwd.reverse()
            x = parent[1][meet]
            while x != -1:
                fwd.append(x)
                x = parent[1][x]
            return fwd
        myseen, otherseen, mypar, myheap = seen[side], seen[1 - side], parent[side], heaps[side]
        for row in adj[side][v]:
            w, cost = row[0], row[1]
            cand = dist + cost
            if w in mine:
                continue
            old = myseen.get(w)
            if old is None or cand < old:
                myseen[w] = cand
                heappush(myheap, (cand, pushes, w))
                pushes += 1
                mypar[w] = v
                o = otherseen.get(w)
                if o is not None:
                    tot = cand + o
                    if best is None or best > tot:
                        best, meet = tot, w
    return None


def path_cost(pg, path):
    """Left-to-right fp64 sum of per-hop min travel_time."""
    c = 0.0
    for a, b in zip(path[:-1], path[1:]):
        for row in pg.succ[a]:
            if row[0] == b:
                c += row[1]
                break
    return c

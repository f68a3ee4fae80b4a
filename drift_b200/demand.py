"""Vectorised trip-chain generators for the device-driven demand mode (synthetic workloads).

In the reference every departure calls the s-t plugin once
(``get_source_target(agent_type, current_location, trip_count) -> (source, target)``,
lib/st_selection.py:23-35) and ``source`` is always the previous target (lib/agent.py:47-53).  A
trip chain is therefore just the sequence of targets; these generators produce, for all agents at
once, chains with the same structure as the five reference models so that the large BASELINE.json
configs can be driven without a Python call per departure.  They are workload generators, not
parity objects: bit-exact parity runs use the reference's own selector objects through
``drift_b200.simulation_thread`` or recorded traces.
"""
from __future__ import annotations

import numpy as np


def _no_self(rng, dst, prev, V):
    """Redraw targets equal to the trip's source (st_selection.py:47-49)."""
    bad = dst == prev
    while bad.any():
        dst[bad] = rng.integers(0, V, size=int(bad.sum()))
        bad = dst == prev
    return dst


def random_chains(rng, V, start, k):
    """RandomSelection (st_selection.py:37-51): uniform target != source."""
    N = len(start)
    dst = np.empty((N, k), dtype=np.int32)
    prev = np.asarray(start, dtype=np.int64).copy()
    for j in range(k):
        d = _no_self(rng, rng.integers(0, V, size=N), prev, V)
        dst[:, j] = d
        prev = d
    return np.arange(N + 1, dtype=np.int64) * k, dst.reshape(-1)


def hub_chains(rng, lg, start, k, hub_trip_probability=0.3, hub_percentage=0.1):
    """HubAndSpokeSelection (st_selection.py:469-648), vectorised: hubs = top ``hub_percentage`` nodes by degree
    (degree centrality only above 1 000 nodes, :498-511).  With probability ``hub_trip_probability`` a trip
    involves a hub -- from a hub to any other node, from a non-hub to a hub (:566-582) -- otherwise it ends
    at a non-hub (:584-606); the target never equals the source."""
    V = lg.V
    deg = np.diff(lg.fwd_rowptr) + np.diff(lg.bwd_rowptr)
    n_hubs = max(1, int(V * hub_percentage))
    order = np.argsort(-deg, kind="stable")
    hubs, non_hubs = order[:n_hubs], order[n_hubs:]
    is_hub = np.zeros(V, dtype=bool)
    is_hub[hubs] = True
    N = len(start)
    dst = np.empty((N, k), dtype=np.int32)
    prev = np.asarray(start, dtype=np.int64).copy()

    def draw(hub_trip, from_hub):
        n = len(hub_trip)
        anywhere = rng.integers(0, V, size=n)
        to_hub = hubs[rng.integers(0, n_hubs, size=n)]
        to_non = non_hubs[rng.integers(0, len(non_hubs), size=n)] if len(non_hubs) else anywhere
        return np.where(hub_trip, np.where(from_hub, anywhere, to_hub), to_non)

    for j in range(k):
        hub_trip = rng.random(N) < hub_trip_probability
        from_hub = is_hub[prev]
        d = draw(hub_trip, from_hub)
        bad = d == prev
        while bad.any():  # the reference removes the source from the candidate list before drawing
            d[bad] = draw(hub_trip[bad], from_hub[bad])
            bad = d == prev
        dst[:, j] = d
        prev = d
    return np.arange(N + 1, dtype=np.int64) * k, dst.reshape(-1)


def _activity_pools(rng, lg):
    """The node pools of ActivityBasedSelection (st_selection.py:690-890): sizes capped by config.py:294-303
    (activity 50 / business 80 / work 100 / home 100 / distribution 20), drawn by distance ring from the
    network centre (centre <= 0.3, middle <= 0.7 of the largest distance)."""
    V = lg.V
    pos = lg.pos if lg.pos is not None else rng.random((V, 2))
    c = 0.5 * (pos.min(axis=0) + pos.max(axis=0))
    r = np.hypot(pos[:, 0] - c[0], pos[:, 1] - c[1])
    rn = r / max(r.max(), 1e-12)
    centre = np.nonzero(rn <= 0.3)[0]
    middle = np.nonzero((rn > 0.3) & (rn <= 0.7))[0]
    border = np.nonzero(rn > 0.7)[0]

    def pick(pool, n):
        pool = pool if len(pool) else np.arange(V)
        return rng.choice(pool, size=min(n, len(pool)), replace=False)

    return dict(homes=np.concatenate([pick(centre, 5), pick(middle, 35), pick(border, 60)]),
                works=np.concatenate([pick(centre, 20), pick(middle, 40), pick(border, 40)]),
                business=np.concatenate([pick(centre, 56), pick(middle, 24)]),
                leisure=np.concatenate([pick(centre, 40), pick(middle, 10)]),
                depots=np.concatenate([pick(middle, 10), pick(border, 10)]))


def activity_chains(rng, lg, start, k, shares=(0.4, 0.3, 0.2, 0.1)):
    """ActivityBasedSelection (st_selection.py:650-1082), trip law by agent type (shares commuter / leisure /
    business / delivery = .4/.3/.2/.1, config.py:323-328), vectorised over the agents:
      commuter  even trips -> a work node, odd trips -> a home node (:941-955)
      delivery  70 % -> a distribution node, else any node (:972-975)
      leisure   60 % -> an activity centre, else any node (:1000-1003)
      business  80 % -> a business node, else any node (:1034-1037)
    and a target equal to the source is redrawn (:978-984, 1006-1016, 1040-1048).  About a fifth of the
    trips therefore end at an arbitrary node, and ~6 % run between two arbitrary nodes."""
    V = lg.V
    P = _activity_pools(rng, lg)
    N = len(start)
    typ = rng.choice(4, size=N, p=list(shares))  # 0 commuter, 1 leisure, 2 business, 3 delivery
    com, lei, bus, dly = typ == 0, typ == 1, typ == 2, typ == 3
    dst = np.empty((N, k), dtype=np.int32)
    prev = np.asarray(start, dtype=np.int64).copy()

    def mixed(mask, pool, p_pool):
        n = int(mask.sum())
        to_pool = rng.random(n) < p_pool
        return np.where(to_pool, pool[rng.integers(0, len(pool), size=n)], rng.integers(0, V, size=n))

    for j in range(k):
        d = np.empty(N, dtype=np.int64)
        pool = P["works"] if j % 2 == 0 else P["homes"]
        d[com] = pool[rng.integers(0, len(pool), size=int(com.sum()))]
        d[lei] = mixed(lei, P["leisure"], 0.6)
        d[bus] = mixed(bus, P["business"], 0.8)
        d[dly] = mixed(dly, P["depots"], 0.7)
        for _ in range(100):  # redraws (commuters keep a target equal to the source, like the reference)
            bad = (d == prev) & ~com
            if not bad.any():
                break
            d[bad & lei] = mixed(bad & lei, P["leisure"], 0.6)
            d[bad & bus] = mixed(bad & bus, P["business"], 0.8)
            d[bad & dly] = rng.integers(0, V, size=int((bad & dly).sum()))  # :981: any node
        dst[:, j] = d
        prev = d
    return np.arange(N + 1, dtype=np.int64) * k, dst.reshape(-1)


def activity_pools_chains(rng, lg, start, k):
    """Pool-to-pool variant of the activity model (every trip has at least one end in a small pool; only the
    delivery out-legs go to arbitrary nodes): fixed home / work per agent, out legs on even trips, return legs
    on odd ones.  The best case for the tree cache; kept as a separate workload for comparison."""
    V = lg.V
    P = _activity_pools(rng, lg)
    homes, works, business, leisure, depots = P["homes"], P["works"], P["business"], P["leisure"], P["depots"]
    N = len(start)
    typ = rng.choice(4, size=N, p=[0.4, 0.3, 0.2, 0.1])
    home = homes[rng.integers(0, len(homes), size=N)]
    work = works[rng.integers(0, len(works), size=N)]
    dst = np.empty((N, k), dtype=np.int32)
    for j in range(k):
        out = j % 2 == 0  # even legs go out, odd legs return home
        d = np.empty(N, dtype=np.int64)
        com, lei, bus, dly = typ == 0, typ == 1, typ == 2, typ == 3
        d[com] = work[com] if out else home[com]
        d[lei] = leisure[rng.integers(0, len(leisure), size=int(lei.sum()))] if out else home[lei]
        d[bus] = business[rng.integers(0, len(business), size=int(bus.sum()))]
        d[dly] = depots[rng.integers(0, len(depots), size=int(dly.sum()))] if not out else \
            rng.integers(0, V, size=int(dly.sum()))
        dst[:, j] = d
    return np.arange(N + 1, dtype=np.int64) * k, dst.reshape(-1)


def zone_chains(rng, lg, start, k, intra_zone_probability=0.7, grid=3):
    """ZoneBasedSelection (st_selection.py:1084-1252), vectorised: 3 x 3 position zones; a trip stays in the
    source's zone with probability 0.7 (config.py:283, any other node of the zone, :1186-1197); otherwise it goes
    to a uniformly chosen OTHER zone, with probability 0.7 to one of that zone's major nodes (its top third
    by degree, :1142-1160) else to any of its nodes (:1199-1231).  The target never equals the source."""
    V = lg.V
    pos = lg.pos if lg.pos is not None else rng.random((V, 2))
    lo, hi = pos.min(axis=0), pos.max(axis=0)
    cell = np.minimum(((pos - lo) / np.maximum(hi - lo, 1e-12) * grid).astype(np.int64), grid - 1)
    zone = cell[:, 0] * grid + cell[:, 1]  # :1134-1136: x cell * grid + y cell
    Z = grid * grid
    deg = np.diff(lg.fwd_rowptr) + np.diff(lg.bwd_rowptr)
    order = np.lexsort((np.arange(V), -deg, zone))  # by zone, then degree (descending), then node id
    zstart = np.searchsorted(zone[order], np.arange(Z + 1))
    zsize = np.diff(zstart)
    major = np.maximum(1, zsize // 3)
    non_empty = np.nonzero(zsize > 0)[0]
    N = len(start)
    dst = np.empty((N, k), dtype=np.int32)
    prev = np.asarray(start, dtype=np.int64).copy()

    def draw(src):
        n = len(src)
        zs = zone[src]
        intra = (rng.random(n) < intra_zone_probability) & (zsize[zs] > 1)  # alone in its zone: inter-zone trip
        other = non_empty[rng.integers(0, len(non_empty), size=n)]
        for _ in range(64):  # a different, non-empty zone
            same = other == zs
            if not same.any() or len(non_empty) < 2:
                break
            other[same] = non_empty[rng.integers(0, len(non_empty), size=int(same.sum()))]
        z = np.where(intra, zs, other)
        to_major = ~intra & (rng.random(n) < 0.7)
        span = np.where(to_major, major[z], zsize[z])
        return order[zstart[z] + (rng.random(n) * span).astype(np.int64)]

    for j in range(k):
        d = draw(prev)
        bad = d == prev
        for _ in range(100):
            if not bad.any():
                break
            d[bad] = draw(prev[bad])
            bad = d == prev
        dst[:, j] = d
        prev = d
    return np.arange(N + 1, dtype=np.int64) * k, dst.reshape(-1)


def gravity_chains(rng, lg, start, k, alpha=1.0, beta=2.0, candidates=32):
    """GravitySelection (st_selection.py:53-466) structure, sampled: T_ij ~ A_j / d_ij^beta with
    A_j = 0.01 + alpha * degree centrality (:288-303) and Euclidean d_ij (:327-363); the target is
    drawn among ``candidates`` uniformly pre-drawn nodes instead of all V (the exact model is
    O(V) per draw)."""
    V = lg.V
    pos = lg.pos if lg.pos is not None else rng.random((V, 2))
    deg = (np.diff(lg.fwd_rowptr) + np.diff(lg.bwd_rowptr)) / max(V - 1, 1)
    attr = 0.01 + alpha * deg
    N = len(start)
    dst = np.empty((N, k), dtype=np.int32)
    prev = np.asarray(start, dtype=np.int64).copy()
    for j in range(k):
        cand = rng.integers(0, V, size=(N, candidates))
        dx = pos[cand, 0] - pos[prev, 0][:, None]
        dy = pos[cand, 1] - pos[prev, 1][:, None]
        dij = np.maximum(np.hypot(dx, dy), 0.1)
        w = attr[cand] / dij ** beta
        w[cand == prev[:, None]] = 0.0
        cw = np.cumsum(w, axis=1)
        u = rng.random(N) * cw[:, -1]
        pickj = np.minimum((cw < u[:, None]).sum(axis=1), candidates - 1)
        d = cand[np.arange(N), pickj]
        dst[:, j] = d
        prev = d
    return np.arange(N + 1, dtype=np.int64) * k, dst.reshape(-1)


MODELS = {
    "random": lambda rng, lg, start, k: random_chains(rng, lg.V, start, k),
    "hub": hub_chains,
    "activity": activity_chains,
    "activity_pools": activity_pools_chains,
    "zone": zone_chains,
    "gravity": gravity_chains,
}

"""TEST INFRASTRUCTURE ONLY -- sequential CPU restatement ("port") of DRIFT's step.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
(``drift_b200/``) never does.

Parity status: PINNED BY EXECUTION.  The reference ships no tests or golden
vectors for this path, so this port is pinned against the
live reference run in the build container: ``oracle/make_golden.py`` records
the reference (``oracle/ref_harness.py``) and ``tests/test_oracle_*.py`` check
this file against those fixtures (and against the live reference when
``/root/reference`` is present).

What is restated (reference file:line):
  * tick loop, agent order, trip bookkeeping   lib/simulation_thread.py:288-355, 384-390, 483-489
  * waiting/moving state machine                lib/agent.py:171-203, 243-259
  * departure + routing + link entry            lib/agent.py:46-160
  * occupancy lists, capacities, BPR factor     lib/managers/traffic_manager.py:33-107
  * trip records                                lib/data_logger.py:20-120
  * shortest path                               networkx 3.6.1 algorithms/shortest_paths/weighted.py:2387-2459
                                                 (third-party; restated in oracle/nx_route.py)

The port keeps the reference's *sequential* (Gauss-Seidel, agent-list order)
semantics on purpose: the CUDA path uses a parallel formulation (events ordered
by (link, agent) + running count) and must be shown equal to this.
"""
from __future__ import annotations

import datetime
import random as _random

import numpy as np

from .nx_route import bidirectional_route

# constants (reference config.py:178-190, 275-276, 336-340)
SECONDS_PER_HOUR = 3600
START_OFFSET = 21600
DEFAULT_SPEED_KPH = 30.0
DEFAULT_EDGE_LENGTH = 1000.0
KPH_TO_MPS = 1000 / 3600
SIMULATION_DT = 0.1
DEFAULT_HOURLY = {
    0: 0.05, 1: 0.02, 2: 0.01, 3: 0.01, 4: 0.02, 5: 0.05,
    6: 0.15, 7: 0.25, 8: 0.35, 9: 0.20, 10: 0.15, 11: 0.18,
    12: 0.22, 13: 0.18, 14: 0.16, 15: 0.20, 16: 0.25, 17: 0.30,
    18: 0.35, 19: 0.25, 20: 0.15, 21: 0.12, 22: 0.08, 23: 0.06,
}
START_DATETIME = datetime.datetime(2025, 7, 10, 6, 0, 0)  # simulation_thread.py:154


class PlainGraph:
    """Index-based view of a loader-processed ``nx.MultiDiGraph`` (oracle's own lowering).

    nodes           ``list(G.nodes)``
    links           ``list(G.edges(keys=True))`` -> link id = position in that list
    link_len/v0/tt  float64[L]  ('length', 'speed_kph', 'travel_time' with the reference's defaults)
    link_cap        int[L]      traffic_manager.py:33-61
    succ[u]         [(v, w_min, link_id, len_min)] in ``G._succ[u]`` order
    pred[v]         [(u, w_min)]                    in ``G._pred[v]`` order
    where w_min = min travel_time over parallel (u,v,*) edges (nx weighted.py:76-77),
    link_id = first key with minimal travel_time (agent.py:122-130) and
    len_min = min 'length' over the parallel edges (data_logger.py:113-114).
    """

    def __init__(self, nodes, links, link_len, link_v0, link_tt, link_cap, succ, pred):
        self.nodes = nodes
        self.links = links
        self.link_len = link_len
        self.link_v0 = link_v0
        self.link_tt = link_tt
        self.link_cap = link_cap
        self.succ = succ
        self.pred = pred
        self.node_index = {n: i for i, n in enumerate(nodes)}
        self.V = len(nodes)
        self.L = len(links)
        self.hop = [dict((v, (lid, lmin)) for v, _, lid, lmin in row) for row in succ]


def link_capacity(lanes, speed_kph, length_m):
    """traffic_manager.py:50-61"""
    base = 2000 if speed_kph >= 50 else 1500
    theoretical = int(lanes * base)
    storage = max(1, int(length_m / 7.5))
    return min(theoretical, storage * 4)


def lower_nx(G) -> PlainGraph:
    nodes = list(G.nodes)
    idx = {n: i for i, n in enumerate(nodes)}
    links = list(G.edges(keys=True))
    lid = {e: i for i, e in enumerate(links)}
    L = len(links)
    link_len = np.empty(L)
    link_v0 = np.empty(L)
    link_tt = np.empty(L)
    link_cap = np.empty(L, dtype=np.int64)
    for i, (u, v, k) in enumerate(links):
        data = G[u][v][k]
        try:  # agent.py:145-151
            link_len[i] = float(data.get("length", DEFAULT_EDGE_LENGTH))
            link_v0[i] = float(data.get("speed_kph", DEFAULT_SPEED_KPH))
        except (TypeError, ValueError):
            link_len[i] = DEFAULT_EDGE_LENGTH
            link_v0[i] = DEFAULT_SPEED_KPH
        link_tt[i] = data.get("travel_time", 1)
        try:  # traffic_manager.py:40-48
            lanes = float(data.get("lanes", 1))
            sp = float(data.get("speed_kph", DEFAULT_SPEED_KPH))
            ln = float(data.get("length", DEFAULT_EDGE_LENGTH))
        except (TypeError, ValueError):
            lanes, sp, ln = 1, DEFAULT_SPEED_KPH, DEFAULT_EDGE_LENGTH
        link_cap[i] = link_capacity(lanes, sp, ln)
    succ = []
    for u in nodes:
        row = []
        for v, par in G._succ[u].items():
            wmin = min(a.get("travel_time", 1) for a in par.values())
            best_key, best_tt = None, None
            for k, a in par.items():
                tt = a.get("travel_time", float("inf"))
                if best_tt is None or tt < best_tt:
                    best_key, best_tt = k, tt
            lmin = min(par.values(), key=lambda a: a.get("length", float("inf"))).get(
                "length", DEFAULT_EDGE_LENGTH)
            row.append((idx[v], wmin, lid[(u, v, best_key)], lmin))
        succ.append(row)
    pred = []
    for v in nodes:
        row = []
        for u, par in G._pred[v].items():
            wmin = min(a.get("travel_time", 1) for a in par.values())
            row.append((idx[u], wmin))
        pred.append(row)
    return PlainGraph(nodes, links, link_len, link_v0, link_tt, link_cap, succ, pred)


def bpr_factor(count, capacity, alpha, beta):
    """traffic_manager.py:88-107 (``**`` on Python floats = libm pow)."""
    if count == 0:
        return 1.0
    r = count / capacity
    m = 1 + alpha * (r ** beta)
    return max(0.1, 1 / m)


def travel_probability(t, hourly):
    """agent.py:243-259"""
    hour = int(((t + START_OFFSET) / SECONDS_PER_HOUR) % 24)
    return hourly.get(hour, 0.05) * 4.0


class RandomSelector:
    """Same draws as the reference's RandomSelection (st_selection.py:41-51)."""

    def __init__(self, nodes):
        self.nodes = list(nodes)

    def get_source_target(self, agent_type=None, current_location=None, trip_count=0):
        src = current_location if current_location is not None else _random.choice(self.nodes)
        while True:
            dst = _random.choice(self.nodes)
            if dst != src:
                return src, dst


class _A:
    __slots__ = ("idx", "state", "agent_type", "source", "target", "trip_count", "path", "pidx",
                 "d", "speed", "elen", "link", "trip", "has_path")


class PortSim:
    """Sequential restatement of SimulationThread + Agent + TrafficManager + data logger.

    ``departure_trace`` (optional) = {tick: {agent_idx: (src_idx, dst_idx, path_idx_list|None)}}:
    replay mode, no RNG is consumed; otherwise the global ``random`` stream is
    consumed exactly as the reference does (H4 in SURVEY.md).
    """

    def __init__(self, pg: PlainGraph, num_agents, selector=None, hourly=None, alpha=0.15, beta=4.0,
                 accel=10, hours=1, departure_trace=None, init_nodes=None, agent_types=None,
                 route_fn=None, record_entries=False, chains=None):
        self.pg = pg
        self.alpha, self.beta = alpha, beta
        self.hourly = DEFAULT_HOURLY if hourly is None else hourly
        self.dt = SIMULATION_DT
        self.accel = accel
        self.max_time = hours * SECONDS_PER_HOUR
        self.t = 0.0
        self.step = 0
        self.trace = departure_trace
        self.selector = selector
        self.chains = chains  # (trip_off, trip_dst): host-generated target chains instead of a selector
        self.chain_pos = [0] * num_agents
        self.route_fn = route_fn or (lambda s, t: bidirectional_route(pg, s, t))
        self.occ = [0] * pg.L
        self.trips = []
        self.trip_counter = 0
        self.entries = [] if record_entries else None
        self.arrivals = []
        self.departures = []
        self.routes_computed = 0
        self.agents = []
        for i in range(num_agents):
            a = _A()
            a.idx = i
            a.state = 0
            a.agent_type = agent_types[i] if agent_types is not None else "random"
            if init_nodes is not None:
                n0 = int(init_nodes[i])
            else:
                n0 = pg.node_index[_random.choice(pg.nodes)]  # agent.py:40
            a.source = a.target = n0
            a.trip_count = 0
            a.path = None
            a.has_path = False
            a.pidx = 0
            a.d = 0.0
            a.speed = 0.0
            a.elen = 0.0
            a.link = -1
            a.trip = None
            self.agents.append(a)
        for a in self.agents:  # simulation_thread.py:274-276
            self._start_trip(a, 0.0)

    # ---- trip records (data_logger.py:20-101) ---------------------------
    def _stamp(self, t):
        return (START_DATETIME + datetime.timedelta(seconds=t)).strftime("%H:%M:%S")

    def _start_trip(self, a, t):
        self.trip_counter += 1
        a.trip = dict(trip_id="T%06d" % self.trip_counter, agent_index=a.idx, start_time=self._stamp(t),
                      origin=a.source, destination=a.target, route=[a.source], t0=t)

    def _end_trip(self, a, t):
        tr = a.trip
        if tr is None:
            return
        dur_min = (t - tr["t0"]) / 60.0
        if a.has_path and a.path is None:
            raise TypeError("object of type 'NoneType' has no len()")  # data_logger.py:68 quirk
        if a.has_path and len(a.path) > 1:
            total = 0.0
            for i in range(len(a.path) - 1):
                total += self.pg.hop[a.path[i]][a.path[i + 1]][1]
            km = total / 1000.0
        else:
            km = 0.0
        kmh = (km / dur_min) * 60 if dur_min > 0 else 0.0
        if a.target not in tr["route"]:
            tr["route"].append(a.target)
        nodes = self.pg.nodes
        self.trips.append(dict(
            trip_id=tr["trip_id"], agent_index=a.idx, start_time=tr["start_time"], end_time=self._stamp(t),
            origin_node=nodes[tr["origin"]], destination_node=nodes[tr["destination"]],
            route_taken=str([nodes[i] for i in tr["route"]]),
            trip_distance_km=km, trip_duration_min=dur_min, average_speed_kmh=kmh))
        a.trip = None

    # ---- agent (agent.py:46-203) ---------------------------------------
    def _enter_or_arrive(self, a):
        if a.pidx >= len(a.path) - 1:
            if a.link >= 0:
                self.occ[a.link] -= 1
                a.link = -1
            a.state = 0
            return
        lid, _ = self.pg.hop[a.path[a.pidx]][a.path[a.pidx + 1]]
        if a.link >= 0:
            self.occ[a.link] -= 1
        self.occ[lid] += 1
        a.link = lid
        a.elen = float(self.pg.link_len[lid])
        f = bpr_factor(self.occ[lid], int(self.pg.link_cap[lid]), self.alpha, self.beta)
        a.speed = (float(self.pg.link_v0[lid]) * f) * KPH_TO_MPS
        if self.entries is not None:
            self.entries.append((self.step, a.idx, lid, a.speed, a.elen))

    def _depart(self, a):
        pg = self.pg
        if self.trace is not None:
            s, t, path = self.trace[self.step][a.idx]
        elif self.chains is not None:
            off, dst = self.chains
            s, t = a.target, int(dst[off[a.idx] + self.chain_pos[a.idx]])
            self.chain_pos[a.idx] += 1
            path = "route"
        else:
            sl, tl = self.selector.get_source_target(a.agent_type, pg.nodes[a.target], a.trip_count)
            s, t = pg.node_index.get(sl, -1), pg.node_index.get(tl, -1)
            sel = self.selector  # agent.py:56-68: hub / zone models relabel the agent
            if hasattr(sel, "last_trip_type") and type(sel).__name__ == "HubAndSpokeSelection":
                a.agent_type = sel.last_trip_type
            elif hasattr(sel, "node_to_zone") and type(sel).__name__ == "ZoneBasedSelection":
                zs, zt = sel.node_to_zone.get(sl), sel.node_to_zone.get(tl)
                if zs is not None and zt is not None:
                    a.agent_type = "intra_zone" if zs == zt else "inter_zone"
            path = "route"
        a.source, a.target = s, t
        a.trip_count += 1
        a.has_path = True
        if path == "route":
            path = None if (s < 0 or t < 0) else self.route_fn(s, t)
            self.routes_computed += 1
        self.departures.append((self.step, a.idx, s, t, path))
        if path is None:
            a.path = None
            a.state = 0
            return
        a.path = path
        a.pidx = 0
        a.d = 0.0
        self._enter_or_arrive(a)

    def tick(self):
        """simulation_thread.py:288-340; returns False once the horizon is reached."""
        if self.t >= self.max_time:
            return False
        delta = self.dt * self.accel
        self.t += delta
        self.step += 1
        t = self.t
        trace_now = self.trace.get(self.step) if self.trace is not None else None
        p_dep = travel_probability(t, self.hourly) * delta / SECONDS_PER_HOUR
        for a in self.agents:
            old = a.state
            if old == 0:
                if self.trace is not None:
                    go = trace_now is not None and a.idx in trace_now
                elif self.chains is not None and self.chains[0][a.idx] + self.chain_pos[a.idx] >= self.chains[0][a.idx + 1]:
                    go = False  # chain exhausted
                else:
                    go = _random.random() < p_dep
                if go:
                    a.state = 1
                    self._depart(a)
            else:
                a.d += a.speed * delta
                if a.d >= a.elen:
                    a.pidx += 1
                    a.d = 0.0
                    self._enter_or_arrive(a)
            if old == 0 and a.state == 1:
                if a.trip is None:
                    self._start_trip(a, t)
            elif old == 1 and a.state == 0:
                if a.trip is not None:
                    self._end_trip(a, t)
                self.arrivals.append((self.step, a.idx))
            if a.state == 1 and a.trip is not None:
                node = a.path[a.pidx]
                # data_logger.py:52 -- falsy labels (int 0) are skipped
                if self.pg.nodes[node] and node not in a.trip["route"]:
                    a.trip["route"].append(node)
        return True

    def viz_snapshot(self):
        """simulation_thread.py:394-416 (_prepare_agent_data): what the reference does on EVERY tick as shipped
        (UPDATE_FREQUENCY = 1, config.py:191) -- a throw-away class per agent carrying the fields the widget reads.
        ~80 % of the reference's wall time; the headless rate switches it off."""
        out = []
        t = self.t
        for a in self.agents:
            rec = type("Agent", (), {
                "id": a.idx + 1, "position": (0, 0), "state": "moving" if a.state else "waiting",
                "agent_type": a.agent_type, "trip_count": a.trip_count, "speed": a.speed, "source": a.source,
                "target": a.target, "path": a.path if a.has_path else [], "path_index": a.pidx,
                "current_node": None, "distance_on_edge": a.d, "edge_length": a.elen})
            rec.simulation_time = t  # :311-312
            out.append(rec)
        return out

    # ---- extension: congestion-aware rerouting (no reference behaviour; SURVEY.md 8c last row) -----
    # Composed from reference pieces: the BPR factor of traffic_manager.py:92-107 turned into a link
    # travel time tt = t0 / factor(occupancy), NetworkX-order bidirectional Dijkstra on those
    # weights, new route spliced after the current link.  A (u, v) pair keeps the link chosen on
    # free-flow travel times (differs from a literal NetworkX composition only on graphs with
    # parallel edges).
    def update_travel_times(self):
        pg = self.pg
        tt = [float(pg.link_tt[l]) / bpr_factor(self.occ[l], int(pg.link_cap[l]), self.alpha, self.beta)
              for l in range(pg.L)]
        self.tt = tt
        succ = [[(v, tt[lid], lid, lmin) for v, _w, lid, lmin in row] for row in pg.succ]
        pair = [{v: w for v, w, _l, _m in row} for row in succ]
        pred = [[(u, pair[u][v]) for u, _w in row] for v, row in enumerate(pg.pred)]

        class _G:
            pass

        g = _G()
        g.succ, g.pred = succ, pred
        self.route_fn = lambda s, t, _g=g: bidirectional_route(_g, s, t)

    def reroute_moving(self):
        n = 0
        for a in self.agents:
            if a.state != 1 or a.pidx >= len(a.path) - 2:
                continue  # not moving, or already on its last link
            new = self.route_fn(a.path[a.pidx + 1], a.target)
            n += 1
            if new is not None and len(new) > 1:
                a.path = a.path[:a.pidx + 1] + new
        return n

    def finish(self):
        """simulation_thread.py:483-489"""
        for a in self.agents:
            if a.trip is not None:
                self._end_trip(a, self.t)

    # ---- views used by tests -------------------------------------------
    def snapshot(self):
        n = len(self.agents)
        st = np.fromiter((a.state for a in self.agents), dtype=np.uint8, count=n)
        d = np.fromiter((a.d for a in self.agents), dtype=np.float64, count=n)
        sp = np.fromiter((a.speed for a in self.agents), dtype=np.float64, count=n)
        pi = np.fromiter((a.pidx for a in self.agents), dtype=np.int32, count=n)
        return st, d, sp, pi


def plain_from_arrays(nodes, link_u, link_v, link_len, link_v0, link_tt, link_cap, fwd_rowptr, fwd_col, fwd_w,
                      fwd_link, fwd_minlen, bwd_rowptr, bwd_col, bwd_w):
    """PlainGraph from CSR arrays (synthetic networks that never existed as a NetworkX object)."""
    V = len(fwd_rowptr) - 1
    succ = [[(int(fwd_col[e]), float(fwd_w[e]), int(fwd_link[e]), float(fwd_minlen[e]))
             for e in range(fwd_rowptr[u], fwd_rowptr[u + 1])] for u in range(V)]
    pred = [[(int(bwd_col[e]), float(bwd_w[e])) for e in range(bwd_rowptr[v], bwd_rowptr[v + 1])] for v in range(V)]
    links = [(nodes[u], nodes[v], 0) for u, v in zip(link_u.tolist(), link_v.tolist())]
    return PlainGraph(list(nodes), links, np.asarray(link_len, dtype=np.float64), np.asarray(link_v0, dtype=np.float64),
                      np.asarray(link_tt, dtype=np.float64), np.asarray(link_cap, dtype=np.int64), succ, pred)

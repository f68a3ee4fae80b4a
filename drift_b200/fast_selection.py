"""Vectorised s-t selectors behind the reference's plugin interface.

Same constructor arguments, same ``get_source_target(agent_type, current_location, trip_count)``,
same attributes the application reads (``hubs`` / ``non_hubs`` / ``last_trip_type``, ``zones`` /
``node_to_zone`` / ``major_nodes_by_zone``, ``size_variables`` / ``attraction_variables`` /
``node_importance``) and -- seeded the same way -- **the same (source, target) sequence** as
lib/st_selection.py, because every draw goes through the global ``random`` module in the
reference's order with the reference's arguments:

* ``random.choice(seq)`` is ``seq[_randbelow(len(seq))]``: the list comprehensions the reference
  rebuilds on every call (``[n for n in self.nodes if n != source]``, lib/st_selection.py:572, 592,
  1194, 1212-1229) only ever drop the source, so the chosen element is found by index arithmetic;
* ``random.choices(population, weights)`` is one ``random()`` and a bisection over the running sums of
  the weights (``itertools.accumulate`` = sequential float adds = ``numpy.cumsum``): the gravity
  model's O(V) Python loop per call (:408-423) becomes one vector expression per *source*, cached,
  and its O(V^2) distance table (:305-363) is never built -- rows are computed when first needed.
  ``d ** beta`` and ``(dx*dx + dy*dy) ** 0.5`` go through the C library's ``pow`` (``drift_host_pow``)
  like CPython's float power; NumPy's own power differs in the last bit for ~0.1-5 % of inputs.

The class names are the reference's on purpose: lib/agent.py:57,61 switches on
``st_selector.__class__.__name__``.  ``tests/test_fast_selection.py`` checks the sequences against
fixtures recorded from the reference classes (and live against them where the reference tree exists).
"""
from __future__ import annotations

import random
from collections import OrderedDict

import networkx as nx
import numpy as np


def _host_pow(x, y):
    """x ** y element-wise exactly as CPython computes it (C library pow)."""
    from . import _capi

    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _capi.check(_capi.load().drift_host_pow(x.ctypes.data_as(_capi.c_f64p), float(y),
                                            out.ctypes.data_as(_capi.c_f64p), x.size))
    return out


def _pick_except(seq, index_of, banned):
    """``random.choice([n for n in seq if n != banned])`` without building the list (None if empty)."""
    pos = index_of.get(banned, -1)
    n = len(seq) - (1 if pos >= 0 else 0)
    if n <= 0:
        return None
    k = random.choice(range(n))  # same _randbelow(n) as random.choice(list of n)
    return seq[k if pos < 0 or k < pos else k + 1]


class _Base:
    def __init__(self, graph):
        self.graph = graph
        self.nodes = list(graph.nodes())  # lib/st_selection.py:21
        self._node_pos = {n: i for i, n in enumerate(self.nodes)}

    def _any_other(self, source):
        """``random.choice([n for n in self.nodes if n != source])``"""
        return _pick_except(self.nodes, self._node_pos, source)


class HubAndSpokeSelection(_Base):
    """lib/st_selection.py:469-648 with O(1) draws."""

    def __init__(self, graph, config=None, hub_trip_probability=None, hub_percentage=None):
        super().__init__(graph)
        self.config = config
        self.hub_trip_probability = hub_trip_probability if hub_trip_probability is not None else \
            getattr(config, "DEFAULT_HUB_TRIP_PROBABILITY", 0.3)
        self.hub_percentage = hub_percentage if hub_percentage is not None else \
            getattr(config, "DEFAULT_HUB_PERCENTAGE", 0.1)
        degree = nx.degree_centrality(graph)
        if len(self.nodes) <= 1000:  # :498-511
            between = nx.betweenness_centrality(graph)
            self.node_centrality = {n: 0.6 * degree.get(n, 0) + 0.4 * between.get(n, 0) for n in self.nodes}
        else:
            self.node_centrality = degree
        self._split()

    def _split(self):
        order = sorted(self.nodes, key=lambda x: self.node_centrality[x], reverse=True)  # :514
        num_hubs = max(1, int(len(self.nodes) * self.hub_percentage))
        self.hubs, self.non_hubs = order[:num_hubs], order[num_hubs:]
        self._hub_pos = {n: i for i, n in enumerate(self.hubs)}
        self._non_hub_pos = {n: i for i, n in enumerate(self.non_hubs)}

    def get_source_target(self, agent_type=None, current_location=None, trip_count=0):
        source = random.choice(self.nodes) if current_location is None else current_location
        if random.random() < self.hub_trip_probability:  # :556
            if source in self._hub_pos:
                target = self._any_other(source)
            else:
                target = _pick_except(self.hubs, self._hub_pos, source)
                if target is None:
                    target = self._any_other(source)
            self.last_trip_type = "hub"
        else:
            target = _pick_except(self.non_hubs, self._non_hub_pos, source)
            if target is None:
                target = self._any_other(source)
            self.last_trip_type = "non-hub"
        if target is None:
            raise IndexError("Cannot choose from an empty sequence")  # what random.choice([]) raises
        return source, target

    def set_parameters(self, hub_trip_probability=None, hub_percentage=None):
        if hub_trip_probability is not None:
            self.hub_trip_probability = max(0.0, min(1.0, hub_trip_probability))
        if hub_percentage is not None:
            self.hub_percentage = max(0.01, min(1.0, hub_percentage))
            self._split()

    def get_model_info(self):
        c = list(self.node_centrality.values())
        return {"model_type": "hub_and_spoke", "hub_trip_probability": self.hub_trip_probability,
                "hub_percentage": self.hub_percentage, "num_nodes": len(self.nodes), "num_hubs": len(self.hubs),
                "num_non_hubs": len(self.non_hubs), "max_centrality": max(c), "min_centrality": min(c),
                "avg_centrality": np.mean(c)}

    def get_hub_nodes(self):
        return {"hubs": self.hubs, "non_hubs": self.non_hubs, "centrality": self.node_centrality}


class ZoneBasedSelection(_Base):
    """lib/st_selection.py:1084-1252 with O(1) draws."""

    def __init__(self, graph, config=None, intra_zone_probability=None):
        super().__init__(graph)
        self.config = config
        self.intra_zone_probability = intra_zone_probability if intra_zone_probability is not None else \
            getattr(config, "DEFAULT_ZONE_INTRA_PROBABILITY", 0.7)
        self.grid_size = getattr(config, "ZONE_GRID_SIZE", 3)
        pos = {n: graph.nodes[n]["pos"] for n in self.nodes}  # :1107-1109
        xs = [p[0] for p in pos.values()]
        ys = [p[1] for p in pos.values()]
        min_x, max_x, min_y, max_y = min(xs), max(xs), min(ys), max(ys)
        x_step = (max_x - min_x) / self.grid_size
        y_step = (max_y - min_y) / self.grid_size
        g = self.grid_size
        self.zones = {i * g + j: [] for i in range(g) for j in range(g)}
        self.node_to_zone = {}
        for n, p in pos.items():  # :1130-1137
            zone_id = min(g - 1, int((p[0] - min_x) / x_step)) * g + min(g - 1, int((p[1] - min_y) / y_step))
            self.zones[zone_id].append(n)
            self.node_to_zone[n] = zone_id
        self.major_nodes_by_zone = {}
        for zone_id, members in self.zones.items():  # :1142-1160
            if not members:
                self.major_nodes_by_zone[zone_id] = []
                continue
            deg = {n: graph.degree(n) for n in members}
            self.major_nodes_by_zone[zone_id] = sorted(members, key=lambda x: deg[x], reverse=True)[:max(1, len(members) // 3)]
        self._zone_pos = {z: {n: i for i, n in enumerate(m)} for z, m in self.zones.items()}
        self._major_pos = {z: {n: i for i, n in enumerate(m)} for z, m in self.major_nodes_by_zone.items()}
        self._zone_ids = list(self.zones.keys())
        self._zone_id_pos = {z: i for i, z in enumerate(self._zone_ids)}

    def get_source_target(self, agent_type=None, current_location=None, trip_count=0):
        source = random.choice(self.nodes) if current_location is None else current_location
        source_zone = self.node_to_zone[source]
        if random.random() < self.intra_zone_probability:  # :1177
            target = _pick_except(self.zones[source_zone], self._zone_pos[source_zone], source)
            if target is None:
                target = self._inter_zone(source_zone, source)
        else:
            target = self._inter_zone(source_zone, source)
        return source, target

    def _inter_zone(self, source_zone, source):
        dest = _pick_except(self._zone_ids, self._zone_id_pos, source_zone)  # :1204-1207
        if dest is None:
            raise IndexError("Cannot choose from an empty sequence")
        major = self.major_nodes_by_zone[dest]
        use_major = random.random() < 0.7 and bool(major)  # :1210 (random() is drawn first, like `and`)
        seq, idx = (major, self._major_pos[dest]) if use_major else (self.zones[dest], self._zone_pos[dest])
        # the three fallbacks of :1217-1223 test emptiness before drawing
        if len(seq) - (1 if source in idx else 0) <= 0:
            seq, idx = self.zones[dest], self._zone_pos[dest]
        if len(seq) - (1 if source in idx else 0) <= 0:
            seq, idx = self.nodes, self._node_pos
        target = _pick_except(seq, idx, source)
        if target is None:
            raise IndexError("Cannot choose from an empty sequence")
        return target

    def set_parameters(self, intra_zone_probability=None):
        if intra_zone_probability is not None:
            self.intra_zone_probability = max(0.0, min(1.0, intra_zone_probability))

    def get_zone_info(self):
        return {z: {"total_nodes": len(m), "major_nodes": len(self.major_nodes_by_zone[z]), "nodes": m[:5]}
                for z, m in self.zones.items()}


class GravitySelection(_Base):
    """lib/st_selection.py:53-466 (Voorhees gravity model) with one vector expression per source.

    ``row_cache`` bounds the number of cached cumulative-weight rows (8 B x V each)."""

    def __init__(self, graph, config=None, alpha=None, beta=None, row_cache=4096):
        super().__init__(graph)
        self.config = config
        self.beta = beta if beta is not None else getattr(config, "DEFAULT_GRAVITY_BETA", 2.0)
        self.alpha = alpha if alpha is not None else getattr(config, "DEFAULT_GRAVITY_ALPHA", 1.0)
        self.distance_cutoff = getattr(config, "GRAVITY_DISTANCE_CUTOFF", None)
        degree = nx.degree_centrality(graph)  # :291
        self.size_variables = {n: 0.01 + degree.get(n, 0) for n in self.nodes}
        self.attraction_variables = {n: 0.01 + degree.get(n, 0) * self.alpha for n in self.nodes}
        self._attr = np.array([self.attraction_variables[n] for n in self.nodes], dtype=np.float64)
        self._euclid = len(self.nodes) > 2000  # :310
        if self._euclid:
            xy = []
            for n in self.nodes:  # :330-337
                if "pos" in graph.nodes[n]:
                    p = graph.nodes[n]["pos"]
                else:
                    h = hash(str(n))
                    p = (h % 1000, (h // 1000) % 1000)
                xy.append((float(p[0]), float(p[1])))
            self._xy = np.array(xy, dtype=np.float64)
        self._rows = OrderedDict()
        self._row_cache = int(row_cache)
        self._src_cum = None
        self._cache_valid = True

    # -- distances of one source to all nodes (inf = not connected), in self.nodes order --
    def _distance_row(self, source):
        i = self._node_pos[source]
        if self._euclid:  # :346-355
            dx = self._xy[i, 0] - self._xy[:, 0]
            dy = self._xy[i, 1] - self._xy[:, 1]
            d = np.maximum(_host_pow(dx * dx + dy * dy, 0.5), 0.1)
            d[i] = 0.1
            return d
        d = np.full(len(self.nodes), np.inf)
        for n, hops in nx.single_source_shortest_path_length(self.graph, source).items():  # = row of :318
            d[self._node_pos[n]] = hops
        return d

    def _row(self, source):
        row = self._rows.get(source)
        if row is not None:
            self._rows.move_to_end(source)
            return row
        d = self._distance_row(source)
        ok = np.isfinite(d) & (d > 0)  # :409-410
        if self.distance_cutoff:
            ok &= ~(d > self.distance_cutoff)  # :412
        ok[self._node_pos[source]] = False  # :403
        idx = np.nonzero(ok)[0]
        s_i = self.size_variables.get(source, 1.0)
        t_ij = (s_i * self._attr[idx]) / _host_pow(d[idx], self.beta)  # :416
        row = (idx, np.cumsum(t_ij))
        self._rows[source] = row
        if len(self._rows) > self._row_cache:
            self._rows.popitem(last=False)
        return row

    def get_source_target(self, agent_type=None, current_location=None, trip_count=0):
        if current_location is None:  # :381-383, 393-399
            if self._src_cum is None:
                self._src_cum = np.cumsum(np.array([self.size_variables[n] for n in self.nodes], dtype=np.float64))
            source = self.nodes[self._choices(self._src_cum)]
        else:
            source = current_location
        if source not in self._node_pos:  # :401-404
            return source, self._fallback(source)
        idx, cum = self._row(source)
        if len(idx) == 0:  # :421-424
            return source, self._fallback(source)
        return source, self.nodes[idx[self._choices(cum)]]

    def _fallback(self, source):
        target = self._any_other(source)
        return source if target is None else target

    @staticmethod
    def _choices(cum):
        """``random.choices(population, weights)[0]`` as an index (random.py: bisect over the running sums)."""
        total = cum[-1] + 0.0
        if total <= 0.0:
            raise ValueError("Total of weights must be greater than zero")
        if not np.isfinite(total):
            raise ValueError("Total of weights must be finite")
        # bisect.bisect(cum, x, 0, n - 1) on the running sums
        return int(np.searchsorted(cum[:len(cum) - 1], random.random() * total, side="right"))

    def set_parameters(self, beta=None, distance_cutoff=None):
        if beta is not None:
            self.beta = beta
        if distance_cutoff is not None:
            self.distance_cutoff = distance_cutoff
        self._rows.clear()

    @property
    def node_importance(self):
        return self.attraction_variables.copy()

    def get_model_info(self):
        return {"model_type": "voorhees_gravity", "beta": self.beta, "alpha": self.alpha,
                "distance_cutoff": self.distance_cutoff, "num_nodes": len(self.nodes),
                "avg_size": np.mean(list(self.size_variables.values())) if self.size_variables else 0,
                "avg_attraction": np.mean(list(self.attraction_variables.values())) if self.attraction_variables else 0,
                "cache_valid": self._cache_valid}


class ActivityBasedSelection(_Base):
    """lib/st_selection.py:650-1082 without the reference package (the GPU box does not have it) and without its
    O(V x pool) list-membership filters (:763-764, :797-799, :842-845, :886-888 become set look-ups).

    Set-up draws the five node pools with the reference's ``random.sample`` calls on the same lists in the same
    order (activity centres, business, work, home, distribution: :732-900), so the pools -- and every later
    (source, target) -- equal the reference's for the same seed.  A trip is one table row per agent type:
    (pool a first trip starts from, pool the target is drawn from, probability of drawing from that pool,
    whether a redraw keeps that probability) -- :926-1050."""

    _DEFAULT_TYPES = {"commuter": 0.4, "leisure": 0.3, "business": 0.2, "delivery": 0.1}

    def __init__(self, graph, config=None, agent_type_distributions=None):
        super().__init__(graph)
        if config is None:
            try:
                from config import MODELS as config  # inside the DRIFT application
            except Exception:
                config = None
        self.config = config
        cfg = lambda name, default: getattr(config, name, default) if config is not None else default  # noqa: E731
        self.agent_type_distributions = agent_type_distributions if agent_type_distributions is not None else \
            (cfg("DEFAULT_AGENT_TYPE_DISTRIBUTIONS", None) or dict(self._DEFAULT_TYPES))
        self.center_threshold = cfg("ACTIVITY_CENTER_THRESHOLD", 0.3)
        self.middle_threshold = cfg("ACTIVITY_MIDDLE_THRESHOLD", 0.7)
        self.activity_size_factor, self.activity_max_size = cfg("ACTIVITY_SIZE_FACTOR", 6), cfg("ACTIVITY_MAX_SIZE", 50)
        self.business_size_factor, self.business_max_size = cfg("BUSINESS_SIZE_FACTOR", 5), cfg("BUSINESS_MAX_SIZE", 80)
        self.work_size_factor, self.work_max_size = cfg("WORK_SIZE_FACTOR", 4), cfg("WORK_MAX_SIZE", 100)
        self.home_size_factor, self.home_max_size = cfg("HOME_SIZE_FACTOR", 4), cfg("HOME_MAX_SIZE", 100)
        self.distribution_size_factor = cfg("DISTRIBUTION_SIZE_FACTOR", 10)
        self.distribution_max_size = cfg("DISTRIBUTION_MAX_SIZE", 20)
        self.activity_center_ratio = cfg("ACTIVITY_CENTER_RATIO", 0.8)
        self.business_center_ratio = cfg("BUSINESS_CENTER_RATIO", 0.7)
        self.work_ratios = cfg("WORK_RATIOS", (0.2, 0.4, 0.4))
        self.home_ratios = cfg("HOME_RATIOS", (0.05, 0.35, 0.6))
        self.distribution_ratios = cfg("DISTRIBUTION_RATIOS", (0.5, 0.5))
        self._distribute_nodes_spatially(graph)

    # ---- pools (:690-900) ----------------------------------------------------------------------
    def _distribute_nodes_spatially(self, graph):
        pos = [graph.nodes[n]["pos"] for n in self.nodes]
        xs, ys = [p[0] for p in pos], [p[1] for p in pos]
        cx, cy = (min(xs) + max(xs)) / 2, (min(ys) + max(ys)) / 2
        dist = [((p[0] - cx) ** 2 + (p[1] - cy) ** 2) ** 0.5 for p in pos]  # the reference's float expression
        far = max([0] + dist)
        rel = [d / far for d in dist]  # ZeroDivisionError on a one-point layout, like the reference (:714)
        rings = ([], [], [])  # centre, middle, border -- in self.nodes order
        for n, r in zip(self.nodes, rel):
            rings[0 if r <= self.center_threshold else (1 if r <= self.middle_threshold else 2)].append(n)
        taken = set()

        def pool(size, ring_ids, counts, filtered, fallback):
            out = []
            avail = [[n for n in rings[r] if n not in taken] if f else rings[r] for r, f in zip(ring_ids, filtered)]
            for a, c in zip(avail, counts):
                c = min(c, len(a))
                if c > 0:
                    out.extend(random.sample(a, c))
            if not out:
                rest = [n for n in self.nodes if n not in taken] if taken else []
                out = random.sample(rest, min(fallback, len(rest))) if rest else \
                    random.sample(self.nodes, min(fallback, len(self.nodes)))
            return out

        V = len(self.nodes)
        centre, middle, border = rings
        # activity centres (:732-755): counts are capped by the UNFILTERED ring sizes
        size = min(V // self.activity_size_factor, self.activity_max_size)
        c0 = min(int(size * self.activity_center_ratio), len(centre))
        c1 = min(size - c0, len(middle))
        self.activity_centers = pool(size, (0, 1), (c0, c1), (False, False), 3)
        taken.update(self.activity_centers)
        # business (:757-786): same capping first, then against the filtered rings
        size = min(V // self.business_size_factor, self.business_max_size)
        c0 = min(int(size * self.business_center_ratio), len(centre))
        c1 = min(size - c0, len(middle))
        self.business_nodes = pool(size, (0, 1), (c0, c1), (True, True), 5)
        taken.update(self.business_nodes)
        # work (:788-828): the border ring is NOT filtered (border_nodes.copy())
        size = min(V // self.work_size_factor, self.work_max_size)
        c0, c1 = int(size * self.work_ratios[0]), int(size * self.work_ratios[1])
        self.work_nodes = pool(size, (0, 1, 2), (c0, c1, size - c0 - c1), (True, True, False), 5)
        taken.update(self.work_nodes)
        # home (:830-870)
        size = min(V // self.home_size_factor, self.home_max_size)
        c0, c1 = int(size * self.home_ratios[0]), int(size * self.home_ratios[1])
        self.home_nodes = pool(size, (0, 1, 2), (c0, c1, size - c0 - c1), (True, True, True), 5)
        taken.update(self.home_nodes)
        # distribution (:872-900)
        size = min(V // self.distribution_size_factor, self.distribution_max_size)
        c1 = int(size * self.distribution_ratios[0])
        self.distribution_nodes = pool(size, (1, 2), (c1, size - c1), (True, True), 3)

    # ---- trips (:903-1068) ---------------------------------------------------------------------
    def get_source_target(self, agent_type="random", current_location=None, trip_count=0):
        if agent_type == "commuter":  # :926-954: alternates work / home, never redraws
            if current_location is None:
                source = random.choice(self.home_nodes or self.nodes)
                return source, random.choice(self.work_nodes or self.nodes)
            pool = self.work_nodes if trip_count % 2 == 0 else self.home_nodes
            return current_location, random.choice(pool or self.nodes)
        first, pool, p, keep = {"delivery": (self.distribution_nodes, self.distribution_nodes, 0.7, False),
                                "leisure": (self.home_nodes, self.activity_centers, 0.6, True),
                                "business": (self.business_nodes, self.business_nodes, 0.8, True)}.get(
            agent_type, (None, None, None, False))
        source = random.choice(first or self.nodes) if current_location is None else current_location

        def target():
            if p is not None and random.random() < p and pool:
                return random.choice(pool)
            return random.choice(self.nodes)

        t = target()
        attempts = 0
        while t == source and attempts < 100:  # a redraw of the delivery / fallback pattern is uniform (:981, :1060)
            attempts += 1
            t = target() if keep else random.choice(self.nodes)
            if len(self.nodes) <= 1:
                break
        return source, t

    def get_agent_types_distribution(self):
        return self.agent_type_distributions

    def get_activity_nodes(self):
        return {"activity_centers": self.activity_centers, "business_nodes": self.business_nodes,
                "work_nodes": self.work_nodes, "home_nodes": self.home_nodes,
                "distribution_nodes": self.distribution_nodes}


FAST_MODELS = {"gravity": GravitySelection, "hub": HubAndSpokeSelection, "zone": ZoneBasedSelection,
               "activity": ActivityBasedSelection}


def make_fast_selector(mode, graph, settings=None):
    """Counterpart of ``drift_b200.st_selection.make_selector`` for the models restated here (same
    arguments as lib/simulation_thread.py:201-261 passes); None for the models that are not."""
    s = settings or {}
    if mode == "gravity":
        return GravitySelection(graph, alpha=s.get("gravity_alpha"), beta=s.get("gravity_beta"))
    if mode == "hub":
        return HubAndSpokeSelection(graph, hub_trip_probability=s.get("hub_trip_probability"),
                                    hub_percentage=s.get("hub_percentage"))
    if mode == "zone":
        return ZoneBasedSelection(graph, intra_zone_probability=s.get("zone_intra_probability"))
    if mode == "activity":
        return ActivityBasedSelection(graph, agent_type_distributions=s.get("agent_type_distributions"))
    return None

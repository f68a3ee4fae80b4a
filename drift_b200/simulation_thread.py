"""Drop-in replacement for the reference's ``SimulationThread`` (lib/simulation_thread.py:19-582).

Same constructor, attributes, methods, signals and payloads; the per-tick work that the reference
does in a Python loop over ``Agent`` objects (``update_simulation_step`` ->
``_update_agent_batch`` -> ``Agent.update`` -> ``TrafficManager``) runs on the B200 through the C
ABI (``drift_b200.engine.Engine``).  Install it by rebinding the name the controller imported:

    import lib.simulation_controller as sc
    from drift_b200.simulation_thread import SimulationThread
    sc.SimulationThread = SimulationThread

What stays on the host, on purpose (north_star: "OD pairs ... generated on host by the reference's
own s-t model with the same seed"):
  * the Bernoulli departure draw -- one ``random.random()`` per waiting agent per tick in agent
    order (lib/agent.py:180), interleaved with the selector's own draws, so that a run seeded like
    the reference consumes the identical MT19937 stream;
  * the s-t plugin call (lib/agent.py:47-68) and the trip records (lib/data_logger.py).
Everything downstream of a departure -- the route (NetworkX-exact bidirectional Dijkstra on the
device), link entry order, BPR speeds, crossing and arrival ticks -- is computed by the CUDA path.
Results are bit-identical to the reference (tests/test_gpu_thread.py).
"""
from __future__ import annotations

import ast
import datetime
import random
import time
from collections import deque

import numpy as np

from .data_logger import SimulationDataLogger
from .engine import Engine, ROUTE_NOPATH, ROUTE_OK, ROUTE_TRIVIAL, ROUTER_NX_EXACT
from .graph import LoweredGraph, lower_graph
from .qt_compat import QMutex, QMutexLocker, QThread, pyqtSignal
from .st_selection import make_selector
from .traffic_manager import TrafficManager

# reference constants (config.py:154-192, 204, 336-340)
SECONDS_PER_HOUR = 3600
SIMULATION_START_OFFSET = 21600
SIMULATION_DT = 0.1
DEFAULT_TIME_ACCELERATION = 10
UPDATE_FREQUENCY = 1
BATCH_SIZE = 50
STATS_UPDATE_STEP_FREQUENCY = 100
AGENT_UPDATE_BUFFER_MAXLEN = 1000
DEFAULT_HOURLY_PROBABILITIES = {
    0: 0.05, 1: 0.02, 2: 0.01, 3: 0.01, 4: 0.02, 5: 0.05,
    6: 0.15, 7: 0.25, 8: 0.35, 9: 0.20, 10: 0.15, 11: 0.18,
    12: 0.22, 13: 0.18, 14: 0.16, 15: 0.20, 16: 0.25, 17: 0.30,
    18: 0.35, 19: 0.25, 20: 0.15, 21: 0.12, 22: 0.08, 23: 0.06,
}


class AgentView:
    """Host-side face of one agent: the attributes the reference's ``Agent`` exposes
    (lib/agent.py:15-44) that UI managers, the data logger and selectors read.  Kinematic state
    lives on the device and is fetched on demand."""

    _next_id = 1

    def __init__(self, thread, index, graph, agent_type, st_selector, traffic_manager, hourly_probabilities,
                 initial_node):
        self.id = AgentView._next_id
        AgentView._next_id += 1
        self._thread = thread
        self.index = index
        self.graph = graph
        self.agent_type = agent_type
        self.st_selector = st_selector
        self.traffic_manager = traffic_manager
        self.hourly_probabilities = hourly_probabilities
        self.trip_count = 0
        self.state = "waiting"
        self.waiting = False
        self.wait_time = 0
        self.source = initial_node
        self.target = initial_node
        self.current_trip_data = None
        self._departures = 0  # routed departures so far (decides whether 'speed' exists yet)
        self._pidx_base = 0   # path_index of device route index 0 (changes when the agent is re-planned)

    # -- device-backed attributes -------------------------------------------------------------
    def _dev(self, name):
        return self._thread._snapshot()[name][self.index]

    @property
    def path_index(self):
        if not hasattr(self, "path") or self.path is None:
            return 0
        return self._pidx_base + int(self._dev("route_idx"))

    @property
    def speed(self):
        return float(self._dev("speed"))

    @property
    def distance_on_edge(self):
        return float(self._dev("dist"))

    @property
    def edge_length(self):
        return float(self._dev("elen"))

    @property
    def current_edge_key(self):
        l = int(self._dev("cur_link"))
        return None if l < 0 else self._thread.lowered.link_tuple(l)

    @property
    def position(self):
        return self._thread._agent_position(self.index)


class _AgentSnapshot:
    """Lightweight per-agent record of the ``agents_updated`` payload (lib/simulation_thread.py:394-416)."""

    __slots__ = ("id", "position", "state", "agent_type", "trip_count", "speed", "source", "target", "path",
                 "path_index", "current_node", "target_node", "progress", "simulation_time")


class SimulationThread(QThread):
    log_message = pyqtSignal(str)
    agents_updated = pyqtSignal(list)
    simulation_finished = pyqtSignal(dict)
    trip_completed = pyqtSignal(dict)
    status_updated = pyqtSignal(dict)
    st_selector_ready = pyqtSignal(object, str)

    def __init__(self, graph, num_agents, selection_mode, duration_hours=24, device=0, lowered=None):
        super().__init__()
        self.graph = graph
        self.num_agents = num_agents
        self.selection_mode = selection_mode
        self.duration_hours = duration_hours
        self.running = False
        self.paused = False
        self.agents = []
        self.traffic_manager = None
        self.data_logger = None
        self.simulation_time = 0.0
        self.dt = SIMULATION_DT
        self.time_acceleration = DEFAULT_TIME_ACCELERATION
        self.max_simulation_time = duration_hours * SECONDS_PER_HOUR
        self.agent_trip_states = {}
        self.last_report_time = 0
        self.step = 0
        self.agent_update_buffer = deque(maxlen=AGENT_UPDATE_BUFFER_MAXLEN)
        self.update_mutex = QMutex()
        self.batch_size = BATCH_SIZE
        self.update_frequency = UPDATE_FREQUENCY
        self.snapshot_stride = 1  # agents_updated carries every k-th agent (1 = all, the reference's behaviour)
        self.last_visual_update = 0
        self.agent_positions = []
        self.agent_states = []
        self.settings = None
        self.st_selector = None
        self.fast_selectors = False  # True: vectorised gravity / hub / zone models (same sequences, fast_selection.py)
        self._cached_agent_type_percentages = {}
        self._cached_network_stats = {}
        self._last_network_stats_step = 0
        # engine side
        self.device = device
        self.lowered: LoweredGraph | None = lowered
        self.engine: Engine | None = None
        self.guard_reference_crashes = True  # see finish_simulation
        # Extension, off by default (the reference never updates travel_time): every `reroute_period`
        # simulated seconds link travel times become t0 / BPR factor(volume), en-route agents are
        # re-planned from the head of their current link and later departures route on those times.
        self.reroute_period = None
        self._last_reroute = 0.0
        self._congested = False
        self.reroutes = 0
        self._waiting = []
        self._open = None
        self._snap = None
        self._snap_step = -1
        self._hourly = DEFAULT_HOURLY_PROBABILITIES
        self._moving_count = 0

    # ---- settings (lib/simulation_thread.py:71-138) -----------------------------------------
    def apply_settings(self, settings):
        self.settings = settings
        if self.traffic_manager is not None:
            self.traffic_manager.set_bpr_parameters(settings["bpr_alpha"], settings["bpr_beta"])

    def update_runtime_settings(self, settings):
        self.settings = settings
        if self.traffic_manager is not None:
            self.traffic_manager.set_bpr_parameters(settings["bpr_alpha"], settings["bpr_beta"])
        if self.agents and "hourly_probabilities" in settings:
            for agent in self.agents:
                agent.hourly_probabilities = settings["hourly_probabilities"]
            self._hourly = settings["hourly_probabilities"]
            self.log_message.emit("Updated hourly probabilities for %d agents" % len(self.agents))
        if self.st_selector:
            self.update_st_selector_parameters(settings)
            self.log_message.emit("Runtime settings updated successfully")

    def update_st_selector_parameters(self, settings):
        try:
            sel, mode = self.st_selector, self.selection_mode
            if mode == "zone" and hasattr(sel, "set_parameters"):
                sel.set_parameters(intra_zone_probability=settings["zone_intra_probability"])
                self.log_message.emit("Zone model updated: intra-zone probability = %.2f" % settings["zone_intra_probability"])
            elif mode == "gravity" and hasattr(sel, "set_parameters"):
                # the reference passes alpha= to a method that only takes beta/distance_cutoff
                # (lib/simulation_thread.py:114-118); the resulting TypeError is swallowed below
                sel.set_parameters(alpha=settings["gravity_alpha"], beta=settings["gravity_beta"])
                self.log_message.emit("Gravity model updated")
            elif mode == "hub" and hasattr(sel, "set_parameters"):
                sel.set_parameters(hub_trip_probability=settings["hub_trip_probability"],
                                   hub_percentage=settings["hub_percentage"])
                self.log_message.emit("Hub model updated")
            elif mode == "activity":
                self.log_message.emit("Activity model: parameters will apply to new agents in next simulation")
            else:
                self.log_message.emit("Random selection: no parameters to update")
        except Exception as exc:
            self.log_message.emit("Error updating ST selector parameters: %s" % exc)

    # ---- initialisation (lib/simulation_thread.py:140-283) ------------------------------------
    def initialize_simulation(self):
        try:
            if self.lowered is None:
                self.lowered = lower_graph(self.graph)
            lg = self.lowered
            alpha, beta = 0.15, 4.0
            if self.settings:
                alpha, beta = self.settings["bpr_alpha"], self.settings["bpr_beta"]
            self.engine = Engine(lg, self.num_agents, alpha=alpha, beta=beta, device=self.device)
            self.traffic_manager = TrafficManager(self.graph, lg, self.engine, thread=self)
            self.traffic_manager.bpr_alpha, self.traffic_manager.bpr_beta = alpha, beta
            self.data_logger = SimulationDataLogger(output_dir="data",
                                                    simulation_start_time=datetime.datetime(2025, 7, 10, 6, 0, 0))
            self.log_message.emit("Initializing %d agents with %s selection..." % (self.num_agents, self.selection_mode))
            hourly = None
            if self.settings and "hourly_probabilities" in self.settings:
                hourly = self.settings["hourly_probabilities"]
            self._hourly = hourly if hourly is not None else DEFAULT_HOURLY_PROBABILITIES
            precomputed = self.st_selector is not None
            if precomputed:
                self.log_message.emit("Using pre-computed ST selector model")
                selector = self.st_selector
            else:
                selector = make_selector(self.selection_mode, self.graph, self.settings,
                                         fast=getattr(self, "fast_selectors", False))
            nodes = lg.nodes
            self.agents = []
            if self.selection_mode == "activity":
                if self.settings and "agent_type_distributions" in self.settings:
                    dist = self.settings["agent_type_distributions"]
                else:
                    dist = selector.get_agent_types_distribution()
                types, weights = list(dist.keys()), list(dist.values())
                for i in range(self.num_agents):
                    agent_type = random.choices(types, weights=weights)[0]  # :183 / :215
                    self.agents.append(self._make_agent(i, agent_type, selector, hourly, random.choice(nodes)))
            else:
                kind = self.selection_mode if self.selection_mode in ("gravity", "zone", "hub") else "random"
                for i in range(self.num_agents):
                    self.agents.append(self._make_agent(i, kind, selector, hourly, random.choice(nodes)))  # agent.py:40
            self.st_selector = selector
            self.st_selector_ready.emit(selector, self.selection_mode)
            self._open = [False] * self.num_agents
            for agent in self.agents:  # :274-276 -- every agent starts with an open trip (first-trip quirk)
                self.data_logger.start_trip(agent, self.simulation_time)
                self.agent_trip_states[id(agent)] = True
                self._open[agent.index] = True
            self._waiting = list(range(self.num_agents))
            self._moving_count = 0
            self.log_message.emit("Starting %s-hour simulation..." % self.duration_hours)
            return True
        except Exception as exc:
            self.log_message.emit("Simulation initialization error: %s" % exc)
            self._init_error = exc
            return False

    def _make_agent(self, index, agent_type, selector, hourly, initial_node):
        return AgentView(self, index, self.graph, agent_type, selector, self.traffic_manager,
                         hourly if hourly is not None else DEFAULT_HOURLY_PROBABILITIES, initial_node)

    # ---- one tick (lib/simulation_thread.py:288-390) ------------------------------------------
    def update_simulation_step(self):
        if not self.running or self.simulation_time >= self.max_simulation_time:
            return False
        delta = self.dt * self.time_acceleration
        self.simulation_time += delta
        self.step += 1
        t = self.simulation_time
        lg, eng, agents = self.lowered, self.engine, self.agents
        # Agent.update for waiting agents (lib/agent.py:172-183): one MT19937 draw each, in agent order
        hour = int(((t + SIMULATION_START_OFFSET) / SECONDS_PER_HOUR) % 24)
        threshold = (self._hourly.get(hour, 0.05) * 4.0) * delta / SECONDS_PER_HOUR
        departing = self._draw_departures(threshold)
        started = []
        if departing:
            index = lg.node_index
            src = [index.get(agents[i].source, -1) for i in departing]
            dst = [index.get(agents[i].target, -1) for i in departing]
            status, n_links, links = eng.route(src, dst, router=ROUTER_NX_EXACT, congested=self._congested)
            nodes = lg.nodes
            for i, s, st, ln in zip(departing, src, status, links):
                a = agents[i]
                if st == ROUTE_OK:
                    a.path = [nodes[k] for k in lg.links_to_nodes(s, ln)]
                    a._pidx_base = 0
                    a.state = "moving"
                    a._departures += 1
                    started.append(i)
                elif st == ROUTE_TRIVIAL:  # path == [s]: arrival branch at once (lib/agent.py:107-116)
                    a.path = [a.source]
                    a.wait_time = 0
                else:  # NetworkXNoPath / NodeNotFound (lib/agent.py:87-99)
                    a.path = None
            eng.commit_departures(departing)
        eng.step(1, delta)
        arr_agent, _ = eng.drain_arrivals()
        arrived = sorted(arr_agent.tolist())
        self._snap = None
        # trip bookkeeping in agent order (lib/simulation_thread.py:326-355)
        if started or arrived:
            moved = set(started)
            for i in sorted(moved.union(arrived)):
                a = agents[i]
                if i in moved:
                    if not self._open[i]:
                        self.data_logger.start_trip(a, t)
                        self._open[i] = True
                        self.agent_trip_states[id(a)] = True
                else:
                    a.state = "waiting"
                    a.wait_time = 0
                    if self._open[i]:
                        self._end_trip(a, t, len(a.path) - 2)
                        self._queue_trip_completion(a)
            if started:
                gone = set(started)
                self._waiting = [i for i in self._waiting if i not in gone]
            if arrived:
                self._waiting = sorted(self._waiting + arrived)
            self._moving_count += len(started) - len(arrived)
        if self.reroute_period and t - self._last_reroute >= self.reroute_period:
            self._reroute(t)
        if self.step % self.update_frequency == 0:
            with QMutexLocker(self.update_mutex):
                data = self._prepare_agent_data()
                for rec in data:
                    rec.simulation_time = self.simulation_time
                self.agents_updated.emit(data)
                self.last_visual_update = self.step
        if self.step % STATS_UPDATE_STEP_FREQUENCY == 0:
            self._emit_status_update()
        if self.simulation_time - self.last_report_time >= 1800:
            self._emit_progress_report()
        return True

    def _draw_departures(self, threshold):
        """Agent.update for waiting agents (lib/agent.py:172-183): one ``random.random()`` per waiting agent in agent
        order, a hit starts a journey (whose selector draws from the same generator before the next agent's draw).
        The reference loops in Python; here the draws between two hits are produced in bulk: the MT19937 state of
        ``random`` is handed to ``numpy.random.MT19937`` (same generator, same tempering), uniforms are built like
        ``genrand_res53`` -- (a >> 5) * 2**26 + (b >> 6) over 2**53 -- the first hit is located, the generator is
        advanced to just after that draw and handed back, the journey is started, and so on.  Same hits, same
        (source, target) answers, same generator state afterwards (tests/test_host_cpu.py)."""
        waiting, agents = self._waiting, self.agents
        n = len(waiting)
        if n < 64:  # not worth two state transfers
            draw = random.random
            departing = []
            for i in waiting:
                if draw() < threshold:
                    departing.append(i)
                    self._start_new_journey(agents[i])
            return departing
        departing = []
        pos = 0
        bg = np.random.MT19937()
        chunk = 1 << 16
        while pos < n:
            version, internal, gauss = random.getstate()
            key = np.array(internal[:-1], dtype=np.uint32)
            start = {"bit_generator": "MT19937", "state": {"key": key, "pos": int(internal[-1])}}
            bg.state = start
            m = min(n - pos, chunk)
            raw = bg.random_raw(2 * m)
            u = ((raw[0::2] >> np.uint64(5)).astype(np.float64) * 67108864.0 +
                 (raw[1::2] >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)
            hits = np.nonzero(u < threshold)[0]
            used = m if len(hits) == 0 else int(hits[0]) + 1
            if used != m:  # rewind to just after the hit's draw
                bg.state = start
                bg.random_raw(2 * used)
            st = bg.state["state"]
            random.setstate((version, tuple(int(x) for x in st["key"]) + (int(st["pos"]),), gauss))
            if len(hits):
                i = waiting[pos + used - 1]
                departing.append(i)
                self._start_new_journey(agents[i])
            pos += used
        return departing

    def _reroute(self, t):
        """Extension step: congested travel times, re-plan en-route agents, keep the host paths in sync.
        The device restarts a re-planned route at index 0 (route = current link + new path), the host
        keeps the reference's convention path = path[:path_index + 1] + new_path."""
        eng, lg = self.engine, self.lowered
        before = eng.agents(fields=("route_idx",))["route_idx"]
        eng.update_travel_times()
        self._congested = True
        self.reroutes += eng.reroute_moving(ROUTER_NX_EXACT)
        self._last_reroute = t
        self._snap = None
        after = self._snapshot()["route_idx"]
        nodes = lg.nodes
        for a in self.agents:
            if a.state != "moving" or int(after[a.index]) != 0:
                continue
            pidx = a._pidx_base + int(before[a.index])
            links = eng.agent_route(a.index)
            a.path = a.path[:pidx] + [nodes[k] for k in lg.links_to_nodes(lg.link_u[links[0]], links)]
            a._pidx_base = pidx

    def _start_new_journey(self, a):
        """Agent.start_new_journey up to the routing call (lib/agent.py:46-91)."""
        a.state = "moving"  # lib/agent.py:181; reverted below when no route exists
        sel = a.st_selector
        current_location = a.target
        a.source, a.target = sel.get_source_target(a.agent_type, current_location, a.trip_count)
        if hasattr(sel, "last_trip_type") and sel.__class__.__name__ == "HubAndSpokeSelection":
            a.agent_type = sel.last_trip_type
        elif hasattr(sel, "node_to_zone") and sel.__class__.__name__ == "ZoneBasedSelection":
            zs, zt = sel.node_to_zone.get(a.source), sel.node_to_zone.get(a.target)
            if zs is not None and zt is not None:
                a.agent_type = "intra_zone" if zs == zt else "inter_zone"
        a.trip_count += 1
        a.state = "waiting"  # becomes 'moving' once the device returned a route

    def _end_trip(self, a, t, upto):
        """Materialise ``route_taken`` (the reference appends path[path_index] every moving tick,
        lib/simulation_thread.py:384-390 + lib/data_logger.py:44-53) and close the record."""
        rec = a.current_trip_data
        path = getattr(a, "path", None)
        if rec is not None and path is not None:
            route = rec["route_taken"]
            seen = set()
            try:
                seen.update(route)
            except TypeError:
                seen = None
            for node in path[:max(upto, -1) + 1]:
                if not node:
                    continue  # falsy labels (int 0) are skipped by update_trip
                if seen is not None:
                    if node in seen:
                        continue
                    seen.add(node)
                elif node in route:
                    continue
                route.append(node)
        metres = None
        if path is not None and len(path) > 1:
            idx = self.lowered.node_index
            metres = self.lowered.path_length_m([idx[n] for n in path])
        if path is None and hasattr(a, "path") and self.guard_reference_crashes:
            # reference: TypeError in data_logger.py:68 (first departure failed to route); guarded here
            saved = a.path
            del a.path
            try:
                self.data_logger.end_trip(a, t, self.traffic_manager)
            finally:
                a.path = saved
        else:
            self.data_logger.end_trip(a, t, self.traffic_manager, path_length_m=metres)
        self._open[a.index] = False
        self.agent_trip_states[id(a)] = False

    def _queue_trip_completion(self, agent):
        if not self.data_logger.trips_data:
            return
        last = self.data_logger.trips_data[-1]
        route = last.get("route_taken", "[]")
        if isinstance(route, str):
            try:
                route = ast.literal_eval(route)
            except Exception:
                route = []
        self.trip_completed.emit({
            "trip_id": len(self.data_logger.trips_data),
            "agent_id": getattr(agent, "id", "N/A"),
            "agent_type": agent.agent_type,
            "start_node": last.get("origin_node", "Unknown"),
            "end_node": last.get("destination_node", "Unknown"),
            "start_time": last.get("start_time", "00:00:00"),
            "duration": last.get("trip_duration_min", 0) * 60,
            "distance": last.get("trip_distance_km", 0) * 1000,
            "avg_speed": last.get("average_speed_kmh", 0) / 3.6,
            "path_nodes": route,
        })

    # ---- device snapshots -------------------------------------------------------------------------
    def _snapshot(self):
        if self._snap is None or self._snap_step != self.step:
            self._snap = self.engine.agents(fields=("state", "dist", "speed", "elen", "cur_link", "route_idx"))
            self._snap_step = self.step
        return self._snap

    def _agent_position(self, i):
        lg = self.lowered
        a = self.agents[i]
        snap = self._snapshot()
        pos = lg.pos
        if pos is None:
            return np.zeros(2)
        link = int(snap["cur_link"][i])
        if link < 0:
            return np.array(pos[lg.node_index[a.target]], dtype=float)
        u, v = lg.link_u[link], lg.link_v[link]
        length = float(snap["elen"][i])
        direction = (pos[v] - pos[u]) / length if length > 0 else np.zeros(2)
        return pos[u] + direction * float(snap["dist"][i])

    def _prepare_agent_data(self):
        """lib/simulation_thread.py:394-416.  ``snapshot_stride`` (default 1 = every agent, like the reference) thins
        the snapshot for large runs: the untouched PyQt widget then draws every k-th agent."""
        snap = self._snapshot()
        speed, ridx = snap["speed"].tolist(), snap["route_idx"].tolist()  # one bulk conversion, not one per agent
        out = []
        for a in self.agents[::max(1, int(getattr(self, "snapshot_stride", 1)))]:
            rec = _AgentSnapshot()
            rec.id = a.id
            rec.position = self._agent_position(a.index)
            rec.state = a.state
            rec.agent_type = a.agent_type
            rec.trip_count = a.trip_count
            rec.speed = speed[a.index] if a._departures else 0
            rec.source, rec.target = a.source, a.target
            rec.path = getattr(a, "path", [])
            rec.path_index = a._pidx_base + ridx[a.index] if a._departures else 0
            rec.current_node = rec.target_node = None
            rec.progress = 0.0
            out.append(rec)
        return out

    # ---- status / progress (lib/simulation_thread.py:421-478) -----------------------------------
    def _emit_status_update(self):
        moving = self._moving_count
        pct = {}
        if self.step % (STATS_UPDATE_STEP_FREQUENCY * 10) == 0:
            counts = {}
            for a in self.agents:
                if a.state == "moving":
                    counts[a.agent_type] = counts.get(a.agent_type, 0) + 1
            if moving > 0:
                pct = {k: v / moving * 100 for k, v in counts.items()}
            self._cached_agent_type_percentages = pct
        else:
            pct = self._cached_agent_type_percentages
        if (self.step - self._last_network_stats_step) >= 100:
            self._cached_network_stats = self.traffic_manager.get_network_statistics()
            self._last_network_stats_step = self.step
        self.status_updated.emit({
            "simulation_time": self.simulation_time,
            "moving_agents": moving,
            "total_agents": len(self.agents),
            "network_utilization": self._cached_network_stats.get("network_utilization", 0.0),
            "active_agent_types": pct,
        })

    def _emit_progress_report(self):
        now = datetime.datetime(2025, 7, 10, 6, 0, 0) + datetime.timedelta(seconds=self.simulation_time)
        stats = self.traffic_manager.get_network_statistics()
        self.log_message.emit("Time: %s | Progress: %.1f/%sh | Trips: %d | On roads: %d | Utilization: %.1f%% | Speed: %sx" % (
            now.strftime("%H:%M"), self.simulation_time / 3600, self.duration_hours, len(self.data_logger.trips_data),
            stats["total_agents_on_roads"], 100 * stats["network_utilization"], self.time_acceleration))
        self.last_report_time = self.simulation_time

    # ---- finish (lib/simulation_thread.py:483-500) ------------------------------------------------
    def finish_simulation(self):
        snap = self._snapshot()
        for a in self.agents:
            if self._open[a.index]:
                upto = a._pidx_base + int(snap["route_idx"][a.index]) if a.state == "moving" else (
                    len(a.path) - 2 if getattr(a, "path", None) else -1)
                self._end_trip(a, self.simulation_time, upto)
        csv_file = self.data_logger.save_to_csv()
        stats = self.data_logger.get_statistics()
        stats["csv_file"] = csv_file
        self.log_message.emit("Simulation completed! Total trips: %s" % stats.get("total_trips", 0))
        self.log_message.emit("Data saved to: %s" % csv_file)
        self.running = False
        self.simulation_finished.emit(stats)

    # ---- controls ---------------------------------------------------------------------------------
    def set_performance_mode(self, mode):
        if mode == "ultra_smooth":
            self.update_frequency, self.batch_size = 1, 10
        elif mode == "balanced":
            self.update_frequency, self.batch_size = 1, 50
        elif mode == "performance":
            self.update_frequency, self.batch_size = 1, 75
        self.log_message.emit("Performance mode set to: %s" % mode)

    def set_speed_multiplier(self, multiplier):
        self.time_acceleration = multiplier
        self.log_message.emit("Speed multiplier set to: %sx" % multiplier)

    def get_performance_metrics(self):
        return {"update_frequency": self.update_frequency, "batch_size": self.batch_size,
                "agent_count": len(self.agents), "step": self.step, "simulation_time": self.simulation_time}

    def run(self):
        self.running = True
        if not self.initialize_simulation():
            return
        n = len(self.agents)
        self.set_performance_mode("ultra_smooth" if n < 200 else ("balanced" if n < 800 else "performance"))
        while self.running and self.simulation_time < self.max_simulation_time:
            t0 = time.time()
            if self.paused:
                self.msleep(5)
                continue
            self.update_simulation_step()
            pause = max(0, 0.033 - (time.time() - t0))  # the reference paces the GUI at ~30 ticks/s
            if pause > 0 and self.update_frequency < 10 ** 6:
                self.msleep(int(pause * 1000))
        if self.running:
            self.finish_simulation()

    def stop(self):
        self.running = False

    def pause(self):
        self.paused = True

    def resume(self):
        self.paused = False

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None

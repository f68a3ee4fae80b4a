"""TEST INFRASTRUCTURE ONLY -- live-reference harness (container only).

Runs the UNMODIFIED reference (``/root/reference``) headless and records what
its hot path does, so that ``oracle/make_golden.py`` can freeze the result as
small fixtures under ``tests/golden/``.  ``/root/reference`` does not exist on
the GPU box, so nothing outside ``oracle/make_golden.py`` and the
"reference present" CPU tests may import this module.

How the reference is driven:
  * ``lib/simulation_thread.py:1`` imports ``PyQt5.QtCore.{QThread, pyqtSignal,
    QMutex, QMutexLocker}``; PyQt5 is absent, so a tiny stand-in is injected in
    ``sys.modules`` and the file imports and runs unchanged.
  * one call of ``SimulationThread.update_simulation_step``
    (``lib/simulation_thread.py:288-324``) is one tick.
  * ``update_frequency = 10**9`` (the reference's own knob, ``:53,306``)
    switches the per-tick visualisation snapshot off.
"""
from __future__ import annotations

import os
import random
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("DRIFT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "simulation_thread.py"))


# --------------------------------------------------------------------------
# PyQt5 stand-in
# --------------------------------------------------------------------------
class _BoundSignal:
    def __init__(self):
        self._slots = []

    def connect(self, slot):
        self._slots.append(slot)

    def disconnect(self, slot=None):
        if slot is None:
            self._slots.clear()
        elif slot in self._slots:
            self._slots.remove(slot)

    def emit(self, *args):
        for s in list(self._slots):
            s(*args)


class _Signal:
    """Class-level ``pyqtSignal(...)``: yields one bound signal per instance."""

    def __init__(self, *types_):
        self._name = None

    def __set_name__(self, owner, name):
        self._name = "_sig_" + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        sig = obj.__dict__.get(self._name)
        if sig is None:
            sig = obj.__dict__[self._name] = _BoundSignal()
        return sig


class _QThread:
    def __init__(self, *a, **k):
        self._running = False

    def msleep(self, ms):
        pass

    def start(self):
        self._running = True
        try:
            self.run()
        finally:
            self._running = False

    def run(self):
        pass

    def wait(self, *a):
        return True

    def isRunning(self):
        return self._running

    def deleteLater(self):
        pass


class _QMutex:
    def lock(self):
        pass

    def unlock(self):
        pass


class _QMutexLocker:
    def __init__(self, m):
        self.m = m

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install_pyqt_shim():
    if "PyQt5.QtCore" in sys.modules:
        return
    pkg = types.ModuleType("PyQt5")
    core = types.ModuleType("PyQt5.QtCore")
    core.QThread = _QThread
    core.pyqtSignal = _Signal
    core.QMutex = _QMutex
    core.QMutexLocker = _QMutexLocker
    pkg.QtCore = core
    sys.modules["PyQt5"] = pkg
    sys.modules["PyQt5.QtCore"] = core


_ref_mods = None


def import_reference():
    """Import the reference's hot-path modules unmodified; returns a namespace."""
    global _ref_mods
    if _ref_mods is not None:
        return _ref_mods
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    install_pyqt_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    ns.config = importlib.import_module("config")
    ns.agent = importlib.import_module("lib.agent")
    ns.traffic_manager = importlib.import_module("lib.managers.traffic_manager")
    ns.data_logger = importlib.import_module("lib.data_logger")
    ns.st_selection = importlib.import_module("lib.st_selection")
    ns.graph_loader = importlib.import_module("lib.graph_loader")
    ns.simulation_thread = importlib.import_module("lib.simulation_thread")
    _ref_mods = ns
    return ns


class _quiet:
    """Silence the reference's prints (tqdm, DEBUG lines, loader chatter)."""

    def __enter__(self):
        self._o, self._e = sys.stdout, sys.stderr
        self._f = open(os.devnull, "w")
        sys.stdout = sys.stderr = self._f

    def __exit__(self, *a):
        sys.stdout, sys.stderr = self._o, self._e
        self._f.close()
        return False


def load_graph(path, seed=0):
    """``GraphLoader.load_graph`` (``lib/graph_loader.py:47-134``), seeded."""
    import numpy as np

    ref = import_reference()
    random.seed(seed)
    np.random.seed(seed)
    with _quiet():
        G = ref.graph_loader.GraphLoader().load_graph(path)
    return G


def ensure_edge_weights(G):
    """``GraphLoader._ensure_edge_weights`` (``lib/graph_loader.py:708-776``)."""
    ref = import_reference()
    with _quiet():
        ref.graph_loader.GraphLoader()._ensure_edge_weights(G)
    return G


def run_reference(G, num_agents, mode="random", seed=0, accel=10, hours=1,
                  settings=None, max_ticks=None, snapshot_every=50,
                  st_selector=None, finish=True, record_occupancy=True):
    """Run the reference headless and record its hot path.

    Returns a dict of plain-Python / numpy data:
      nodes, links            graph order (``list(G.nodes)``, ``G.edges(keys=True)``)
      init_nodes              per-agent initial node index
      agent_types0            per-agent agent_type at init
      departures              list of (tick, agent_idx, src_idx, dst_idx, path|None)
      entries                 list of (tick, agent_idx, link_idx, speed, elen)
      arrivals                list of (tick, agent_idx)
      occ                     int32[ticks, L] occupancy after each tick
      snaps                   {tick: (state u8[N], d f64[N], speed f64[N], path_index i32[N])}
      times                   simulation_time after each tick (fp64)
      trips                   final ``data_logger.trips_data`` with agent index
      stats                   ``get_statistics()`` after finish
    """
    import numpy as np

    ref = import_reference()
    Agent = ref.agent.Agent
    SimulationThread = ref.simulation_thread.SimulationThread

    nodes = list(G.nodes)
    node_idx = {n: i for i, n in enumerate(nodes)}
    links = list(G.edges(keys=True))
    link_idx = {e: i for i, e in enumerate(links)}

    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="drift_ref_")
    os.chdir(tmp)  # the logger does os.makedirs('data') in cwd (data_logger.py:15)
    rec = dict(departures=[], entries=[], arrivals=[])
    orig_journey = Agent.start_new_journey
    orig_setup = Agent.setup_current_edge
    cur = dict(tick=0)
    agent_index = {}

    def journey(self):
        orig_journey(self)
        p = getattr(self, "path", None)
        rec["departures"].append(
            (cur["tick"], agent_index[id(self)], node_idx.get(self.source, -1),
             node_idx.get(self.target, -1),
             None if p is None else [node_idx[x] for x in p]))

    def setup(self):
        orig_setup(self)
        after = self.current_edge_key
        if after is not None:  # non-arrival branch (agent.py:118-160): a link entry
            rec["entries"].append((cur["tick"], agent_index[id(self)], link_idx[after],
                                   float(self.speed), float(self.edge_length)))

    try:
        Agent.start_new_journey = journey
        Agent.setup_current_edge = setup
        random.seed(seed)
        np.random.seed(seed)
        Agent._next_id = 1
        with _quiet():
            st = SimulationThread(G, num_agents, mode, duration_hours=hours)
            st.time_acceleration = accel
            if settings is not None:
                st.apply_settings(settings)
            if st_selector is not None:
                st.st_selector = st_selector
            st.running = True
            ok = st.initialize_simulation()
        if not ok:
            raise RuntimeError("reference initialize_simulation failed")
        st.update_frequency = 10 ** 9
        for i, a in enumerate(st.agents):
            agent_index[id(a)] = i
        N = len(st.agents)
        L = len(links)
        init_nodes = np.array([node_idx[a.source] for a in st.agents], dtype=np.int32)
        agent_types0 = [a.agent_type for a in st.agents]
        occ_rows, times, snaps = [], [], {}
        tm = st.traffic_manager
        prev_state = [a.state for a in st.agents]
        with _quiet():
            while True:
                cur["tick"] = st.step + 1
                if not st.update_simulation_step():
                    break
                t = st.step
                times.append(st.simulation_time)
                if record_occupancy:
                    row = np.zeros(L, dtype=np.int32)
                    for e, lst in tm.edge_agents.items():
                        if lst:
                            row[link_idx[e]] = len(lst)
                    occ_rows.append(row)
                for i, a in enumerate(st.agents):
                    s = a.state
                    if prev_state[i] == "moving" and s == "waiting":
                        rec["arrivals"].append((t, i))
                    prev_state[i] = s
                if snapshot_every and t % snapshot_every == 0:
                    snaps[t] = (
                        np.array([1 if a.state == "moving" else 0 for a in st.agents], dtype=np.uint8),
                        np.array([float(getattr(a, "distance_on_edge", 0.0)) for a in st.agents]),
                        np.array([float(getattr(a, "speed", 0.0)) for a in st.agents]),
                        np.array([int(getattr(a, "path_index", 0)) for a in st.agents], dtype=np.int32),
                    )
                if max_ticks is not None and st.step >= max_ticks:
                    break
            trips_before_finish = len(st.data_logger.trips_data)
            stats = None
            finish_error = None
            if finish:
                st.running = True
                try:
                    st.finish_simulation()
                    stats = st.data_logger.get_statistics()
                except TypeError as e:  # open trip with path None
                    finish_error = repr(e)
        # agent_id in the records is 'A' + id(obj): translate to the agent index
        id2idx = {"A%06d" % id(a): i for i, a in enumerate(st.agents)}
        trips = []
        for row in st.data_logger.trips_data:
            r = dict(row)
            r["agent_index"] = id2idx[r.pop("agent_id")]
            trips.append(r)
        out = dict(
            nodes=nodes, links=links, init_nodes=init_nodes, agent_types0=agent_types0,
            departures=rec["departures"], entries=rec["entries"], arrivals=rec["arrivals"],
            occ=np.array(occ_rows, dtype=np.int32) if record_occupancy else None,
            snaps=snaps, times=np.array(times), trips=trips, stats=stats,
            trips_before_finish=trips_before_finish, finish_error=finish_error,
            final_state=np.array([1 if a.state == "moving" else 0 for a in st.agents], dtype=np.uint8),
            trip_counts=np.array([a.trip_count for a in st.agents], dtype=np.int32),
            capacities=np.array([tm.edge_capacities[e] for e in links], dtype=np.int32),
            ticks=st.step, sim=st,
        )
        return out
    finally:
        Agent.start_new_journey = orig_journey
        Agent.setup_current_edge = orig_setup
        os.chdir(cwd)

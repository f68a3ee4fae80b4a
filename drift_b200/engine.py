"""Python handle on one device engine (thin wrapper over the C ABI, numpy in / numpy out)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from . import bpr as _bpr
from .graph import LoweredGraph

ROUTER_NX_EXACT = _capi.ROUTER_NX_EXACT
ROUTER_BATCHED = _capi.ROUTER_BATCHED
ROUTE_OK, ROUTE_TRIVIAL, ROUTE_NOPATH = _capi.ROUTE_OK, _capi.ROUTE_TRIVIAL, _capi.ROUTE_NOPATH


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """One engine = one CUDA device = one rank.

    ``lg``           lowered graph (``drift_b200.graph``)
    ``num_agents``   local agents; ``agent_base`` = global id of local agent 0
    ``alpha/beta``   BPR parameters (lib/managers/traffic_manager.py:25-31)
    """

    def __init__(self, lg: LoweredGraph, num_agents: int, alpha: float = 0.15, beta: float = 4.0, device: int = 0,
                 agent_base: int = 0, route_arena_links: int = 0, total_agents: int = 0):
        self._lib = _capi.load()
        self.lg = lg
        self.N = int(num_agents)
        self.agent_base = int(agent_base)
        # agents of the whole (possibly sharded) run: no link ever holds more -> bound of the BPR factor tables
        self.total_agents = max(int(total_agents), self.agent_base + self.N)
        self._h = _capi.ENGINE()
        self._keep = dict(
            fwd_rowptr=_i32(lg.fwd_rowptr), fwd_col=_i32(lg.fwd_col), fwd_w=_f64(lg.fwd_w), fwd_link=_i32(lg.fwd_link),
            bwd_rowptr=_i32(lg.bwd_rowptr), bwd_col=_i32(lg.bwd_col), bwd_w=_f64(lg.bwd_w), bwd_link=_i32(lg.bwd_link),
            link_len=_f64(lg.link_len), link_v0=_f64(lg.link_v0), link_t0=_f64(lg.link_t0), link_cap=_i32(lg.link_cap),
            link_u=_i32(lg.link_u), link_v=_i32(lg.link_v))
        k = self._keep
        d = _capi.GraphDesc(
            V=lg.V, L=lg.L, Ef=len(k["fwd_col"]), Eb=len(k["bwd_col"]),
            fwd_rowptr=_p(k["fwd_rowptr"], C.c_int32), fwd_col=_p(k["fwd_col"], C.c_int32),
            fwd_w=_p(k["fwd_w"], C.c_double), fwd_link=_p(k["fwd_link"], C.c_int32),
            bwd_rowptr=_p(k["bwd_rowptr"], C.c_int32), bwd_col=_p(k["bwd_col"], C.c_int32),
            bwd_w=_p(k["bwd_w"], C.c_double), bwd_link=_p(k["bwd_link"], C.c_int32),
            link_len=_p(k["link_len"], C.c_double), link_v0=_p(k["link_v0"], C.c_double),
            link_t0=_p(k["link_t0"], C.c_double), link_cap=_p(k["link_cap"], C.c_int32),
            link_u=_p(k["link_u"], C.c_int32), link_v=_p(k["link_v"], C.c_int32))
        _capi.check(self._lib.drift_create(C.byref(self._h), int(device), C.byref(d)))
        try:
            self.set_bpr(alpha, beta)
            _capi.check(self._lib.drift_add_agents(self._h, self.N, self.agent_base, int(route_arena_links)))
        except Exception:
            self.close()
            raise

    # ---- lifecycle --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.drift_destroy(self._h)
            self._h = _capi.ENGINE()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        _capi.check(self._lib.drift_sync(self._h))

    # ---- configuration ------------------------------------------------------------------------
    def set_bpr(self, alpha, beta):
        table, off, cls = _bpr.build_tables(self.lg.link_cap, float(alpha), float(beta), max_count=self.total_agents)
        self.alpha, self.beta = float(alpha), float(beta)
        _capi.check(self._lib.drift_set_bpr(self._h, self.alpha, self.beta, _p(table, C.c_double),
                                            _p(off, C.c_int64), len(off) - 1, _p(cls, C.c_int32)))

    def set_demand(self, seed, hourly24, start_node, chain_capacity=4, router=ROUTER_NX_EXACT):
        h = _f64(hourly24)
        assert h.shape == (24,)
        s = _i32(start_node)
        assert len(s) == self.N
        _capi.check(self._lib.drift_set_demand(self._h, int(seed), _p(h, C.c_double), _p(s, C.c_int32),
                                               int(chain_capacity), int(router)))

    def push_trips(self, candidates):
        """Top up the per-agent destination rings from an [N, k] int32 host array (asynchronous copy;
        the array is kept alive by the engine object)."""
        c = _i32(candidates)
        assert c.ndim == 2 and c.shape[0] == self.N
        self._trip_keep = c
        _capi.check(self._lib.drift_push_trips(self._h, _p(c, C.c_int32), c.shape[1]))

    def store_trip_pool(self, pool):
        """Upload [n_blocks, N, k] candidate destinations once; ``refill_from_pool(b)`` tops up from block b."""
        c = _i32(pool)
        assert c.ndim == 3 and c.shape[1] == self.N
        _capi.check(self._lib.drift_store_trip_pool(self._h, _p(c, C.c_int32), c.shape[0], c.shape[2]))

    def refill_from_pool(self, block):
        _capi.check(self._lib.drift_refill_from_pool(self._h, int(block)))

    # ---- routing ------------------------------------------------------------------------------
    def route(self, src, dst, router=ROUTER_NX_EXACT, congested=False, fetch=True):
        """Route Q (src, dst) node-index pairs on the device.

        Returns (status int32[Q], n_links int32[Q], links list-of-arrays | None).  The routes stay
        staged on the device for ``commit_departures``.
        """
        src, dst = _i32(src), _i32(dst)
        Q = len(src)
        status = np.zeros(Q, dtype=np.int32)
        nl = np.zeros(Q, dtype=np.int32)
        _capi.check(self._lib.drift_route_od(self._h, int(router), Q, _p(src, C.c_int32), _p(dst, C.c_int32),
                                             int(bool(congested)), _p(status, C.c_int32), _p(nl, C.c_int32)))
        links = None
        if fetch:
            total = int(nl.sum())
            flat = np.zeros(max(total, 1), dtype=np.int32)
            if total:
                _capi.check(self._lib.drift_fetch_routes(self._h, _p(flat, C.c_int32), total))
            cuts = np.cumsum(nl)[:-1]
            links = np.split(flat[:total], cuts) if Q else []
        return status, nl, links

    def route_costs(self, Q):
        out = np.zeros(max(Q, 1))
        _capi.check(self._lib.drift_fetch_route_costs(self._h, _p(out, C.c_double), Q))
        return out[:Q]

    def commit_departures(self, agents):
        a = _i32(agents)
        _capi.check(self._lib.drift_commit_departures(self._h, len(a), _p(a, C.c_int32)))

    def submit_departures(self, agents, routes):
        """Departures with host-supplied link sequences (trace replay)."""
        a = _i32(agents)
        off = np.zeros(len(a) + 1, dtype=np.int64)
        if len(a):
            off[1:] = np.cumsum([len(r) for r in routes])
        flat = _i32(np.concatenate([np.asarray(r, dtype=np.int32) for r in routes])) if off[-1] else np.zeros(1, np.int32)
        _capi.check(self._lib.drift_submit_departures(self._h, len(a), _p(a, C.c_int32), _p(off, C.c_int64),
                                                      _p(flat, C.c_int32)))

    # ---- stepping -----------------------------------------------------------------------------
    def step(self, n_ticks, delta):
        _capi.check(self._lib.drift_step(self._h, int(n_ticks), float(delta)))

    def run(self, n_ticks, delta, t0):
        _capi.check(self._lib.drift_run(self._h, int(n_ticks), float(delta), float(t0)))

    def update_travel_times(self):
        _capi.check(self._lib.drift_update_travel_times(self._h))

    def update_travel_times_with(self, volumes_dev_ptr):
        """``update_travel_times`` from a device vector of link volumes (int32[L]; e.g. the all-reduced recount)."""
        _capi.check(self._lib.drift_update_travel_times_with(self._h, C.c_void_p(volumes_dev_ptr)))

    def reroute_moving(self, router=ROUTER_NX_EXACT):
        """Extension: re-plan every en-route agent from the head of its current link on the congested
        travel times (call ``update_travel_times`` first).  Returns the number of agents rerouted."""
        n = C.c_int64(0)
        _capi.check(self._lib.drift_reroute_moving(self._h, int(router), C.byref(n)))
        return n.value

    def invalidate_prefetch(self, t0, delta, n_ticks):
        _capi.check(self._lib.drift_invalidate_prefetch(self._h, float(t0), float(delta), int(n_ticks)))

    def tick_begin(self, delta):
        n = C.c_int64(0)
        _capi.check(self._lib.drift_tick_begin(self._h, float(delta), C.byref(n)))
        return n.value

    def tick_begin_demand(self, delta, t):
        n = C.c_int64(0)
        _capi.check(self._lib.drift_tick_begin_demand(self._h, float(delta), float(t), C.byref(n)))
        return n.value

    def set_stream(self, cuda_stream_handle):
        _capi.check(self._lib.drift_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def tick_begin_async(self, delta, t, send_ptr, cap):
        _capi.check(self._lib.drift_tick_begin_async(self._h, float(delta), float(t), C.c_void_p(send_ptr), int(cap)))

    def tick_finish_gathered(self, recv_ptr, world, cap):
        _capi.check(self._lib.drift_tick_finish_gathered(self._h, C.c_void_p(recv_ptr), int(world), int(cap)))

    def p2p_export(self, rank, world, n_total, cap):
        """Allocate the peer-memory inbox / flags of this rank; returns its 4 CUDA IPC handles (256 bytes)."""
        buf = (C.c_ubyte * 256)()
        _capi.check(self._lib.drift_p2p_export(self._h, int(rank), int(world), int(n_total), int(cap), buf))
        return bytes(buf)

    def p2p_connect(self, all_handles):
        """``all_handles``: the handles of all ranks in rank order (world x 256 bytes)."""
        buf = (C.c_ubyte * len(all_handles)).from_buffer_copy(all_handles)
        _capi.check(self._lib.drift_p2p_connect(self._h, buf))

    def max_events(self, reset=True):
        n = C.c_int64(0)
        _capi.check(self._lib.drift_max_events(self._h, C.byref(n), int(bool(reset))))
        return n.value

    def events_ptr(self):
        p, cap = C.c_void_p(), C.c_int64(0)
        _capi.check(self._lib.drift_events_ptr(self._h, C.byref(p), C.byref(cap)))
        return p.value, cap.value

    def tick_finish(self, events_dev_ptr, n_events):
        _capi.check(self._lib.drift_tick_finish(self._h, C.c_void_p(events_dev_ptr), int(n_events)))

    # ---- readback -----------------------------------------------------------------------------
    def occupancy(self):
        out = np.zeros(max(self.lg.L, 1), dtype=np.int32)
        _capi.check(self._lib.drift_read_occupancy(self._h, _p(out, C.c_int32)))
        return out[:self.lg.L]

    def recount_occupancy(self):
        out = np.zeros(max(self.lg.L, 1), dtype=np.int32)
        _capi.check(self._lib.drift_recount_occupancy(self._h, _p(out, C.c_int32)))
        return out[:self.lg.L]

    def recount_device(self):
        """Volume recount left in the device buffer ``volume_recount`` (asynchronous)."""
        _capi.check(self._lib.drift_recount_device(self._h))

    def travel_times(self):
        out = np.zeros(max(self.lg.L, 1))
        _capi.check(self._lib.drift_read_travel_times(self._h, _p(out, C.c_double)))
        return out[:self.lg.L]

    def agents(self, fields=("state", "dist", "speed", "elen", "cur_link", "route_idx")):
        """dict of per-agent arrays; fields among state, dist, speed, elen, cur_link, route_idx,
        route_len, trip_count, cur_node."""
        n = self.N
        spec = [("state", np.uint8, C.c_uint8), ("dist", np.float64, C.c_double), ("speed", np.float64, C.c_double),
                ("elen", np.float64, C.c_double), ("cur_link", np.int32, C.c_int32), ("route_idx", np.int32, C.c_int32),
                ("route_len", np.int32, C.c_int32), ("trip_count", np.int32, C.c_int32), ("cur_node", np.int32, C.c_int32)]
        out, args = {}, []
        for name, dt, ct in spec:
            if name in fields:
                out[name] = np.zeros(n, dtype=dt)
                args.append(_p(out[name], ct))
            else:
                args.append(None)
        _capi.check(self._lib.drift_read_agents(self._h, *args))
        return out

    def agent_route(self, agent):
        n = C.c_int64(0)
        _capi.check(self._lib.drift_read_agent_route(self._h, int(agent), None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.int32)
        if n.value:
            _capi.check(self._lib.drift_read_agent_route(self._h, int(agent), _p(out, C.c_int32), n.value, C.byref(n)))
        return out[:n.value]

    def export_routes(self, agents, with_links=False):
        """(n_links int32[], metres float64[], links list-of-arrays | None) of the agents' current routes."""
        a = _i32(agents)
        n = len(a)
        nl = np.zeros(max(n, 1), dtype=np.int32)
        m = np.zeros(max(n, 1))
        if n == 0:
            return nl[:0], m[:0], ([] if with_links else None)
        _capi.check(self._lib.drift_export_routes(self._h, n, _p(a, C.c_int32), _p(nl, C.c_int32), _p(m, C.c_double), None, 0))
        links = None
        if with_links:
            total = int(nl[:n].sum())
            flat = np.zeros(max(total, 1), dtype=np.int32)
            _capi.check(self._lib.drift_export_routes(self._h, n, _p(a, C.c_int32), _p(nl, C.c_int32), _p(m, C.c_double),
                                                      _p(flat, C.c_int32), total))
            links = np.split(flat[:total], np.cumsum(nl[:n])[:-1])
        return nl[:n], m[:n], links

    def _drain(self, fn, specs, chunk=1 << 18):
        cols = [[] for _ in specs]
        while True:
            bufs = [np.zeros(chunk, dtype=dt) for dt, _ in specs]
            n = C.c_int64(0)
            _capi.check(fn(self._h, *[_p(b, ct) for b, (_, ct) in zip(bufs, specs)], chunk, C.byref(n)))
            for c, b in zip(cols, bufs):
                c.append(b[:n.value])
            if n.value < chunk:
                break
        return [np.concatenate(c) for c in cols]

    def drain_arrivals(self):
        """(agent int32[], tick int32[]) of the arrivals since the last drain (device order)."""
        i32 = (np.int32, C.c_int32)
        return self._drain(self._lib.drift_drain_arrivals, [i32, i32])

    def drain_departures(self):
        """(agent, tick, src, dst, status) of device-driven departures since the last drain."""
        i32 = (np.int32, C.c_int32)
        return self._drain(self._lib.drift_drain_departures, [i32] * 5)

    def drain_departures_routes(self):
        """(agent, tick, src, dst, status, route_off, route_len, route_epoch): the departure log with the arena
        position of every route, for ``export_logged_routes``."""
        i32 = (np.int32, C.c_int32)
        return self._drain(self._lib.drift_drain_departures_routes, [i32] * 5 + [(np.int64, C.c_int64), i32, i32])

    def export_logged_routes(self, route_off, route_len, route_epoch, with_links=False):
        """(metres float64[], links list-of-arrays | None) of routes logged at departure; call before the next
        planning window can compact the arena (i.e. after every block of at most ``route_window`` ticks)."""
        off, ln, ep = _i64(route_off), _i32(route_len), _i32(route_epoch)
        n = len(off)
        m = np.zeros(max(n, 1))
        if n == 0:
            return m[:0], ([] if with_links else None)
        total = int(ln.sum()) if with_links else 0
        flat = np.zeros(max(total, 1), dtype=np.int32)
        _capi.check(self._lib.drift_export_logged_routes(self._h, n, _p(off, C.c_int64), _p(ln, C.c_int32),
                                                         _p(ep, C.c_int32), _p(m, C.c_double),
                                                         _p(flat, C.c_int32) if with_links else None, total))
        links = np.split(flat[:total], np.cumsum(ln)[:-1]) if with_links else None
        return m[:n], links

    def discard_logs(self):
        """Advance the arrival / departure rings past everything logged so far without copying it out
        (null output pointers): for runs that never read the logs but must not overflow them."""
        n = C.c_int64(0)
        big = 1 << 40
        _capi.check(self._lib.drift_drain_arrivals(self._h, None, None, big, C.byref(n)))
        _capi.check(self._lib.drift_drain_departures(self._h, None, None, None, None, None, big, C.byref(n)))

    def enable_entry_trace(self, cap=1 << 20):
        _capi.check(self._lib.drift_enable_entry_trace(self._h, int(cap)))

    def drain_entries(self):
        """(tick, agent, link, speed) of traced link entries since the last drain (device order)."""
        i32 = (np.int32, C.c_int32)
        return self._drain(self._lib.drift_drain_entries, [i32, i32, i32, (np.float64, C.c_double)])

    def stats(self):
        s = _capi.StatsOut()
        _capi.check(self._lib.drift_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in _capi.StatsOut._fields_}

    @property
    def tick(self):
        t = C.c_int64(0)
        _capi.check(self._lib.drift_tick_index(self._h, C.byref(t)))
        return t.value

    def phase_times(self, reset=True):
        a, i, r, pl = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0)
        n = C.c_int64(0)
        _capi.check(self._lib.drift_phase_times(self._h, C.byref(a), C.byref(i), C.byref(r), C.byref(pl), C.byref(n),
                                                int(bool(reset))))
        t = max(n.value, 1)
        return dict(ticks=n.value, advance_us_per_tick=a.value / t / 1e3, inject_us_per_tick=i.value / t / 1e3,
                    resolve_us_per_tick=r.value / t / 1e3, plan_ms=pl.value)

    def p2p_phase_detail(self, ticks, reset=True):
        """Microseconds per tick of the peer-exchange tick kernel's sections (block 0's clock)."""
        out = np.zeros(8)
        _capi.check(self._lib.drift_p2p_phase_detail(self._h, _p(out, C.c_double), int(bool(reset))))
        names = ("advance", "block_fence", "grid_barrier_1", "publish", "wait_peer_flags", "link_events",
                 "grid_barrier_2", "resolve_and_fence")
        return {n: float(v) / max(int(ticks), 1) / 1e3 for n, v in zip(names, out)}

    def tree_stats(self, reset=True):
        ms = C.c_double(0)
        v = [C.c_int64(0) for _ in range(7)]
        _capi.check(self._lib.drift_tree_stats(self._h, C.byref(ms), *[C.byref(x) for x in v], int(bool(reset))))
        return dict(build_ms=ms.value, launches=v[0].value, trees=v[1].value, edges_relaxed=v[2].value,
                    improved=v[3].value, nodes_expanded=v[4].value, waves=v[5].value, far_visits=v[6].value)

    def time_passes(self, iters=20):
        """Per-launch microseconds of the volume recount (K3') and of the BPR travel-time pass (K4)."""
        a, b = C.c_double(0), C.c_double(0)
        _capi.check(self._lib.drift_time_passes(self._h, int(iters), C.byref(a), C.byref(b)))
        return dict(recount_us=a.value, travel_times_us=b.value)

    def set_landmarks(self, K=8, seed=0, tables=None):
        """Pick K far-apart landmarks and upload their free-flow distance arrays (host: scipy Dijkstra, a
        one-off setup cost of ~0.2 s per landmark on a 1 M-link graph).  ``tables`` = (from, to) of an earlier
        ``landmark_tables`` call on the same graph skips the Dijkstras."""
        if K <= 0 or self.lg.V < 2:
            _capi.check(self._lib.drift_set_landmarks(self._h, 0, None, None))
            return []
        picks, frm, to = tables if tables is not None else landmark_tables(self.lg, K, seed)
        _capi.check(self._lib.drift_set_landmarks(self._h, len(picks), _p(frm, C.c_double), _p(to, C.c_double)))
        return picks

    def set_hierarchy(self, hier):
        """Upload a contraction hierarchy (``drift_b200.hierarchy.build``) and build the hub labels on the device.
        From then on ``ROUTER_BATCHED`` answers queries on the static travel times by label lookups.  Raises
        ``_capi.DriftError`` when the graph has no usable hierarchy (labels too large): the engine keeps its
        tree / bounded-search router."""
        d = hier.desc()
        self._hier_keep = hier
        _capi.check(self._lib.drift_set_hierarchy(self._h, C.byref(d)))

    def label_stats(self, reset=True):
        ent, lau, q, rd = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        bms, qms = C.c_double(0), C.c_double(0)
        _capi.check(self._lib.drift_label_stats(self._h, C.byref(ent), C.byref(bms), C.byref(qms), C.byref(lau),
                                                C.byref(q), C.byref(rd), int(bool(reset))))
        return dict(label_entries=ent.value, label_bytes=20 * ent.value, build_ms=bms.value, query_ms=qms.value,
                    launches=lau.value, queries=q.value, entries_read=rd.value)

    def set_option(self, name, value):
        _capi.check(self._lib.drift_set_option(self._h, name.encode(), float(value)))

    def set_profiling(self, on=True):
        _capi.check(self._lib.drift_set_profiling(self._h, int(bool(on))))

    def kernel_times(self):
        a, t = C.c_double(0), C.c_double(0)
        an, tn = C.c_int64(0), C.c_int64(0)
        _capi.check(self._lib.drift_kernel_times(self._h, C.byref(a), C.byref(an), C.byref(t), C.byref(tn)))
        return dict(advance_ms=a.value, advance_launches=an.value, total_ms=t.value, total_launches=tn.value)

    def device_ptr(self, name):
        p, n = C.c_void_p(), C.c_int64(0)
        _capi.check(self._lib.drift_device_ptr(self._h, name.encode(), C.byref(p), C.byref(n)))
        return p.value, n.value


def landmark_tables(lg, K=8, seed=0):
    """(landmark nodes, distances landmark -> node [K][V], distances node -> landmark [K][V]) of K landmarks picked by
    farthest-point sampling on the free-flow travel times."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    V = lg.V
    W = csr_matrix((lg.fwd_w, lg.fwd_col, lg.fwd_rowptr), shape=(V, V))
    WT = csr_matrix((lg.bwd_w, lg.bwd_col, lg.bwd_rowptr), shape=(V, V))
    rng = np.random.default_rng(seed)
    first = int(rng.integers(0, V))
    d0 = dijkstra(W, directed=True, indices=first)
    d0[~np.isfinite(d0)] = -1.0
    picks = [int(np.argmax(d0))]  # farthest node from a random one
    frm, to = [], []
    mind = None
    for k in range(K):
        f = dijkstra(W, directed=True, indices=picks[-1])
        t = dijkstra(WT, directed=True, indices=picks[-1])
        frm.append(f)
        to.append(t)
        reach = np.where(np.isfinite(f), f, -1.0)
        mind = reach if mind is None else np.minimum(mind, reach)
        if k + 1 < K:
            picks.append(int(np.argmax(mind)))  # farthest-point sampling
    frm = np.ascontiguousarray(np.stack(frm), dtype=np.float64)
    to = np.ascontiguousarray(np.stack(to), dtype=np.float64)
    return picks, frm, to

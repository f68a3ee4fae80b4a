"""Large-scale headless driver: device-driven demand in blocks of ticks, trip records streamed to CSV.

The reference keeps every trip as a dict in ``SimulationDataLogger.trips_data`` and writes one CSV at
the end (lib/data_logger.py:122-152); at 10 M agents x 48 h that list would hold ~10^8 dicts.  Here
the departure / arrival logs of each block come back from the device as arrays, are matched per
agent and appended to the CSV in chunks, with the reference's ten columns, ``T%06d`` ids in trip
start order, ``HH:MM:SS`` stamps counted from 06:00:00 and -- like the reference -- every agent's
first row carrying origin = destination = its initial node and the start time 06:00:00
(lib/simulation_thread.py:274-276).  Trips still open at ``finish`` are flushed with
the end-of-run clock, never-departed agents get their zero-length row.
"""
from __future__ import annotations

import csv
import datetime

import numpy as np

from .data_logger import FIELDNAMES, TripJsonStream, TripStatistics
from .engine import Engine, ROUTE_OK, ROUTER_BATCHED

START = datetime.datetime(2025, 7, 10, 6, 0, 0)  # lib/simulation_thread.py:154


def _stamp(t):
    return (START + datetime.timedelta(seconds=float(t))).strftime("%H:%M:%S")


class BlockSimulation:
    def __init__(self, lg, start_nodes, hourly24, seed=0, csv_path=None, ticks_per_block=300, delta=1.0,
                 router=ROUTER_BATCHED, reroute=False, record_routes=False, device=0, alpha=0.15, beta=4.0,
                 chain_capacity=4, route_arena_links=0, keep_log=False, hierarchy=None, json_path=None):
        self.lg = lg
        self.N = len(start_nodes)
        self.start_nodes = np.asarray(start_nodes, dtype=np.int32)
        self.delta = float(delta)
        self.ticks_per_block = int(ticks_per_block)
        self.reroute = reroute
        self.router = router
        self.record_routes = record_routes
        self.engine = Engine(lg, self.N, alpha=alpha, beta=beta, device=device,
                             route_arena_links=route_arena_links or max(1 << 22, self.N * 1024))
        self.keep_log = keep_log
        self.keep_rows = False  # set to keep the rows in self.rows when no CSV is written
        self.departure_log = []  # (tick, agent, src, dst, status) of every departure (keep_log only)
        if hierarchy is not None:  # static travel times: route by hub labels (drift_b200.hierarchy)
            self.engine.set_hierarchy(hierarchy)
        self.engine.set_demand(seed, np.asarray(hourly24, dtype=np.float64), self.start_nodes,
                               chain_capacity=chain_capacity, router=router)
        # one planning window per block: the arena is only compacted at a window start, so the routes logged during
        # a block are still in place when the block's logs are drained (export_logged_routes)
        self.engine.set_option("route_window", self.ticks_per_block)
        self.t = 0.0
        self.trip_counter = self.N  # ids T000001..T%06d(N) belong to the init-time trips of the reference
        self.rows_written = 0
        self.reroutes = 0
        # open trip per agent: (trip number, start time, origin, destination, metres, route)
        self._open_id = np.arange(1, self.N + 1, dtype=np.int64)
        self._open_t0 = np.zeros(self.N)
        self._open_src = self.start_nodes.copy()
        self._open_dst = self.start_nodes.copy()
        self._open_m = np.zeros(self.N)
        self._target = self.start_nodes.copy()        # Agent.target: what end_trip appends to route_taken
        self._open = np.ones(self.N, dtype=bool)      # the reference opens a trip for every agent at init
        self._first = np.ones(self.N, dtype=bool)     # first-trip quirk still pending
        self._routes = {}                             # agent -> node list (record_routes only)
        self.statistics = TripStatistics()            # get_statistics() without the list of trips
        self.rows = []
        self._fh = self._csv = None
        self._json = TripJsonStream(json_path, START) if json_path else None  # the reference's JSON export, streamed
        if csv_path:
            self._fh = open(csv_path, "w", newline="", encoding="utf-8")
            self._csv = csv.DictWriter(self._fh, fieldnames=FIELDNAMES)
            self._csv.writeheader()

    # ---- one block --------------------------------------------------------------------------------
    def run_block(self, candidates):
        """``candidates``: int32 [N, k] destinations produced by the s-t model for this block."""
        eng = self.engine
        eng.push_trips(candidates)
        if self.reroute and eng.tick > 0:
            eng.update_travel_times()
            self.reroutes += eng.reroute_moving(self.router)
            eng.invalidate_prefetch(self.t, self.delta, self.ticks_per_block)
        eng.run(self.ticks_per_block, self.delta, self.t)
        t_block0 = self.t
        for _ in range(self.ticks_per_block):
            self.t += self.delta
        dep = eng.drain_departures_routes()
        arr = eng.drain_arrivals()
        self._account(dep, arr, eng.tick - self.ticks_per_block, t_block0)
        return len(dep[0]), len(arr[0])

    def _clock(self, tick, tick0, t_block0):
        # the engine's clock is a repeated fp64 sum; for representable deltas this product is identical
        return t_block0 + (tick - tick0) * self.delta

    def _account(self, dep, arr, tick0, t_block0):
        d_agent, d_tick, d_src, d_dst, d_status, d_off, d_len, d_epoch = dep
        a_agent, a_tick = arr
        ok = d_status == ROUTE_OK
        ok_idx = np.nonzero(ok)[0]
        # length (and links) of EVERY departure's route, by its arena position at departure time: exact also for
        # agents that depart twice in a block or whose finished route was dropped by a compaction later on
        metres, links = self.engine.export_logged_routes(d_off[ok_idx], d_len[ok_idx], d_epoch[ok_idx],
                                                         with_links=self.record_routes)
        if self.keep_log:
            self.departure_log.extend(zip(d_tick.tolist(), d_agent.tolist(), d_src.tolist(), d_dst.tolist(),
                                          d_status.tolist()))
        # merge the two logs in (tick, agent) order; a trip starts before it can end
        ev = [(int(t), int(a), 0, k) for k, (t, a) in enumerate(zip(d_tick[ok], d_agent[ok]))] + \
             [(int(t), int(a), 1, k) for k, (t, a) in enumerate(zip(a_tick, a_agent))] + \
             [(int(d_tick[k]), int(d_agent[k]), 2, int(k)) for k in np.nonzero(~ok)[0]]
        ev.sort()
        rows = []
        for tick, agent, kind, k in ev:
            t = self._clock(tick, tick0, t_block0)
            if kind == 2:  # no route / trivial route: the agent stays where it is, but its target is already
                self._target[agent] = d_dst[k]  # overwritten (lib/agent.py:51) -- end_trip appends it
                continue
            if kind == 0:
                kk = int(ok_idx[k])
                if not self._open[agent]:  # lib/simulation_thread.py:344-347
                    self.trip_counter += 1
                    self._open_id[agent] = self.trip_counter
                    self._open_t0[agent] = t
                    self._open_src[agent] = d_src[kk]
                    self._open_dst[agent] = d_dst[kk]
                    self._open[agent] = True
                self._open_m[agent] = float(metres[k])
                self._target[agent] = d_dst[kk]
                if self.record_routes:
                    self._routes[agent] = self.lg.links_to_nodes(int(d_src[kk]), links[k])
            else:
                if self._open[agent]:
                    rows.append(self._close(agent, t))
        self._write(rows)

    def _route_taken(self, agent, upto=None):
        """``route_taken`` as the reference materialises it: [origin] at start_trip (lib/data_logger.py:29), then
        path[path_index] on every moving tick unless falsy or already present (:44-53), the target at end_trip
        (:75-76).  ``upto``: path index reached (None = arrived)."""
        nodes = self.lg.nodes
        o = nodes[int(self._open_src[agent])]
        path = self._routes.pop(agent, None)
        route = [o]
        if path is None:
            return None
        seen = {o}
        last = len(path) - 2 if upto is None else min(int(upto), len(path) - 1)
        for x in path[:last + 1]:
            lab = nodes[x]
            if not lab or lab in seen:
                continue
            seen.add(lab)
            route.append(lab)
        tgt = nodes[int(self._target[agent])]
        if tgt not in seen:
            route.append(tgt)
        return route

    def _close(self, agent, t, upto=None, moving=False):
        minutes = (float(t) - float(self._open_t0[agent])) / 60.0
        km = float(self._open_m[agent]) / 1000.0
        self.statistics.add(km, minutes, (km / minutes) * 60 if minutes > 0 else 0.0)
        if self._csv is None and self._json is None and not self.keep_rows:  # nobody reads the row: count it only
            self._routes.pop(agent, None)
            self._open[agent] = False
            self._first[agent] = False
            return None
        nodes = self.lg.nodes
        o, d = int(self._open_src[agent]), int(self._open_dst[agent])
        route = self._route_taken(agent, upto) if self.record_routes else None
        if route is None:  # routes are not recorded (or the agent never left): origin, and the target once under way
            route = [nodes[o]]
            tgt = nodes[int(self._target[agent])]
            if (moving or upto is None) and self._open_m[agent] > 0 and tgt != nodes[o]:
                route.append(tgt)
        row = {
            "trip_id": "T%06d" % self._open_id[agent],
            "agent_id": "A%06d" % (agent + 1),
            "start_time": _stamp(self._open_t0[agent]),
            "end_time": _stamp(t),
            "origin_node": nodes[o],
            "destination_node": nodes[d],
            "route_taken": str(route),
            "trip_distance_km": km,
            "trip_duration_min": minutes,
            "average_speed_kmh": (km / minutes) * 60 if minutes > 0 else 0.0,
        }
        self._open[agent] = False
        self._first[agent] = False
        return row

    def _write(self, rows):
        if self._csv is not None:
            self._csv.writerows(rows)
        if self._json is not None:
            self._json.write(rows)
        if self.keep_rows and self._csv is None:
            self.rows.extend(rows)
        self.rows_written += len(rows)

    def finish(self):
        """lib/simulation_thread.py:483-489: close every open trip in agent order -- agents still under way with the
        part of the route they have covered, never-departed agents with their zero-length row."""
        ag = self.engine.agents(fields=("state", "route_idx"))
        rows = []
        for a in np.nonzero(self._open)[0]:
            a = int(a)
            mv = ag["state"][a] == 1
            rows.append(self._close(a, self.t, upto=int(ag["route_idx"][a]) if mv else (None if a in self._routes else 0),
                                    moving=bool(mv)))
        self._write(rows)
        if self._fh is not None:
            self._fh.close()
            self._fh = None
        if self._json is not None:
            self._json.close()
        return self.rows_written

    def get_statistics(self):
        """lib/data_logger.py:154-171 over every trip closed so far (running sums, no list of trips kept)."""
        return self.statistics.result()

    def status(self):
        """Payload of the reference's ``status_updated`` signal (lib/simulation_thread.py:455-462)."""
        s = self.engine.stats()
        return {
            "simulation_time": self.t,
            "moving_agents": s["moving_agents"],
            "total_agents": self.N,
            "network_utilization": s["total_agents_on_roads"] / s["total_capacity"] if s["total_capacity"] else 0,
            "active_agent_types": {},
        }

    def close(self):
        if self._fh is not None:
            self._fh.close()
        if self._json is not None:
            self._json.close()
        self.engine.close()

"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the *parallel* tick formulation, shardable.

The reference updates agents sequentially (lib/simulation_thread.py:299-303); the CUDA path turns
a tick into link events and resolves them in (link, agent) order.  This model does
the same on the CPU for an agent block [lo, hi) of a run sharded over ranks: ``tick_begin`` yields
the block's events, ``tick_finish`` consumes the events of *all* ranks.  tests/test_dist_cpu.py
shows, over gloo, that the sharded model equals the sequential port (oracle/port.py) bit for bit,
which is the property the multi-GPU path relies on.
"""
import numpy as np

from .port import KPH_TO_MPS, bpr_factor


class EventModel:
    def __init__(self, pg, n_total, lo, hi, alpha=0.15, beta=4.0):
        self.pg, self.lo, self.hi = pg, lo, hi
        self.alpha, self.beta = alpha, beta
        n = hi - lo
        self.state = np.zeros(n, dtype=np.uint8)
        self.dist = np.zeros(n)
        self.speed = np.zeros(n)
        self.elen = np.zeros(n)
        self.cur_link = np.full(n, -1, dtype=np.int64)
        self.route = [None] * n
        self.ridx = np.zeros(n, dtype=np.int64)
        self.occ = np.zeros(pg.L, dtype=np.int64)
        self.arrivals = []
        self.tick = 0

    def tick_begin(self, delta, departures):
        """departures: {global agent: (src, dst, node path | None)} of agents in [lo, hi)."""
        self.tick += 1
        ev = []
        mv = np.nonzero(self.state == 1)[0]
        nd = self.dist[mv] + self.speed[mv] * delta  # two roundings, elementwise like the reference
        cross = nd >= self.elen[mv]
        self.dist[mv[~cross]] = nd[~cross]
        for i in mv[cross].tolist():
            self.dist[i] = 0.0
            self.ridx[i] += 1
            ev.append((self.cur_link[i], self.lo + i, -1, 0))
            r = self.route[i]
            if self.ridx[i] < len(r):
                self.cur_link[i] = r[self.ridx[i]]
                ev.append((self.cur_link[i], self.lo + i, 1, 0))
            else:
                self.cur_link[i] = -1
                self.state[i] = 0
                self.arrivals.append((self.tick, self.lo + i))
        for a, (_s, _t, path) in sorted(departures.items()):
            if path is None or len(path) < 2:
                continue
            i = a - self.lo
            links = [self.pg.hop[u][v][0] for u, v in zip(path[:-1], path[1:])]
            self.route[i], self.ridx[i], self.dist[i] = links, 0, 0.0
            self.state[i] = 1
            self.cur_link[i] = links[0]
            ev.append((links[0], a, 1, 0))
        return np.array(ev, dtype=np.int32).reshape(-1, 4)

    def tick_finish(self, events):
        if len(events) == 0:
            return
        order = np.lexsort((events[:, 1], events[:, 0]))  # by link, then global agent id
        ev = events[order]
        run_link, run = -1, 0
        for link, agent, kind, _ in ev.tolist():
            if link != run_link:
                if run_link >= 0:
                    self.occ[run_link] = run
                run_link, run = link, int(self.occ[link])
            run += kind
            if kind > 0 and self.lo <= agent < self.hi:
                i = agent - self.lo
                f = bpr_factor(run, int(self.pg.link_cap[link]), self.alpha, self.beta)
                self.speed[i] = (float(self.pg.link_v0[link]) * f) * KPH_TO_MPS
                self.elen[i] = float(self.pg.link_len[link])
        self.occ[run_link] = run

    # ---- link-owner protocol (what k_tick_loop_p2p does over NVLink peer memory) ----------------
    def resolve_owned(self, events, rank, world):
        """Owner side: ``events`` = the events of ALL ranks on the links this rank owns (link % world ==
        rank).  Updates the owned links' occupancy and returns [(global agent, speed)] for every entering
        agent, to be delivered to the agent's home rank.  elen is not returned: the home rank knows the
        link it entered (link tables are replicated)."""
        out = []
        if len(events) == 0:
            return out
        assert np.all(events[:, 0] % world == rank)
        order = np.lexsort((events[:, 1], events[:, 0]))
        ev = events[order]
        run_link, run = -1, 0
        for link, agent, kind, _ in ev.tolist():
            if link != run_link:
                if run_link >= 0:
                    self.occ[run_link] = run
                run_link, run = link, int(self.occ[link])
            run += kind
            if kind > 0:
                f = bpr_factor(run, int(self.pg.link_cap[link]), self.alpha, self.beta)
                out.append((agent, (float(self.pg.link_v0[link]) * f) * KPH_TO_MPS))
        self.occ[run_link] = run
        return out

    def apply_speeds(self, results):
        """Home side: entry speeds of this rank's agents as delivered by the owners of their links."""
        for agent, sp in results:
            i = agent - self.lo
            self.speed[i] = sp
            self.elen[i] = float(self.pg.link_len[self.cur_link[i]])

"""Host face of the device-side traffic state: same attributes and methods as the reference's
``TrafficManager`` (lib/managers/traffic_manager.py:10-165).  Occupancy ("link volume") lives in the
engine as one int32 per link; this class only translates between (u, v, key) labels and link ids
for the UI managers, status updates and exports that read it."""
from __future__ import annotations

from . import bpr as _bpr


class _EdgeAgents:
    """Read-only mapping (u, v, key) -> list of agents currently on that link, built on demand from a
    device snapshot (the reference keeps Python lists, traffic_manager.py:19)."""

    def __init__(self, tm):
        self._tm = tm

    def _build(self):
        tm = self._tm
        th = tm._thread
        out = {}
        if th is None:
            return out
        snap = th._snapshot()
        lg = tm._lg
        for i, l in enumerate(snap["cur_link"].tolist()):
            if l >= 0 and snap["state"][i] == 1:
                out.setdefault(lg.link_tuple(l), []).append(th.agents[i])
        return out

    def __getitem__(self, key):
        return self._build().get(key, [])

    def get(self, key, default=None):
        return self._build().get(key, default)

    def items(self):
        return self._build().items()

    def values(self):
        return self._build().values()

    def keys(self):
        return self._build().keys()

    def __iter__(self):
        return iter(self._build())

    def __len__(self):
        return len(self._build())

    def clear(self):
        pass


class TrafficManager:
    def __init__(self, graph, lowered, engine, thread=None):
        self.graph = graph
        self._lg = lowered
        self._engine = engine
        self._thread = thread
        self._caps = None
        self.bpr_alpha = engine.alpha
        self.bpr_beta = engine.beta
        self.edge_agents = _EdgeAgents(self)

    @property
    def edge_capacities(self):
        if self._caps is None:
            lg = self._lg
            self._caps = {lg.link_tuple(l): int(c) for l, c in enumerate(lg.link_cap.tolist())}
        return self._caps

    def set_bpr_parameters(self, alpha, beta):
        self.bpr_alpha, self.bpr_beta = alpha, beta
        self._engine.set_bpr(alpha, beta)  # re-tabulates the factor on the host, uploads it

    # occupancy is maintained by the tick kernels
    def register_agent_on_edge(self, agent, u, v, key=0):
        pass

    def unregister_agent_from_edge(self, agent, u, v, key=0):
        pass

    def clear_edge_agents(self):
        pass

    def _link_id(self, u, v, key):
        lg = self._lg
        ui, vi = lg.node_index.get(u), lg.node_index.get(v)
        if ui is None or vi is None:
            return -1
        import numpy as np

        hits = np.nonzero((lg.link_u == ui) & (lg.link_v == vi))[0]
        for l in hits.tolist():
            if lg.link_key is None or lg.link_key[l] == key:
                return l
        return -1

    def _count(self, u, v, key):
        l = self._link_id(u, v, key)
        return (int(self._engine.occupancy()[l]), int(self._lg.link_cap[l])) if l >= 0 else (0, 10)

    def get_congested_speed(self, agent, u, v, key=0):
        count, cap = self._count(u, v, key)
        return _bpr.factor(count, cap, self.bpr_alpha, self.bpr_beta)

    def get_edge_congestion_level(self, u, v, key=0):
        count, cap = self._count(u, v, key)
        return min(1.0, count / cap)

    def get_network_statistics(self):
        s = self._engine.stats()
        total_edges = self._lg.L
        counts = {}
        th = self._thread
        if th is not None:
            for a in th.agents:
                if a.state == "moving":
                    counts[str(a.agent_type)] = counts.get(str(a.agent_type), 0) + 1
        cap = s["total_capacity"]
        return {
            "total_agents_on_roads": s["total_agents_on_roads"],
            "total_network_capacity": cap,
            "network_utilization": s["total_agents_on_roads"] / cap if cap > 0 else 0,
            "congested_edges": s["links_with_traffic"],
            "total_edges": total_edges,
            "congestion_ratio": s["links_with_traffic"] / total_edges if total_edges > 0 else 0,
            "active_agent_types": counts,
        }

"""TEST INFRASTRUCTURE ONLY -- reader for the ``tests/golden/*.npz`` fixtures.

The fixtures were recorded from the unmodified reference by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from .port import PlainGraph, link_capacity

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class GoldenCase:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.z = z
        self.nodes = json.loads(str(z["nodes_json"]))
        self.link_keys = json.loads(str(z["link_key_json"]))
        self.params = json.loads(str(z["params_json"]))
        self.link_u, self.link_v = z["link_u"], z["link_v"]
        self.link_len, self.link_v0, self.link_tt = z["link_len"], z["link_v0"], z["link_tt"]
        self.link_lanes, self.link_has_lanes = z["link_lanes"], z["link_has_lanes"]
        self.capacities = z["capacities"]
        self.V, self.L = len(self.nodes), len(self.link_u)
        self.ticks = int(z["ticks"])
        self.init_nodes = z["init_nodes"]
        self.agent_types0 = json.loads(str(z["agent_types0_json"]))
        self.N = len(self.init_nodes)
        self.times = z["times"]
        self.trips = json.loads(str(z["trips_json"]))
        self.stats = json.loads(str(z["stats_json"]))
        self.finish_error = str(z["finish_error"])
        self.trips_before_finish = int(z["trips_before_finish"])
        st = self.params.get("settings") or {}
        self.alpha = st.get("bpr_alpha", 0.15)
        self.beta = st.get("bpr_beta", 4.0)
        hp = st.get("hourly_probabilities")
        self.hourly = None if hp is None else {int(k): v for k, v in hp.items()}
        self.accel = self.params["accel"]
        self.hours = self.params["hours"]

    # ---- trace views ----------------------------------------------------
    def departures(self):
        z = self.z
        out = []
        off = z["dep_path_off"]
        for i in range(len(z["dep_tick"])):
            n = int(z["dep_path_len"][i])
            path = None if n < 0 else z["dep_path_nodes"][off[i]:off[i + 1]].tolist()
            out.append((int(z["dep_tick"][i]), int(z["dep_agent"][i]), int(z["dep_src"][i]),
                        int(z["dep_dst"][i]), path))
        return out

    def departure_trace(self):
        tr = {}
        for tick, agent, s, t, path in self.departures():
            tr.setdefault(tick, {})[agent] = (s, t, path)
        return tr

    def entries(self):
        z = self.z
        return list(zip(z["ent_tick"].tolist(), z["ent_agent"].tolist(), z["ent_link"].tolist(),
                        z["ent_speed"].tolist(), z["ent_elen"].tolist()))

    def arrivals(self):
        z = self.z
        return list(zip(z["arr_tick"].tolist(), z["arr_agent"].tolist()))

    def occupancy(self):
        """Dense int32[ticks, L] rebuilt from the sparse change list."""
        z = self.z
        occ = np.zeros((self.ticks, self.L), dtype=np.int32)
        cur = np.zeros(self.L, dtype=np.int32)
        ct, cl, cv = z["occ_tick"], z["occ_link"], z["occ_val"]
        j = 0
        for k in range(1, self.ticks + 1):
            while j < len(ct) and ct[j] == k:
                cur[cl[j]] = cv[j]
                j += 1
            occ[k - 1] = cur
        return occ

    def snapshots(self):
        z = self.z
        return {int(t): (z["snap_state"][i], z["snap_d"][i], z["snap_speed"][i], z["snap_pidx"][i])
                for i, t in enumerate(z["snap_ticks"])}

    # ---- graph views ----------------------------------------------------
    def plain_graph(self) -> PlainGraph:
        """PlainGraph straight from the arrays (no NetworkX involved)."""
        V = self.V
        links = [(self.nodes[u], self.nodes[v], k) for u, v, k in zip(self.link_u, self.link_v, self.link_keys)]
        groups = [dict() for _ in range(V)]  # u -> {v: [link ids]} in succ order
        for i, (u, v) in enumerate(zip(self.link_u.tolist(), self.link_v.tolist())):
            groups[u].setdefault(v, []).append(i)
        succ, pair_w = [], {}
        for u in range(V):
            row = []
            for v, ids in groups[u].items():
                wmin = min(self.link_tt[i] for i in ids)
                best = ids[0]
                for i in ids:
                    if self.link_tt[i] < self.link_tt[best]:
                        best = i
                lmin = min(self.link_len[i] for i in ids)
                row.append((v, float(wmin), best, float(lmin)))
                pair_w[(u, v)] = float(wmin)
            succ.append(row)
        pr, pc = self.z["pred_rowptr"], self.z["pred_col"]
        pred = [[(int(u), pair_w[(int(u), v)]) for u in pc[pr[v]:pr[v + 1]]] for v in range(V)]
        return PlainGraph(self.nodes, links, self.link_len.copy(), self.link_v0.copy(), self.link_tt.copy(),
                          self.capacities.astype(np.int64), succ, pred)

    def nx_graph(self):
        """Rebuild the loader-processed ``nx.MultiDiGraph`` with identical node, ``_succ`` and ``_pred`` order."""
        import networkx as nx

        G = nx.MultiDiGraph()
        pos = self.z["pos"]
        for i, n in enumerate(self.nodes):
            G.add_node(n, pos=(float(pos[i, 0]), float(pos[i, 1])))
        for i in range(self.L):
            u, v = self.nodes[self.link_u[i]], self.nodes[self.link_v[i]]
            attrs = dict(length=float(self.link_len[i]), speed_kph=float(self.link_v0[i]),
                         travel_time=float(self.link_tt[i]))
            if self.link_has_lanes[i]:
                attrs["lanes"] = float(self.link_lanes[i])
            G.add_edge(u, v, key=self.link_keys[i], **attrs)
        pr, pc = self.z["pred_rowptr"], self.z["pred_col"]
        for vi, v in enumerate(self.nodes):
            order = [self.nodes[u] for u in pc[pr[vi]:pr[vi + 1]]]
            old = G._pred[v]
            assert set(order) == set(old)
            new = {u: old[u] for u in order}
            old.clear()
            old.update(new)
        return G

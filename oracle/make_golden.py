"""TEST INFRASTRUCTURE ONLY -- freeze live-reference runs into ``tests/golden/``.

Run in the build container (needs ``/root/reference`` and NetworkX 3.6.1):

    python -m oracle.make_golden            # writes tests/golden/*.npz

Every fixture is produced by the UNMODIFIED reference driven by
``oracle/ref_harness.py``; nothing in the reference is copied, only its
observable behaviour on seeded inputs is recorded.  The fixtures travel to the
GPU box, the reference does not.

Fixture layout (one compressed ``.npz`` per case, see ``oracle/golden.py`` for the reader):
  graph   node labels (JSON), links (u, v, key, length, speed_kph, travel_time, lanes, cap)
          in ``G.edges(keys=True)`` order, plus the ``G._pred`` insertion order
  run     agents, mode, seed, accel, hours, bpr alpha/beta, hourly probabilities
  trace   initial nodes, departures (tick, agent, src, dst, path), link entries
          (tick, agent, link, fp64 speed, fp64 length), arrivals (tick, agent),
          sparse per-tick occupancy changes, fp64 state snapshots every 50 ticks,
          the clock after every tick, and the final trip records (CSV rows).
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

from . import ref_harness as H

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
EX = os.path.join(H.REFERENCE_ROOT, "example-data")


def pack_graph(G):
    nodes = list(G.nodes)
    idx = {n: i for i, n in enumerate(nodes)}
    links = list(G.edges(keys=True, data=True))
    L = len(links)
    out = dict(
        nodes_json=json.dumps(nodes),
        link_u=np.array([idx[u] for u, _, _, _ in links], dtype=np.int32),
        link_v=np.array([idx[v] for _, v, _, _ in links], dtype=np.int32),
        link_key_json=json.dumps([k for _, _, k, _ in links]),
        link_len=np.array([float(d["length"]) for *_, d in links]),
        link_v0=np.array([float(d["speed_kph"]) for *_, d in links]),
        link_tt=np.array([float(d["travel_time"]) for *_, d in links]),
        link_lanes=np.array([float(d.get("lanes", 1)) for *_, d in links]),
        link_has_lanes=np.array(["lanes" in d for *_, d in links], dtype=np.uint8),
    )
    pos = [G.nodes[n].get("pos", (0.0, 0.0)) for n in nodes]
    out["pos"] = np.array(pos, dtype=np.float64)
    pr, pc = [0], []
    for v in nodes:
        for u in G._pred[v]:
            pc.append(idx[u])
        pr.append(len(pc))
    out["pred_rowptr"] = np.array(pr, dtype=np.int64)
    out["pred_col"] = np.array(pc, dtype=np.int32)
    assert L == len(out["link_u"])
    return out


def pack_run(r, params):
    out = {}
    out["params_json"] = json.dumps(params)
    out["init_nodes"] = r["init_nodes"]
    out["agent_types0_json"] = json.dumps(r["agent_types0"])
    dep = r["departures"]
    out["dep_tick"] = np.array([d[0] for d in dep], dtype=np.int32)
    out["dep_agent"] = np.array([d[1] for d in dep], dtype=np.int32)
    out["dep_src"] = np.array([d[2] for d in dep], dtype=np.int32)
    out["dep_dst"] = np.array([d[3] for d in dep], dtype=np.int32)
    off, flat = [0], []
    plen = []
    for d in dep:
        if d[4] is None:
            plen.append(-1)
        else:
            plen.append(len(d[4]))
            flat.extend(d[4])
        off.append(len(flat))
    out["dep_path_len"] = np.array(plen, dtype=np.int32)
    out["dep_path_off"] = np.array(off, dtype=np.int64)
    out["dep_path_nodes"] = np.array(flat, dtype=np.int32)
    en = r["entries"]
    out["ent_tick"] = np.array([e[0] for e in en], dtype=np.int32)
    out["ent_agent"] = np.array([e[1] for e in en], dtype=np.int32)
    out["ent_link"] = np.array([e[2] for e in en], dtype=np.int32)
    out["ent_speed"] = np.array([e[3] for e in en], dtype=np.float64)
    out["ent_elen"] = np.array([e[4] for e in en], dtype=np.float64)
    ar = r["arrivals"]
    out["arr_tick"] = np.array([a[0] for a in ar], dtype=np.int32)
    out["arr_agent"] = np.array([a[1] for a in ar], dtype=np.int32)
    occ = r["occ"]
    prev = np.zeros(occ.shape[1], dtype=np.int32)
    ct, cl, cv = [], [], []
    for k in range(occ.shape[0]):
        ch = np.nonzero(occ[k] != prev)[0]
        ct.extend([k + 1] * len(ch))
        cl.extend(ch.tolist())
        cv.extend(occ[k][ch].tolist())
        prev = occ[k]
    out["occ_tick"] = np.array(ct, dtype=np.int32)
    out["occ_link"] = np.array(cl, dtype=np.int32)
    out["occ_val"] = np.array(cv, dtype=np.int32)
    out["capacities"] = r["capacities"].astype(np.int32)
    ticks = sorted(r["snaps"])
    out["snap_ticks"] = np.array(ticks, dtype=np.int32)
    if ticks:
        out["snap_state"] = np.stack([r["snaps"][t][0] for t in ticks])
        out["snap_d"] = np.stack([r["snaps"][t][1] for t in ticks])
        out["snap_speed"] = np.stack([r["snaps"][t][2] for t in ticks])
        out["snap_pidx"] = np.stack([r["snaps"][t][3] for t in ticks])
    out["times"] = r["times"]
    out["ticks"] = np.int64(r["ticks"])
    out["trips_json"] = json.dumps(r["trips"])
    out["stats_json"] = json.dumps(r["stats"])
    out["trips_before_finish"] = np.int64(r["trips_before_finish"])
    out["finish_error"] = "" if r["finish_error"] is None else r["finish_error"]
    out["final_state"] = r["final_state"]
    out["trip_counts"] = r["trip_counts"]
    return out


def save_case(name, G, r, params):
    os.makedirs(OUT, exist_ok=True)
    d = {}
    d.update(pack_graph(G))
    d.update(pack_run(r, params))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("%-28s V=%d L=%d agents=%d ticks=%d departures=%d arrivals=%d trips=%d  %.0f kB" % (
        name, G.number_of_nodes(), G.number_of_edges(), len(r["init_nodes"]), r["ticks"],
        len(r["departures"]), len(r["arrivals"]), len(r["trips"]), os.path.getsize(path) / 1e3))


def case(name, G, agents, mode="random", seed=0, accel=10, hours=1, settings=None, max_ticks=None):
    r = H.run_reference(G, agents, mode, seed=seed, accel=accel, hours=hours, settings=settings,
                        max_ticks=max_ticks)
    params = dict(agents=agents, mode=mode, seed=seed, accel=accel, hours=hours,
                  settings=_jsonable(settings), max_ticks=max_ticks)
    save_case(name, G, r, params)
    return r


def _jsonable(settings):
    if settings is None:
        return None
    s = dict(settings)
    if "hourly_probabilities" in s:
        s["hourly_probabilities"] = {str(k): v for k, v in s["hourly_probabilities"].items()}
    return s


def congested_g40():
    """SURVEY.md section 8(d) config 1, 'congested variant' (probe P3)."""
    import networkx as nx

    G = H.load_graph(os.path.join(EX, "graphml", "small", "g.40.16.graphml"), seed=0)
    rng = random.Random(123)
    for u, v, k, d in G.edges(keys=True, data=True):
        d["length"] = rng.choice([25, 40, 60, 90, 150, 300])
        d["speed_kph"] = rng.choice([20, 30, 50, 70])
    H.ensure_edge_weights(G)
    return G


def multigraph_case():
    """Small hand-made MultiDiGraph: int labels incl. 0 (falsy-node quirk, data_logger.py:52),
    parallel edges with different travel_time/length, self-loops, lanes, exact fp ties."""
    import networkx as nx

    rng = random.Random(7)
    G = nx.MultiDiGraph()
    n = 30
    for i in range(n):
        G.add_node(i, pos=(rng.random(), rng.random()))
    for i in range(n):  # ring both ways keeps it strongly connected
        for j in ((i + 1) % n, (i - 1) % n):
            G.add_edge(i, j, length=rng.choice([50.0, 100.0, 150.0]), speed_kph=rng.choice([30, 50]),
                       lanes=rng.choice([1, 2]))
    for _ in range(60):
        a, b = rng.randrange(n), rng.randrange(n)
        G.add_edge(a, b, length=rng.choice([50.0, 75.0, 100.0, 200.0]), speed_kph=rng.choice([30, 50, 70]),
                   lanes=rng.choice([1, 2, 3]))
    for _ in range(25):  # parallel edges on existing pairs
        a, b, _k = rng.choice(list(G.edges(keys=True)))
        G.add_edge(a, b, length=rng.choice([40.0, 60.0, 100.0, 300.0]), speed_kph=rng.choice([20, 30, 50]))
    H.ensure_edge_weights(G)
    return G


def main():
    small = os.path.join(EX, "graphml", "small")
    # config 1 (BASELINE.json configs[0] as SURVEY.md 8(d) defines it)
    G100 = H.load_graph(os.path.join(small, "g.100.0.graphml"), seed=0)
    case("g100_random_s0", G100, 300, "random", seed=0, accel=10, hours=1)
    G40 = H.load_graph(os.path.join(small, "g.40.16.graphml"), seed=0)
    case("g40_random_s1", G40, 300, "random", seed=1, accel=10, hours=1)
    # tick lengths: GUI default 50x (5 s), shipped-CSV 100x (10 s), non-representable 7x
    case("g100_random_accel50", G100, 300, "random", seed=2, accel=50, hours=3)
    case("g100_random_accel7", G100, 200, "random", seed=3, accel=7, hours=1)
    # congestion: BPR factor < 1, caps 12-160
    Gc = congested_g40()
    hp = {h: 0.9 for h in range(24)}
    case("g40_congested", Gc, 2500, "random", seed=0, accel=10, hours=1,
         settings=dict(bpr_alpha=0.15, bpr_beta=4.0, hourly_probabilities=hp), max_ticks=2500)
    case("g40_congested_bpr", Gc, 1500, "random", seed=5, accel=10, hours=1,
         settings=dict(bpr_alpha=0.8, bpr_beta=2.5, hourly_probabilities=hp), max_ticks=1500)
    # the other four s-t models (selectors stay on host; their (s,t) draws are replayed)
    for mode in ("activity", "zone", "hub", "gravity"):
        case("g100_%s_s0" % mode, G100, 300, mode, seed=0, accel=10, hours=2)
    # directed music-net graph: unreachable pairs (NetworkXNoPath) and self-loops
    Gm = H.load_graph(os.path.join(EX, "music-net", "chpn_op10_e01.graphml"), seed=0) \
        if os.path.exists(os.path.join(EX, "music-net", "chpn_op10_e01.graphml")) else None
    if Gm is None:
        names = sorted(os.listdir(os.path.join(EX, "music-net")))
        Gm = H.load_graph(os.path.join(EX, "music-net", names[0]), seed=0)
    case("musicnet_random_s0", Gm, 300, "random", seed=0, accel=10, hours=1)
    # unreachable pairs + the finish_simulation TypeError quirk
    Gch = H.load_graph(os.path.join(EX, "music-net", "chpn_op35_4l_0.graphml"), seed=0)
    case("chpn_random_nopath", Gch, 300, "random", seed=0, accel=10, hours=2)
    Glz = H.load_graph(os.path.join(EX, "music-net", "liz_rhap09.graphml"), seed=0)
    case("liz_random_nopath", Glz, 300, "random", seed=0, accel=10, hours=2)
    # multigraph / int labels / falsy node 0
    case("multi_random_s0", multigraph_case(), 400, "random", seed=11, accel=10, hours=1,
         settings=dict(bpr_alpha=0.15, bpr_beta=4.0, hourly_probabilities={h: 0.6 for h in range(24)}))


if __name__ == "__main__":
    sys.exit(main())

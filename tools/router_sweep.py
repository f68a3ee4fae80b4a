"""Option sweep of the tree / bounded-search router on one graph and one batch of cold (source, destination) pairs:
the graph, its landmark tables and the queries are built once, every configuration gets a fresh engine, routes a
warm-up batch and then the timed batch (CUDA events of the search kernel via option time_plan + wall clock around
the call).  All configurations must return the same route costs (the shortest path is unique on continuous
weights); the tool checks that and prints one JSON line per configuration.

    python tools/router_sweep.py --graph grid --nodes 4000000 --queries 20000 \
        --configs "base:;nocache:pot_cache=0;d2:tree_delta_scale=2;narrow:tree_wide_regs=0"
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def parse_configs(text):
    out = []
    for item in text.split(";"):
        if not item.strip():
            continue
        name, _, opts = item.partition(":")
        kv = {}
        for o in opts.split(","):
            if o.strip():
                k, v = o.split("=")
                kv[k.strip()] = float(v)
        out.append((name.strip(), kv))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", default="grid", choices=["grid", "road", "rgg"])
    ap.add_argument("--nodes", type=int, default=4_000_000)
    ap.add_argument("--model", default="zone", choices=sorted(["zone", "gravity", "hub", "activity", "activity_pools", "random"]))
    ap.add_argument("--queries", type=int, default=20_000)
    ap.add_argument("--warm-queries", type=int, default=2_000)
    ap.add_argument("--landmarks", type=int, default=8)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--configs", default="base:")
    ap.add_argument("--arena-links", type=int, default=0,
                    help="size of the route arena (links): set it to the bench's to leave the searches the same memory")
    ap.add_argument("--out", default=None, help="append the JSON lines to this file too")
    args = ap.parse_args()

    from drift_b200 import build as _build

    _build.build()
    from drift_b200 import synth
    from drift_b200.engine import ROUTER_BATCHED, Engine, landmark_tables

    t0 = time.perf_counter()
    if args.graph == "grid":
        side = int(round(args.nodes ** 0.5))
        lg = synth.grid_graph(side, side, seed=args.seed, morton=True)
    elif args.graph == "rgg":
        lg = synth.rgg_graph(args.nodes, seed=args.seed)
    else:
        lg = synth.road_graph(args.nodes, seed=args.seed)
    t_graph = time.perf_counter() - t0
    rng = np.random.default_rng(args.seed + 1)
    nq = args.queries + args.warm_queries
    src = rng.integers(0, lg.V, size=nq).astype(np.int32)
    _, dst = bench.make_chains(args.model, rng, lg, src, 1)
    dst = np.ascontiguousarray(dst.reshape(nq)).astype(np.int32)
    t0 = time.perf_counter()
    tables = landmark_tables(lg, args.landmarks, args.seed) if args.landmarks > 0 else None
    t_lm = time.perf_counter() - t0
    sys.stderr.write("[router_sweep] graph %d nodes / %d links in %.1f s, landmarks %.1f s, %d queries\n" % (
        lg.V, lg.L, t_graph, t_lm, args.queries))
    ws, wd = src[:args.warm_queries], dst[:args.warm_queries]
    qs, qd = src[args.warm_queries:], dst[args.warm_queries:]
    ref_cost = None
    out_f = open(args.out, "a") if args.out else None
    for name, kv in parse_configs(args.configs):
        eng = Engine(lg, 1, route_arena_links=max(1 << 26, 2560 * args.queries, args.arena_links))
        try:
            eng.set_option("time_plan", 1)
            for k, v in kv.items():
                eng.set_option(k, v)
            if tables is not None:
                eng.set_landmarks(tables=tables)
            eng.route(ws, wd, router=ROUTER_BATCHED, fetch=False)
            eng.sync()
            eng.tree_stats(reset=True)
            t0 = time.perf_counter()
            status, nl, _ = eng.route(qs, qd, router=ROUTER_BATCHED, fetch=False)
            eng.sync()
            wall = time.perf_counter() - t0
            st = eng.tree_stats(reset=True)
            cost = eng.route_costs(len(qs))
            ok = status >= 0
            same = None
            if ref_cost is None:
                ref_cost = cost
            else:
                same = bool(np.array_equal(cost[ok], ref_cost[ok]))
            n = max(st["trees"], 1)
            line = dict(config=name, options=kv, graph=args.graph, nodes=lg.V, links=lg.L, queries=len(qs),
                        answered=int(ok.sum()), mean_links=float(nl[ok].mean()) if ok.any() else 0.0,
                        wall_ms=1e3 * wall, search_kernel_ms=st["build_ms"], us_per_search=1e3 * st["build_ms"] / n,
                        searches=st["trees"], expanded_per_search=st["nodes_expanded"] / n,
                        relaxed_per_search=st["edges_relaxed"] / n, improved_per_search=st["improved"] / n,
                        waves_per_search=st["waves"] / n, far_visits_per_search=st["far_visits"] / n,
                        tie_fallbacks=eng.stats().get("tie_fallbacks"), costs_equal_first_config=same)
        except Exception as ex:  # an overflowing configuration is a result too
            line = dict(config=name, options=kv, error=str(ex))
        finally:
            eng.close()
        print(json.dumps(line), flush=True)
        if out_f:
            out_f.write(json.dumps(line) + "\n")
            out_f.flush()
    if out_f:
        out_f.close()


if __name__ == "__main__":
    main()

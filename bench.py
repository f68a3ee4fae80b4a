#!/usr/bin/env python
"""bench.py -- agent-steps/s (and reroutes/s) of the DRIFT simulation step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

One "step" = one block of ``--ticks-per-step`` one-second ticks of the whole hot path (advance,
link events, ordered occupancy + BPR entry speeds, device-driven departures and their routing)
over all agents.  ``value`` = agents x ticks / time with all inputs resident in HBM; ``e2e`` =
the same through the public API with host buffers: every step uploads that step's trip chains
from pinned host memory and reads back arrivals, departures and the link-volume vector.

Workloads (BASELINE.json configs; the city networks of the reference are not shipped, so the
graphs are synthetic):
  regional1m   configs[2]: ~1.16 M-link regional road graph, activity-based demand (the reference's trip law
               per agent type, lib/st_selection.py:926-1050), 1 M agents
  brussels100k configs[1]: ~53 k-node city graph, gravity-structured demand, 100 k agents
  bundled      configs[0]-sized: 100-node graph, random demand, 300 agents (CPU-reference scale)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HOURLY = np.array([0.05, 0.02, 0.01, 0.01, 0.02, 0.05, 0.15, 0.25, 0.35, 0.20, 0.15, 0.18,
                   0.22, 0.18, 0.16, 0.20, 0.25, 0.30, 0.35, 0.25, 0.15, 0.12, 0.08, 0.06])  # config.py:336-340

WORKLOADS = {
    # name: (graph nodes, agents per GPU, demand model, description)
    "regional1m": dict(nodes=400_000, agents=1_000_000, model="activity",
                       desc="configs[2]: regional road graph ~1.16M links, activity-structured demand, 1M agents/GPU"),
    "regional1m_pools": dict(nodes=400_000, agents=1_000_000, model="activity_pools",
                             desc="regional1m with the pool-to-pool variant of the activity demand (every trip has an end "
                                  "in a small pool: the best case for the tree cache)"),
    "brussels100k": dict(nodes=50_000, agents=100_000, model="gravity", reroute=True,
                         desc="configs[1]: city road graph ~53k nodes, gravity-structured demand, 100k agents/GPU, "
                              "congestion-aware reroute of en-route agents every step (5 simulated minutes; extension)"),
    "bundled": dict(nodes=100, agents=300, model="random",
                    desc="configs[0]-sized: 100-node graph, random demand, 300 agents"),
    # cold demand (every trip has its own source and destination): routing is one bounded search per trip
    "grid_zone": dict(graph="grid", nodes=1_000_000, agents=1_000_000, model="zone", lookahead=True, arena_per_agent=768,
                      # no end of a trip is shared by more than a few trips: no tree is worth caching (2 GB of slots for
                      # the odd popular node), the memory goes to resident searches, and a launch takes 64 k queries
                      # (tools/router_sweep.py: half the bucket width = 31 % fewer node expansions, 8 % less time)
                      opts=dict(tree_cache_bytes=float(2 << 30), router_chunk=65536.0, tree_delta_scale=0.5),
                      desc="configs[3] at half linear scale (--nodes 4000000 --agents 10000000 = full size): 4-neighbour "
                           "grid, lengths U[50,300] m, 50 km/h, zone-structured demand, lookahead route planning"),
    "rgg_hub": dict(graph="rgg", nodes=200_000, agents=200_000, model="hub", reroute=True, lookahead=True, tps=30,
                    arena_per_agent=768,
                    desc="configs[4] scaled down: random geometric graph (mean degree 6), hub-and-spoke demand, "
                         "congestion-aware reroute of every en-route agent at every step (30 ticks; extension)"),
}
T_START = 2 * 3600.0  # 08:00 in simulation clock terms (06:00 + 2 h): the morning peak (p = 0.35 * 4 / 3600 per s)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self._stop, self._t = gpu, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_workload(name, agents, seed, rank, nodes=None):
    from drift_b200.synth import grid_graph, rgg_graph, road_graph

    w = WORKLOADS[name]
    kind = w.get("graph", "road")
    if kind == "grid":
        side = int(round((nodes or w["nodes"]) ** 0.5))
        lg = grid_graph(side, side, seed=seed, morton=True)
    elif kind == "rgg":
        lg = rgg_graph(nodes or w["nodes"], seed=seed)
    else:
        lg = road_graph(nodes or w["nodes"], seed=seed, morton=os.environ.get("DRIFT_NO_MORTON") is None)
    rng = np.random.default_rng(seed * 1000 + 17 + rank)
    start = rng.integers(0, lg.V, size=agents).astype(np.int32)
    return lg, start, w["model"], rng, kind


def make_chains(model, rng, lg, start, k):
    from drift_b200 import demand

    return demand.MODELS[model](rng, lg, start, k)


def pinned(arr):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t, t.numpy()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (sequential Python restatement of the reference's step, oracle/port.py)
# ------------------------------------------------------------------------------------------------
def cpu_port_rate(model, seed, budget_s=20.0, agents=10_000, nodes=20_000, ticks=300, reroute=False, as_shipped_s=0.0):
    """Times the oracle port on a bounded sample of the same kind of workload.  1 thread: the
    reference is single-threaded Python (one QThread under the GIL)."""
    import random

    from drift_b200.synth import road_graph
    from oracle.port import PortSim, plain_from_arrays

    lg = road_graph(nodes, seed=seed)
    pg = plain_from_arrays(lg.nodes, lg.link_u, lg.link_v, lg.link_len, lg.link_v0, lg.link_t0, lg.link_cap,
                           lg.fwd_rowptr, lg.fwd_col, lg.fwd_w, lg.fwd_link, lg.fwd_minlen, lg.bwd_rowptr,
                           lg.bwd_col, lg.bwd_w)
    rng = np.random.default_rng(seed)
    start = rng.integers(0, lg.V, size=agents).astype(np.int32)
    off, dst = make_chains(model, rng, lg, start, 4)
    random.seed(seed)
    hourly = {h: float(HOURLY[h]) for h in range(24)}
    sim = PortSim(pg, agents, hourly=hourly, accel=10, hours=48, init_nodes=start, chains=(off, dst))
    sim.t = T_START
    t0 = time.perf_counter()
    done = 0
    n_rer = 0
    if reroute:  # the step starts with the extension's travel-time update + reroute, like the B200 arm
        warm = 0
        while warm < 100:  # put some agents on the road first (not timed)
            sim.tick()
            warm += 1
        t0 = time.perf_counter()
        sim.update_travel_times()
        n_rer = sim.reroute_moving()
    while done < ticks and time.perf_counter() - t0 < budget_s:
        sim.tick()
        done += 1
    dt = time.perf_counter() - t0
    moving = sum(1 for a in sim.agents if a.state == 1)
    # "as shipped" (BASELINE.md 3.2): the reference also builds its visualisation snapshot on every tick
    # (UPDATE_FREQUENCY = 1): the same ticks with the snapshot restated, on a short sample
    shipped = None
    if as_shipped_s > 0:
        t1 = time.perf_counter()
        n2 = 0
        while n2 < ticks and time.perf_counter() - t1 < as_shipped_s:
            sim.tick()
            sim.viz_snapshot()
            n2 += 1
        shipped = agents * n2 / (time.perf_counter() - t1) if n2 else None
    return dict(agent_steps_per_s=agents * done / dt, routes_per_s=(sim.routes_computed + n_rer) / dt, seconds=dt,
                ticks=done, agents=agents, nodes=lg.V, links=lg.L, routes=sim.routes_computed, reroutes=n_rer,
                moving_at_end=moving, as_shipped_agent_steps_per_s=shipped)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    args.ticks_per_step = args.ticks_per_step or w.get("tps", 300)
    vals = []
    r = None
    for _ in range(args.warmup):
        cpu_port_rate(w["model"], args.seed, budget_s=2.0, agents=2000, nodes=5000, ticks=50)
    t_all = time.perf_counter()
    for k in range(args.steps):
        r = cpu_port_rate(w["model"], args.seed + k, budget_s=args.ref_seconds or max(5.0, 120.0 / max(args.steps, 1)), agents=10_000,
                          nodes=20_000, ticks=args.ticks_per_step, reroute=bool(w.get("reroute")) and not args.no_reroute,
                          as_shipped_s=2.0 if k == args.steps - 1 else 0.0)
        vals.append(r["agent_steps_per_s"])
    wall = time.perf_counter() - t_all
    v = float(np.mean(vals))
    sample = ("oracle port (sequential Python restatement of lib/simulation_thread.py + lib/agent.py + NetworkX "
              "bidirectional Dijkstra), %d agents, %d-node / %d-link road graph, %d one-second ticks per step, %s demand, "
              "08:00 peak, COLD start (all agents waiting at the first tick, %d moving at the end: fewer than the B200 arm's "
              "~45 %%, which favours the CPU figure); headless = no per-tick visualisation snapshot, as_shipped_value = with "
              "it (UPDATE_FREQUENCY = 1); the unmodified reference ran 6.1e5 agent-steps/s on g.100.0 in the build container" %
              (r["agents"], r["nodes"], r["links"], r["ticks"], w["model"], r["moving_at_end"]))
    line = dict(
        impl="reference", metric="agent-steps/s", value=v, unit="agent-steps/s", n_gpus=args.gpus, steps=args.steps,
        warmup=args.warmup, ms_per_step=1e3 * wall / max(args.steps, 1), higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f64", data="synthetic",
        config=dict(workload=args.workload, description=w["desc"], ticks_per_step=args.ticks_per_step,
                    reference_sample=dict(agents=r["agents"], nodes=r["nodes"], links=r["links"], ticks_per_step=r["ticks"],
                                          demand=w["model"], moving_agents_at_end=r["moving_at_end"],
                                          note="the reference cannot hold the workload's size (every Agent copies the "
                                               "node list: O(N*V) memory, SURVEY.md H5): rates are compared, measured "
                                               "at this size")),
        cpu_baseline=dict(value=v, unit="agent-steps/s", cores=1, kind="port", sample=sample,
                          routes_per_s=r["routes_per_s"], host_cores=os.cpu_count(),
                          as_shipped_value=r["as_shipped_agent_steps_per_s"]),
        e2e=dict(value=v, unit="agent-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0)
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def progress(msg):
    """Timestamped progress line on stderr (DRIFT_BENCH_VERBOSE=1): where a multi-rank run is, should it stall."""
    if os.environ.get("DRIFT_BENCH_VERBOSE"):
        sys.stderr.write("[bench %s rank %s] %s\n" % (time.strftime("%H:%M:%S"), os.environ.get("RANK", "0"), msg))
        sys.stderr.flush()


def run_b200(args):
    import torch

    if os.environ.get("DRIFT_BENCH_VERBOSE"):  # where is a stalled rank? Python stacks every N seconds
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ.get("DRIFT_BENCH_WATCHDOG", "150")), repeat=True, file=sys.stderr)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from drift_b200 import build as _build

    if rank == 0:
        _build.build()
    if dist is not None:
        dist.barrier()
    from drift_b200.engine import Engine, ROUTER_BATCHED

    w = WORKLOADS[args.workload]
    n_local = args.agents or w["agents"]
    lg, start, model, rng, graph_kind = build_workload(args.workload, n_local, args.seed, rank, nodes=args.nodes)
    K, W, TPS = args.steps, args.warmup, args.ticks_per_step or w.get("tps", 300)
    arena_links = max(1 << 24, n_local * w.get("arena_per_agent", 2560))
    trips_per_step = 2
    delta = 1.0
    sharded = None
    if world > 1:
        from drift_b200.dist import ShardedEngine

        sharded = ShardedEngine(lg, n_local * world, device=local_rank, route_arena_links=arena_links)
        eng = sharded.engine
        if args.exchange == "p2p":
            sharded.enable_p2p()
    else:
        eng = Engine(lg, n_local, device=local_rank, agent_base=0, route_arena_links=arena_links)
        bench_stream = torch.cuda.Stream(device=local_rank)  # the engine launches here, so torch events see its kernels
        eng.set_stream(bench_stream.cuda_stream)
    if sharded is not None:
        bench_stream = torch.cuda.current_stream(local_rank)
    # HBM-resident pass: candidate destinations of all warm-up + timed steps uploaded once
    total_steps = W + K
    # one chain per agent covers both passes (resident + e2e), so that both draw from the same activity
    # pools / hubs and the e2e pass does not pay for trees of a second, unrelated demand pattern
    _, dst_all = make_chains(model, rng, lg, start, 4 + trips_per_step * (total_steps + K + 1))
    dst_all = dst_all.reshape(n_local, -1)
    dst_e2e = dst_all[:, 4 + trips_per_step * total_steps:]
    dst_all = dst_all[:, :4 + trips_per_step * total_steps]
    if w.get("lookahead"):
        eng.set_option("plan_lookahead", 1)
    engine_opts = dict(w.get("opts", {}))
    for kv in args.opt:
        k, v = kv.split("=")
        engine_opts[k] = float(v)
    for k, v in engine_opts.items():
        eng.set_option(k, v)
    # static travel times -> contraction hierarchy (host, cached on disk: rank 0 contracts, the others load) + hub
    # labels built on the device; graphs without a usable hierarchy (grids) keep the tree / bounded-search router
    # (the reroute extension routes on congested times from its first update on: labels of the static times would
    # serve the initial batch only)
    hier_info = None
    if not args.no_labels and graph_kind == "road" and not (w.get("reroute") and not args.no_reroute):
        from drift_b200 import hierarchy
        from drift_b200._capi import DriftError

        try:
            t_h = time.perf_counter()
            if rank == 0:
                hier, cached = hierarchy.build_cached(lg, time_limit_s=args.hier_seconds)
            if dist is not None:
                dist.barrier()
            if rank != 0:
                hier, cached = hierarchy.build_cached(lg, time_limit_s=args.hier_seconds)
            t_h = time.perf_counter() - t_h
            eng.set_hierarchy(hier)
            hier_info = dict(hier.info(), host_seconds=t_h, from_cache=bool(cached), **eng.label_stats(reset=False))
        except DriftError as ex:
            hier_info = dict(unavailable=str(ex))
        progress("hierarchy: %r" % (hier_info,))
    labels_on = hier_info is not None and "unavailable" not in hier_info
    if args.landmarks > 0 and not labels_on:
        eng.set_landmarks(args.landmarks)
    eng.set_demand(args.seed, HOURLY, start, chain_capacity=4, router=ROUTER_BATCHED)
    eng.set_option("route_window", TPS)
    if args.coop_blocks:
        eng.set_option("coop_blocks_per_sm", args.coop_blocks)
    eng.push_trips(dst_all[:, :4])
    pool = np.ascontiguousarray(dst_all[:, 4:].reshape(n_local, total_steps, trips_per_step).transpose(1, 0, 2))
    eng.store_trip_pool(pool)
    eng.sync()
    step_no = [0]

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    t_sim = T_START
    launches0 = 0

    reroute = bool(w.get("reroute")) and not args.no_reroute
    n_reroutes = [0]

    def run_block(resident=True):
        nonlocal t_sim
        if resident:
            eng.refill_from_pool(step_no[0])
            step_no[0] += 1
            eng.discard_logs()  # nothing is read back in the resident pass: free the rings (no copy)
        if reroute and eng.tick > 0:  # K4 + reroute of en-route agents + fresh routes for this step's departures
            if sharded is not None:  # volumes of all ranks: dense all-reduce of the recount, then K4 on every rank
                sharded.update_travel_times()
            else:
                eng.update_travel_times()
            n_reroutes[0] += eng.reroute_moving(ROUTER_BATCHED)
            eng.invalidate_prefetch(t_sim, delta, TPS)
        if sharded is not None:
            t_sim = sharded.run(TPS, delta, t_sim)
        else:
            eng.run(TPS, delta, t_sim)
            for _ in range(TPS):
                t_sim += delta

    progress("setup done, warm-up")
    for _ in range(W):
        run_block()
    barrier()
    progress("timed region")
    launches0 = eng.kernel_times()["total_launches"]
    st0 = eng.stats()
    n_reroutes[0] = 0
    eng.set_option("time_plan", 1)
    eng.phase_times(reset=True)
    if sharded is not None and args.exchange == "p2p":
        eng.p2p_phase_detail(1)
    eng.tree_stats(reset=True)
    eng.label_stats(reset=True)
    with ClockSampler(local_rank) as clk:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(bench_stream)
        for _ in range(K):
            run_block()
        ev1.record(bench_stream)
        barrier()
        dt = ev0.elapsed_time(ev1) * 1e-3  # device time on the engine's stream, host planning gaps included
    st1 = eng.stats()
    phases = eng.phase_times(reset=True)
    p2p_detail = eng.p2p_phase_detail(phases["ticks"]) if (sharded is not None and args.exchange == "p2p") else None
    trees = eng.tree_stats(reset=True)
    lab = eng.label_stats(reset=True)
    eng.set_option("time_plan", 0)
    launches = eng.kernel_times()["total_launches"] - launches0
    if dist is not None:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    total_agents = n_local * world
    value = total_agents * TPS * K / dt
    routes = st1["departures_total"] - st0["departures_total"]
    if dist is not None:  # whole-job count
        rt = torch.tensor([routes], device="cuda", dtype=torch.int64)
        dist.all_reduce(rt)
        routes = int(rt.item())
    reroutes_timed = n_reroutes[0]
    if dist is not None and reroute:
        rt = torch.tensor([reroutes_timed], device="cuda", dtype=torch.int64)
        dist.all_reduce(rt)
        reroutes_timed = int(rt.item())

    progress("timed region done: %.1f ms/step" % (1e3 * dt / K))
    # ---- e2e pass: per step H2D of the step's chains (pinned) + D2H of arrivals/departures/volumes
    e2e_value = dt_e2e = None
    h2d = d2h_total = e2e_fallbacks = 0
    e2e_phases, e2e_trees, st_e1 = {}, {}, eng.stats()
    if not args.skip_e2e:  # --skip-e2e: exploratory runs of the largest configs (the line then carries no e2e number)
        dst_s = np.ascontiguousarray(dst_e2e).reshape(n_local, K + 1, trips_per_step)
        keep_t, blocks_p = zip(*[pinned(dst_s[:, b, :]) for b in range(K + 1)])
        h2d = blocks_p[0].nbytes
        d2h_total = 0
        eng.push_trips(blocks_p[K])
        run_block(False)  # warm the per-step path
        eng.drain_arrivals(), eng.drain_departures(), (sharded.occupancy() if sharded is not None else eng.occupancy())
        barrier()
        st_e0 = eng.stats()
        eng.phase_times(reset=True)
        eng.tree_stats(reset=True)
        eng.set_option("time_plan", 1)
        ev0.record(bench_stream)
        for b in range(K):
            eng.push_trips(blocks_p[b])
            run_block(False)
            a = eng.drain_arrivals()
            d = eng.drain_departures()
            occ = sharded.occupancy() if sharded is not None else eng.occupancy()
            d2h_total += sum(x.nbytes for x in a) + sum(x.nbytes for x in d) + occ.nbytes
        ev1.record(bench_stream)
        barrier()
        dt_e2e = ev0.elapsed_time(ev1) * 1e-3
        e2e_phases = eng.phase_times(reset=True)
        e2e_trees = eng.tree_stats(reset=True)
        eng.set_option("time_plan", 0)
        st_e1 = eng.stats()
        e2e_fallbacks = st_e1["fallback_routes"] - st_e0["fallback_routes"]
        if dist is not None:
            tt = torch.tensor([dt_e2e], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt_e2e = float(tt.item())
        e2e_value = total_agents * TPS * K / dt_e2e

    progress("e2e done")
    # ---- auxiliary link passes on the state the runs left behind: volume recount (K3') and BPR travel-time pass (K4)
    peak, peak_src = peaks()
    st_aux = eng.stats()
    tp = eng.time_passes(20)
    mv_aux, touched = st_aux["moving_agents"], st_aux["links_with_traffic"]
    if sharded is not None:  # with the peer exchange a link's volume lives on its owner: ~1/world of the touched links here
        touched = min(lg.L, mv_aux)
    sc_bytes = 1.0 * n_local + 4.0 * mv_aux + 4.0 * lg.L + 8.0 * touched
    bpr_bytes = 24.0 * lg.L
    roof_scatter = dict(bound="hbm", kernel="zero-fill + k_recount (per-link volume = histogram of cur_link; warp-merged, "
                                            "staged in a shared-memory histogram per 1024-agent tile)",
                        us_per_launch=tp["recount_us"], algorithmic_bytes_per_launch=sc_bytes,
                        achieved=sc_bytes / (tp["recount_us"] * 1e-6) / 1e9, peak=peak, unit="GB/s",
                        frac=sc_bytes / (tp["recount_us"] * 1e-6) / 1e9 / peak,
                        traffic=8.57e6 if (args.workload == "regional1m" and not args.agents and sharded is None) else None,
                        traffic_source="ncu --set full capture of k_recount on this workload "
                                       "(profiles/r2_recount_travel_times_set_full_raw.csv)", peak_source=peak_src,
                        agents=n_local, moving_agents=mv_aux, links=lg.L, links_touched=touched,
                        note="1 B state per agent + 4 B cur_link per moving agent + 4 B zero-fill per link + 8 B RMW per "
                             "touched link (SURVEY.md 8d); %.1f MB at this size = %.1f us at peak: launch-latency-bound"
                             % (sc_bytes / 1e6, sc_bytes / peak / 1e3))
    roof_bpr = dict(bound="hbm", kernel="k_travel_times (tt = t0 / factor(occ), fused elementwise pass over the link arrays)",
                    us_per_launch=tp["travel_times_us"], algorithmic_bytes_per_launch=bpr_bytes,
                    achieved=bpr_bytes / (tp["travel_times_us"] * 1e-6) / 1e9, peak=peak, unit="GB/s",
                    frac=bpr_bytes / (tp["travel_times_us"] * 1e-6) / 1e9 / peak,
                    traffic=18.0e6 if (args.workload == "regional1m" and not args.agents) else None,
                    traffic_source="dram__bytes_read of the ncu --set full capture (the 9.3 MB of tt stores stay in L2 within "
                                   "the capture; profiles/r2_recount_travel_times_set_full_raw.csv)", peak_source=peak_src,
                    links=lg.L, note="24 B per link: occ 4 R + capacity class 4 R + t0 8 R + tt 8 W (SURVEY.md 8d)")
    # ---- multi-GPU: north_star's dense per-tick all-reduce of the volume vector, as the labelled baseline next to
    # the sparse event exchange the tick uses (SURVEY.md H6 "report both")
    dense = None
    if sharded is not None:
        barrier()
        vol = sharded.dense_volume_allreduce()
        same = bool(np.array_equal(vol.cpu().numpy(), sharded.occupancy()))
        iters = 50
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        d0.record(bench_stream)
        for _ in range(iters):
            sharded.dense_volume_allreduce()
        d1.record(bench_stream)
        barrier()
        us = 1e3 * d0.elapsed_time(d1) / iters
        tt = torch.tensor([us], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        us = float(tt.item())
        payload = 4 * lg.L
        dense = dict(us_per_tick=us, payload_bytes=payload, bus_gbs=2.0 * (world - 1) / world * payload / (us * 1e-6) / 1e9,
                     equals_event_exchange_volumes=same,
                     note="recount of this rank's agents + ncclAllReduce(int32[L], sum) per tick (max over ranks); "
                          "volumes only: a sum cannot order one tick's entries across ranks (SURVEY.md 0-5), so the "
                          "tick exchanges link events instead -- compare with tick_only.exchange_tick_us")
    # ---- roofline pass: per-launch CUDA-event timing of the advance kernel on the engine stream
    stm0 = eng.stats()
    eng.set_profiling(True)
    prof_ticks = min(TPS, 100)
    if sharded is not None:
        sharded.run(prof_ticks, delta, t_sim)
    else:
        eng.run(prof_ticks, delta, t_sim)
    kt = eng.kernel_times()
    eng.set_profiling(False)
    stm1 = eng.stats()
    moving = 0.5 * (stm0["moving_agents"] + stm1["moving_agents"])
    waiting = n_local - moving
    adv_ms = kt["advance_ms"] / max(kt["advance_launches"], 1)
    alg_bytes = waiting * 1.0 + moving * 33.0  # 1 B per waiting, 33 B per moving agent-step
    achieved = alg_bytes / (adv_ms * 1e-3) / 1e9 if adv_ms > 0 else 0.0
    roof_adv = dict(bound="hbm", kernel="k_advance (standalone launch, CUDA events on the engine stream)",
                    achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=43.9e6 if (args.workload == "regional1m" and not args.agents) else None,
                    traffic_source="dram__bytes_read+write per launch: 100 launches of this kernel on this workload in the "
                                   "ncu launch list profiles/r2_launches_regional1m_ncu.csv (4.39 GB + 2.5 MB)",
                    peak_source=peak_src, us_per_launch=adv_ms * 1e3, moving_agents=moving,
                    algorithmic_bytes_per_launch=alg_bytes,
                    note="1 B per waiting + 33 B per moving agent-step (SURVEY.md 8d); at this size the kernel is "
                         "latency-bound, see roofline_advance_saturated")
    # the dominant kernel of the timed region: the batched tree builder when routing dominates
    tick_ms = (phases["advance_us_per_tick"] + phases["inject_us_per_tick"] + phases["resolve_us_per_tick"]) * phases["ticks"] / 1e3
    if trees["build_ms"] > tick_ms and trees["launches"] > 0:
        tb = 20.0 * trees["edges_relaxed"] + 12.0 * trees["improved"] + 16.0 * trees["nodes_expanded"]
        per_launch = tb / trees["launches"]
        us = 1e3 * trees["build_ms"] / trees["launches"]
        ach = per_launch / (us * 1e-6) / 1e9
        roof = dict(bound="hbm", kernel="k_build_trees (batched reverse SSSP, CUDA events on the engine stream)",
                    achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=None, peak_source=peak_src,
                    us_per_launch=us, algorithmic_bytes_per_launch=per_launch, launches=trees["launches"],
                    trees=trees["trees"], edges_relaxed=trees["edges_relaxed"], improved=trees["improved"],
                    nodes_expanded=trees["nodes_expanded"], waves=trees["waves"], far_visits=trees["far_visits"],
                    share_of_step=trees["build_ms"] / (1e3 * dt),
                    note="20 B per relaxed edge + 12 B per improving relaxation + 16 B per expanded node (SURVEY.md 8d)")
    else:
        # the tick kernel dominates: roofline of the persistent kernel itself, from its own phase clocks
        # (%globaltimer, grid barriers included) over the timed region
        tick_us = phases["advance_us_per_tick"] + phases["inject_us_per_tick"] + phases["resolve_us_per_tick"]
        ev_tick = float(st1["events_last_tick"])
        mv_t = 0.5 * (st0["moving_agents"] + st1["moving_agents"])
        alg_tick = (n_local - mv_t) * 1.0 + mv_t * 33.0 + ev_tick * 90.0
        ach_t = alg_tick / (tick_us * 1e-6) / 1e9 if tick_us > 0 else 0.0
        ncu_ok = args.workload == "regional1m" and not args.agents and world == 1 and TPS == 300
        roof = dict(bound="hbm", kernel="k_tick_loop (persistent cooperative kernel: advance + resolve phases of "
                                        "%d ticks per launch; in-kernel %%globaltimer phase clocks)" % TPS,
                    achieved=ach_t, peak=peak, unit="GB/s", frac=ach_t / peak,
                    traffic=6.8e9 if ncu_ok else None,
                    traffic_source="dram__bytes_read+write of one 300-tick launch (5.65 / 7.88 GB in the two captured "
                                   "launches = 19-26 MB per tick), ncu --set full capture of this kernel on this workload "
                                   "(profiles/r2_tick_loop_label_query_label_paths_set_full_raw.csv)" if ncu_ok else None,
                    peak_source=peak_src, us_per_launch=tick_us * TPS, us_per_tick=tick_us,
                    algorithmic_bytes_per_launch=alg_tick * TPS, algorithmic_bytes_per_tick=alg_tick,
                    moving_agents=mv_t, events_per_tick=ev_tick, share_of_step=tick_us * TPS / (1e6 * dt / K),
                    note="1 B per waiting + 33 B per moving agent-step + ~90 B per link event (SURVEY.md 8d); at 1 M "
                         "agents a tick is latency-bound between grid barriers (profiles/r1_summary.md): the advance "
                         "phase alone is in roofline_advance (standalone kernel) and roofline_advance_saturated")
    sat = None
    if rank == 0 and world == 1 and args.saturated_agents > 0:
        eng.close()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import advance_roofline

        r = advance_roofline.measure(args.saturated_agents)
        sat = dict(bound="hbm", kernel="advance phase of k_tick_loop / k_advance, all agents moving",
                   agents=r["agents"], achieved=r["persistent_advance_gbs"], peak=peak, unit="GB/s",
                   frac=r["persistent_advance_gbs"] / peak, us_per_tick=r["persistent_advance_us"],
                   standalone_kernel_us=r["k_advance_us"], standalone_kernel_frac=r["k_advance_gbs"] / peak,
                   algorithmic_bytes_per_launch=r["algorithmic_bytes"], peak_source=peak_src)

    # ---- multi-GPU: what a tick costs WITHOUT the exchange (same engine, same agents, peer exchange suspended:
    # timing only, the state is not used afterwards), so that the exchange is visible separately from the
    # embarrassingly parallel route planning
    progress("roofline pass done")
    tick_only = None
    if sharded is not None and args.exchange == "p2p":
        barrier()
        p2p_tick_us = phases["advance_us_per_tick"] + phases["inject_us_per_tick"] + phases["resolve_us_per_tick"]
        eng.set_option("p2p_suspend", 1)
        eng.phase_times(reset=True)
        eng.run(TPS, delta, t_sim)
        eng.sync()
        loc = eng.phase_times(reset=True)
        local_tick_us = loc["advance_us_per_tick"] + loc["inject_us_per_tick"] + loc["resolve_us_per_tick"]
        tt = torch.tensor([local_tick_us, p2p_tick_us], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        local_tick_us, p2p_tick_us = float(tt[0].item()), float(tt[1].item())
        tick_only = dict(local_tick_us=local_tick_us, exchange_tick_us=p2p_tick_us, exchange_sections_us=p2p_detail,
                         tick_only_efficiency=local_tick_us / p2p_tick_us if p2p_tick_us > 0 else None,
                         note="max over ranks; local = the same engine with the peer exchange suspended (k_tick_loop on "
                              "this rank's agents only, timing probe), exchange = k_tick_loop_p2p in the timed region")
        barrier()

    # ---- multi-GPU: N-rank parity self-check (SURVEY.md 8e acceptance), after the timed regions
    progress("tick-only probe done")
    parity = None
    if sharded is not None and not args.no_parity_check:
        sharded.close()
        dist.barrier()
        from drift_b200.selfcheck import sharded_vs_single

        parity = sharded_vs_single(local_rank, agents=args.parity_agents, nodes=20_000, ticks=100, blocks=3,
                                   p2p=(args.exchange == "p2p"))

    progress("parity check done: %r" % (parity,))
    # ---- the drop-in SimulationThread itself (exact mode: host MT19937 draws in the reference's order, the
    # reference's selector interface, device step + NetworkX-exact device router) at the reference's own scale
    thread_rate = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        import tempfile

        from drift_b200.headless import run_headless
        from oracle import golden

        if sat is None:
            eng.close()
        case = golden.GoldenCase("liz_random_nopath")
        th, wall = run_headless(case.nx_graph(), 10_000, "random", hours=0.25, accel=10, seed=0,
                                out_dir=tempfile.mkdtemp())
        thread_rate = dict(value=10_000 * th.step / wall, unit="agent-steps/s", agents=10_000, ticks=th.step,
                           trips=len(th.data_logger.trips_data), seconds=wall,
                           note="drift_b200.SimulationThread.update_simulation_step per tick (headless, no viz snapshot), "
                                "random model, music-net digraph of the reference's example data (1 157 nodes)")
        th.close()
        progress("drop-in thread: %r" % (thread_rate,))
    line = None
    if rank == 0:
        cpu = cpu_port_rate(model, args.seed, budget_s=args.cpu_seconds, ticks=TPS, reroute=reroute,
                            as_shipped_s=min(3.0, 0.2 * args.cpu_seconds)) if args.cpu_seconds > 0 else None
        line = dict(
            metric="agent-steps/s", value=value, unit="agent-steps/s", n_gpus=world, steps=K, warmup=W,
            ms_per_step=1e3 * dt / K, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f64",
            data="synthetic",
            config=dict(workload=args.workload, description=w["desc"], agents_per_gpu=n_local, total_agents=total_agents,
                        nodes=lg.V, links=lg.L, demand=model, ticks_per_step=TPS, tick_seconds=delta,
                        router=("hub labels over a contraction hierarchy of the static travel times (labels built on the "
                                "device, a route = intersection of two sorted labels + shortcut unpacking)" if labels_on else
                                "batched SSSP: cached shortest-path trees on either end of a trip (destination / source), "
                                "bounded goal-directed searches for cold pairs") +
                               "; planning: %s" % ("lookahead" if w.get("lookahead") else "prefetch-all"),
                        hierarchy=hier_info, engine_options=engine_opts or None, clock="08:00 peak",
                        l2="no explicit flush: a tick re-reads the same agent state by construction (%.0f MB hot + ~80 MB "
                           "cold per-agent state at this size, L2 is 126 MB) and every step's route planning walks "
                           "GBs of cached trees and route arena in between; roofline_advance_saturated uses 16 M "
                           "agents (400 MB of hot state, >> L2) for the bandwidth-bound figure" % (n_local * 25 / 1e6),
                        multi_gpu=(("agents sharded by id block, graph replicated, links owned round-robin: events stored "
                                    "into the owner's inbox and entry speeds into the home rank's arrays over NVLink peer "
                                    "memory inside the persistent tick kernel (k_tick_loop_p2p), no collective call"
                                    if args.exchange == "p2p" else
                                    "agents sharded by id block, graph replicated, per-tick all-gather of link events "
                                    "(NCCL), every rank resolves the global event set")) if world > 1 else "n/a"),
            routes_per_s=routes / dt, routes_in_timed_region=int(routes),
            fallback_routes_in_timed_region=int(st1["fallback_routes"] - st0["fallback_routes"]),
            tie_fallbacks_in_timed_region=int(st1["tie_fallbacks"] - st0["tie_fallbacks"]),
            reroutes_per_s=(reroutes_timed / dt) if reroute else None, reroutes_in_timed_region=int(reroutes_timed),
            moving_fraction=float(moving / n_local),
            e2e=dict(value=e2e_value, unit="agent-steps/s", h2d_bytes_per_step=int(h2d),
                     d2h_bytes_per_step=int(d2h_total / K), ms_per_step=(1e3 * dt_e2e / K) if dt_e2e else None,
                     fallback_routes=int(e2e_fallbacks), arena_compactions=int(st_e1["arena_compactions"]),
                     router_stats={k: (round(v, 3) if isinstance(v, float) else v) for k, v in e2e_trees.items()},
                     phases={k: (round(v, 3) if isinstance(v, float) else v) for k, v in e2e_phases.items()}),
            gpu_launches=int(launches), router_stats={k: (round(v, 3) if isinstance(v, float) else v) for k, v in trees.items()},
            roofline=roof, roofline_advance=roof_adv, roofline_advance_saturated=sat,
            roofline_scatter=roof_scatter, roofline_bpr=roof_bpr, dense_allreduce=dense,
            parity_check=parity, tick_only=tick_only, dropin_thread=thread_rate,
            label_router=(dict(queries=lab["queries"], launches=lab["launches"], query_kernel_ms=round(lab["query_ms"], 4),
                               entries_read=lab["entries_read"], algorithmic_bytes=12 * lab["entries_read"],
                               achieved_gbs=(12e-9 * lab["entries_read"] / (lab["query_ms"] * 1e-3)) if lab["query_ms"] > 0 else None,
                               note="k_label_query in the timed region: 12 B (hub id + fp64 distance) per label entry of "
                                    "the two labels a query intersects; CUDA events on the engine stream")
                          if labels_on else None),
            clocks=clk.summary(),
            phases=dict(note="timed region: per-tick phase times of the persistent kernel (%globaltimer, barriers "
                             "included) and total ms in route-planning batches",
                        **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in phases.items()}))
        if cpu is not None:
            line["cpu_baseline"] = dict(
                value=cpu["agent_steps_per_s"], unit="agent-steps/s", cores=1, kind="port",
                routes_per_s=cpu["routes_per_s"], host_cores=os.cpu_count(),
                replicas_upper_bound=cpu["agent_steps_per_s"] * (os.cpu_count() or 1),  # one independent run per core
                as_shipped_value=cpu["as_shipped_agent_steps_per_s"],
                sample="oracle port (sequential Python restatement of the reference step incl. NetworkX-order "
                       "bidirectional Dijkstra), %d agents on a %d-node / %d-link road graph, %d ticks in %.1f s, %s demand, "
                       "cold start (all agents waiting at the first tick, %d moving at the end; the B200 arm runs at ~45 %% "
                       "moving: the comparison favours the CPU); value = headless (no per-tick visualisation snapshot), "
                       "as_shipped_value = with the snapshot of lib/simulation_thread.py:394-416 on every tick"
                       % (cpu["agents"], cpu["nodes"], cpu["links"], cpu["ticks"], cpu["seconds"], model, cpu["moving_at_end"]))
        print(json.dumps(line))
    if sat is None:
        eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: N-rank parity check FAILED: %r\n" % (parity,))
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="regional1m", choices=sorted(WORKLOADS))
    ap.add_argument("--agents", type=int, default=0, help="agents per GPU (default: the workload's)")
    ap.add_argument("--nodes", type=int, default=0, help="graph nodes (default: the workload's)")
    ap.add_argument("--ticks-per-step", type=int, default=0, help="ticks per step (default: the workload's, 300)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg (0 = skip)")
    ap.add_argument("--ref-seconds", type=float, default=0.0,
                    help="--impl reference: CPU seconds per step (default: 120 s spread over the steps, at least 5 s)")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable)")
    ap.add_argument("--landmarks", type=int, default=8, help="ALT landmarks for goal-directed bounded searches (0 = off)")
    ap.add_argument("--no-reroute", action="store_true", help="switch the periodic reroute extension off")
    ap.add_argument("--saturated-agents", type=int, default=16_000_000,
                    help="also time the advance kernel with this many agents all moving (0 = skip)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "allgather"],
                    help="multi-GPU: link events over NVLink peer memory inside the tick kernel, or per-tick NCCL all-gather")
    ap.add_argument("--coop-blocks", type=int, default=0, help="blocks per SM of the persistent tick kernel (0 = default)")
    ap.add_argument("--skip-e2e", action="store_true",
                    help="skip the end-to-end pass (exploratory runs of the largest configs; e2e is then null)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="label of the line: weak = --agents per GPU is fixed as N grows (default), strong = the caller "
                         "divides a fixed total by N in --agents")
    ap.add_argument("--no-labels", action="store_true",
                    help="do not build the contraction hierarchy / hub labels (route with trees + bounded searches)")
    ap.add_argument("--hier-seconds", type=float, default=240.0, help="time limit of the host-side contraction")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="multi-GPU: skip the N-rank parity self-check (sharded run == one engine, bit for bit)")
    ap.add_argument("--parity-agents", type=int, default=200_000, help="agents of the N-rank parity self-check")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())

/*
 * drift_b200 -- C ABI of the B200-native DRIFT simulation step.
 *
 * The reference (jbaudru/DRIFT) is pure Python and has no native boundary; the
 * entry points below are what a ctypes stub inside the reference's
 * `lib/simulation_thread.py` would bind.  Each one cites
 * the reference code it replaces (paths relative to the reference root; `nx:`
 * = NetworkX 3.6.1 `networkx/algorithms/shortest_paths/`).
 *
 * Conventions
 *   - every function returns 0 on success, a negative DRIFT_E* code on error;
 *     `drift_last_error()` returns a human-readable message for the calling thread;
 *   - the caller owns all host buffers, the library owns all device memory;
 *   - one engine per CUDA device / per rank; calls on one engine must be serialised
 *     by the caller (the reference itself is single-threaded);
 *   - work is stream-ordered on the engine's own stream; functions that return data
 *     to host buffers synchronise, `drift_step` does not (see `drift_sync`);
 *   - no exceptions and no C++/torch types cross this boundary.
 *
 * Data model
 *   nodes   0..V-1  in `list(G.nodes)` order
 *   links   0..L-1  in `G.edges(keys=True)` order (one per (u,v,key) edge: the unit
 *           of occupancy, `TrafficManager.edge_agents`, lib/managers/traffic_manager.py:19)
 *   forward CSR  in `G._succ[u]` order, backward CSR in `G._pred[v]` order, one entry per
 *           (u,v) pair carrying min travel_time over parallel edges (nx:weighted.py:76-77)
 *           and, forward only, the link an agent occupies on that hop (first key with
 *           minimal travel_time, lib/agent.py:122-130)
 *   agents  local index 0..N-1, global id = agent_base + index (ordering key across ranks)
 */
#ifndef DRIFT_B200_H
#define DRIFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRIFT_ABI_VERSION 1

#define DRIFT_OK 0
#define DRIFT_EINVAL (-1)    /* bad argument */
#define DRIFT_ECUDA (-2)     /* CUDA runtime error */
#define DRIFT_ENOMEM (-3)    /* device allocation failed */
#define DRIFT_EOVERFLOW (-4) /* a device-side buffer (events, arrivals, routes, heap) overflowed */
#define DRIFT_ESTATE (-5)    /* call sequence error */

/* route status codes (per query) */
#define DRIFT_ROUTE_OK 0
#define DRIFT_ROUTE_TRIVIAL 1 /* source == target: path = [s] (nx:weighted.py:2393-2394) */
#define DRIFT_ROUTE_NOPATH 2  /* NetworkXNoPath / node missing (lib/agent.py:87-99) */

/* routing engines */
#define DRIFT_ROUTER_NX_EXACT 0 /* bit-exact nx bidirectional_dijkstra order (nx:weighted.py:2387-2459) */
#define DRIFT_ROUTER_BATCHED 1  /* batched router: cached shortest-path trees / bounded searches (or hub labels, see
                                  drift_set_hierarchy); a query whose shortest path may not be unique -- a second hop
                                  tight within 1e-12 -- is re-run by the NetworkX-exact router, so the link sequences
                                  equal NetworkX's on tie-prone graphs too (option "detect_ties": automatic up to 20 000 nodes) */

typedef struct drift_engine drift_engine;

typedef struct drift_graph_desc {
  int32_t V;  /* nodes */
  int32_t L;  /* links ((u,v,key) edges) */
  int32_t Ef; /* forward CSR entries ((u,v) pairs) */
  int32_t Eb; /* backward CSR entries (== Ef) */
  const int32_t* fwd_rowptr; /* [V+1] */
  const int32_t* fwd_col;    /* [Ef] successor node */
  const double* fwd_w;       /* [Ef] min travel_time over parallel edges */
  const int32_t* fwd_link;   /* [Ef] link occupied on this hop */
  const int32_t* bwd_rowptr; /* [V+1] */
  const int32_t* bwd_col;    /* [Eb] predecessor node */
  const double* bwd_w;       /* [Eb] */
  const int32_t* bwd_link;   /* [Eb] link occupied on the hop (pred -> node) */
  const double* link_len;    /* [L] metres   (lib/agent.py:146) */
  const double* link_v0;     /* [L] km/h     (lib/agent.py:147) */
  const double* link_t0;     /* [L] seconds  (lib/graph_loader.py:759-766) */
  const int32_t* link_cap;   /* [L]          (lib/managers/traffic_manager.py:33-61) */
  const int32_t* link_u;     /* [L] tail node */
  const int32_t* link_v;     /* [L] head node */
} drift_graph_desc;

typedef struct drift_stats_out {
  int64_t moving_agents;         /* lib/simulation_thread.py:424 */
  int64_t total_agents_on_roads; /* lib/managers/traffic_manager.py:125-128 */
  int64_t total_capacity;        /* :129 */
  int64_t links_with_traffic;    /* :132-135 ('congested_edges') */
  int64_t arrivals_total;        /* completed trips since creation */
  int64_t departures_total;
  int64_t route_arena_used;      /* link ids currently held in the route arena */
  int64_t events_last_tick;
  int64_t arena_compactions;     /* times the live routes were copied into the spare arena */
  int64_t fallback_routes;       /* departures routed inside a tick because no planned route was ready */
  int64_t tie_fallbacks;         /* batched-router queries re-run by the NetworkX-exact router (possible ties) */
} drift_stats_out;

const char* drift_last_error(void);
int drift_abi_version(void);

/* Device-side copy of the graph (replaces the nx.MultiDiGraph walked by lib/agent.py:94,118-151
 * and the TrafficManager tables built at lib/managers/traffic_manager.py:16-61). */
int drift_create(drift_engine** out, int device, const drift_graph_desc* g);
int drift_destroy(drift_engine* e);
int drift_sync(drift_engine* e);

/* BPR congestion factor, lib/managers/traffic_manager.py:76-107.
 * `v_c_ratio ** beta` is libm pow in the reference; to stay bit-exact the host tabulates
 * factor(count) = max(0.1, 1/(1+alpha*(count/cap)**beta)) per capacity class with the same
 * Python float arithmetic; entry 0 is 1.0 and the last entry of a class is the saturated 0.1.
 *   class_off[n_class+1]  start of each class in factor_table
 *   link_class[L]         capacity class of each link */
int drift_set_bpr(drift_engine* e, double alpha, double beta, const double* factor_table,
                  const int64_t* class_off, int32_t n_class, const int32_t* link_class);

/* Agents (lib/simulation_thread.py:181-261 builds N Agent objects; all start 'waiting',
 * lib/agent.py:23).  `agent_base` is the global id of local agent 0. */
int drift_add_agents(drift_engine* e, int64_t n_agents, int64_t agent_base, int64_t route_arena_links);

/* Routing: replaces nx.shortest_path(G, s, t, weight='travel_time') (lib/agent.py:94).
 * Routes Q (src,dst) pairs on the device and stages the link sequences in the route arena.
 *   status[Q]   DRIFT_ROUTE_*
 *   n_links[Q]  hops of each path (0 for TRIVIAL / NOPATH)
 * `use_congested` != 0 routes on the congestion-updated travel times written by
 * drift_update_travel_times (extension; the reference never updates travel_time). */
int drift_route_od(drift_engine* e, int32_t router, int64_t Q, const int32_t* src, const int32_t* dst,
                   int32_t use_congested, int32_t* status, int32_t* n_links);
/* Copy the staged link sequences of the last drift_route_od to the host, concatenated in
 * query order (sum of n_links entries).  The node path of query q is
 * [src] + [link_v[l] for l in links_q]. */
int drift_fetch_routes(drift_engine* e, int32_t* links_out, int64_t cap);
/* Path costs of the last batch (left-to-right fp64 sums of fwd_w), for tolerance checks. */
int drift_fetch_route_costs(drift_engine* e, double* cost_out, int64_t Q);
/* Make the staged routes the pending departures of `agent[q]` for the next tick
 * (Agent.start_new_journey, lib/agent.py:101-104).  TRIVIAL / NOPATH entries leave the agent
 * waiting (lib/agent.py:95-99, 107-116). */
int drift_commit_departures(drift_engine* e, int64_t Q, const int32_t* agent);
/* Departures with host-supplied routes (replay of a recorded trace):
 * route_off[n+1] into route_links; an empty route = TRIVIAL. */
int drift_submit_departures(drift_engine* e, int64_t n, const int32_t* agent, const int64_t* route_off,
                            const int32_t* route_links);

/* The tick: SimulationThread.update_simulation_step / _update_agent_batch
 * (lib/simulation_thread.py:288-340) -> Agent.update (lib/agent.py:171-203) ->
 * setup_current_edge (:106-160) -> TrafficManager register/unregister/get_congested_speed
 * (lib/managers/traffic_manager.py:63-107), for n_ticks ticks of `delta` seconds
 * (delta = dt * time_acceleration, lib/simulation_thread.py:294).  Pending departures are
 * applied in the first tick.  Asynchronous. */
int drift_step(drift_engine* e, int32_t n_ticks, double delta);

/* Device-driven demand (throughput mode): every waiting agent departs at tick k
 * with probability hourly[hour(t_k)]*4*delta/3600 (lib/agent.py:176-180, 243-259) drawn from a
 * counter-based generator keyed (seed, global agent id, tick).  Its trips go from its current
 * node (lib/agent.py:47: source = previous target) to the destinations the host pushes into a
 * per-agent ring of `chain_capacity` upcoming targets (generated by the s-t model on the host).
 * Routes of the next two trips of every agent are prefetched as one large batch every
 * "route_window" ticks (option; the reference's travel_time is static, so a route does not
 * depend on when it is computed); `router` selects the engine.  start_node[N] = initial node of
 * each agent (lib/agent.py:40-44). */
int drift_set_demand(drift_engine* e, uint64_t seed, const double hourly24[24], const int32_t* start_node,
                     int32_t chain_capacity, int32_t router);
/* Top up every agent's ring from `k` candidate destinations per agent (row-major [N][k] host
 * buffer, copied asynchronously: keep it alive until the next synchronising call).  The k
 * candidates of an agent are one unit (e.g. an out leg and its return leg): taken together when the
 * ring has room for all k, otherwise the row is dropped. */
int drift_push_trips(drift_engine* e, const int32_t* candidates, int32_t k);
/* Same, from a pool kept in HBM: n_blocks blocks of [N][k] candidates uploaded once. */
int drift_store_trip_pool(drift_engine* e, const int32_t* candidates, int32_t n_blocks, int32_t k);
int drift_refill_from_pool(drift_engine* e, int32_t block);
/* Run n_ticks ticks with device-driven demand starting at simulation time t0 (seconds). */
int drift_run(drift_engine* e, int32_t n_ticks, double delta, double t0);

/* Extension (no reference behaviour): congestion-aware travel times
 * tt[l] = t0[l] / factor(occ[l]) (formula of lib/managers/traffic_manager.py:92-107 applied
 * per link) and rerouting of en-route agents from the head of their current link. */
int drift_update_travel_times(drift_engine* e);
/* Same from a caller-supplied device vector of link volumes (int32[L]): with agents sharded over several GPUs the
 * volumes of all ranks are summed first (ncclAllReduce of the recount, drift_recount_device) -- north_star's "the
 * per-link volume vector is all-reduced" feeding the BPR update -- so that every rank routes on the same times. */
int drift_update_travel_times_with(drift_engine* e, const void* volumes_dev);
int drift_reroute_moving(drift_engine* e, int32_t router, int64_t* n_rerouted);
/* Device-driven demand + extension: waiting agents that will depart during the next n_ticks ticks
 * (clock t0, tick length delta) drop their prefetched routes; the next drift_run replans them on the
 * current travel times. */
int drift_invalidate_prefetch(drift_engine* e, double t0, double delta, int32_t n_ticks);

/* Readback.  Any output pointer may be NULL. */
int drift_read_occupancy(drift_engine* e, int32_t* occ_out /* [L] */);
int drift_recount_occupancy(drift_engine* e, int32_t* occ_out /* [L] */); /* histogram of cur_link (traffic_manager.py:122-135) */
/* The same recount left on the device (asynchronous; buffer "volume_recount" of drift_device_ptr): the payload of
 * the dense per-tick ncclAllReduce(int32[L]) mode of drift_b200/dist.py (north_star's literal design). */
int drift_recount_device(drift_engine* e);
int drift_read_travel_times(drift_engine* e, double* tt_out /* [L] */);
int drift_read_agents(drift_engine* e, uint8_t* state, double* dist_on_link, double* speed, double* link_len,
                      int32_t* cur_link, int32_t* route_idx, int32_t* route_len, int32_t* trip_count,
                      int32_t* cur_node);
int drift_read_agent_route(drift_engine* e, int64_t agent, int32_t* links_out, int64_t cap, int64_t* n);
/* Trip-record export (lib/data_logger.py:55-120): hops, length in metres (sum of link lengths) and,
 * when links_out != NULL, the link ids (concatenated) of the current route of each listed agent. */
int drift_export_routes(drift_engine* e, int64_t n, const int32_t* agents, int32_t* n_links, double* metres,
                        int32_t* links_out, int64_t cap);
int drift_drain_arrivals(drift_engine* e, int32_t* agent, int32_t* tick, int64_t cap, int64_t* n);
int drift_drain_departures(drift_engine* e, int32_t* agent, int32_t* tick, int32_t* src, int32_t* dst,
                           int32_t* status, int64_t cap, int64_t* n);
/* Same log with the arena position of each departure's route (offset, hops, compaction count at departure), and
 * the export of such routes: metres (left-to-right fp64 sum of the link lengths, lib/data_logger.py:103-120) and,
 * when links_out != NULL, the link ids concatenated in order (sum of route_len entries).  A logged position is
 * valid until the next arena compaction (DRIFT_ESTATE afterwards): drain and export after every block of at most
 * "route_window" ticks.  This is what keeps trip_distance_km / route_taken exact for agents that depart twice in a
 * block or arrive before the block ends (drift_b200/block_runner.py). */
int drift_drain_departures_routes(drift_engine* e, int32_t* agent, int32_t* tick, int32_t* src, int32_t* dst,
                                  int32_t* status, int64_t* route_off, int32_t* route_len, int32_t* route_epoch,
                                  int64_t cap, int64_t* n);
int drift_export_logged_routes(drift_engine* e, int64_t n, const int64_t* route_off, const int32_t* route_len,
                               const int32_t* route_epoch, double* metres, int32_t* links_out, int64_t cap);
/* optional link-entry trace (tick, agent, link, speed) for parity tests */
int drift_enable_entry_trace(drift_engine* e, int64_t cap);
int drift_drain_entries(drift_engine* e, int32_t* tick, int32_t* agent, int32_t* link, double* speed,
                        int64_t cap, int64_t* n);
int drift_stats(drift_engine* e, drift_stats_out* out);
int drift_tick_index(drift_engine* e, int64_t* tick);

/* Multi-GPU (agents sharded by contiguous id block, graph replicated).
 * A tick is split around the exchange of link events:
 *   drift_tick_begin    advance + departures -> local events (device), returns their count
 *   drift_events_ptr    device pointer / capacity of the local event buffer (16 B records)
 *   drift_tick_finish   resolve `n_events` gathered records at `events_dev` (device pointer,
 *                       all ranks' records in any order) */
int drift_tick_begin(drift_engine* e, double delta, int64_t* n_events);
int drift_tick_begin_demand(drift_engine* e, double delta, double t, int64_t* n_events); /* same, device-driven demand at clock t */
int drift_events_ptr(drift_engine* e, void** dev_ptr, int64_t* capacity);
int drift_tick_finish(drift_engine* e, const void* events_dev, int64_t n_events);

/* The same exchange without host round trips (what drift_b200/dist.py uses): the engine runs on the
 * caller's CUDA stream (drift_set_stream, e.g. torch's current stream), drift_tick_begin_async
 * leaves [header, events...] in `send_dev` (cap + 1 records of 16 B, header.link = count), the caller
 * all-gathers the send buffers of all ranks into `recv_dev` ([world][cap + 1] records) on that stream,
 * and drift_tick_finish_gathered resolves them.  `t` is the clock after the tick (device-driven
 * demand only).  drift_max_events reports the largest per-tick count seen, for sizing `cap`. */
int drift_set_stream(drift_engine* e, void* cuda_stream);
int drift_tick_begin_async(drift_engine* e, double delta, double t, void* send_dev, int32_t cap);
int drift_tick_finish_gathered(drift_engine* e, void* recv_dev, int32_t world, int32_t cap);
int drift_max_events(drift_engine* e, int64_t* max_events, int32_t reset);

/* The exchange inside the tick kernel, over NVLink peer memory (one process per GPU on one node; what
 * drift_b200/dist.py uses for device-driven demand).  Links are owned round-robin (link % world): every
 * rank stores the events of its agents straight into the owner's inbox, the owner orders the events of
 * its links (lib/agent.py:106-160 + lib/managers/traffic_manager.py:63-107 for those links) and stores
 * the entry speed straight into the home rank's agent arrays; the ranks meet twice per tick through
 * flags in peer memory.  Work per rank does not grow with `world`; the link volumes are held by the
 * owner only (drift_read_occupancy returns 0 for links owned elsewhere: sum over ranks).
 *   drift_p2p_export   allocates inbox / flags and writes 4 CUDA IPC handles (4 x 64 bytes) to
 *                      handles_out; `cap` = events one rank may send to one owner per tick
 *   drift_p2p_connect  all_handles = the world x 4 handles of all ranks in rank order (exchanged by
 *                      the caller, e.g. torch.distributed.all_gather); from then on drift_run launches
 *                      k_tick_loop_p2p.  All ranks must call drift_run with the same arguments. */
int drift_p2p_export(drift_engine* e, int32_t rank, int32_t world, int64_t n_total, int32_t cap, void* handles_out);
int drift_p2p_connect(drift_engine* e, const void* all_handles);

/* Timing of the kernels launched by the last drift_step/drift_run (CUDA events on the engine
 * stream): total milliseconds spent in the advance kernel and number of advance launches. */
int drift_kernel_times(drift_engine* e, double* advance_ms, int64_t* advance_launches, double* total_ms,
                       int64_t* total_launches);
int drift_set_profiling(drift_engine* e, int32_t on);
/* Phase breakdown of the persistent tick kernel since the last reset: %globaltimer nanoseconds spent
 * in advance / inject / resolve (barriers included) over `ticks` ticks, and milliseconds spent in the
 * route-planning batches (CUDA events; needs option "time_plan" = 1). */
int drift_phase_times(drift_engine* e, double* advance_ns, double* inject_ns, double* resolve_ns, double* plan_ms,
                      int64_t* ticks, int32_t reset);
/* Peer-exchange tick kernel since the last reset: nanoseconds on block 0's clock spent in advance | block fence |
 * grid barrier | flag publish | wait for the peers' flags | list building | grid barrier | resolve + fence
 * (divide by the ticks of drift_phase_times). */
int drift_p2p_phase_detail(drift_engine* e, double ns_out[8], int32_t reset);
/* Work and time of the batched tree builder since the last reset (needs option "time_plan" = 1 for
 * the CUDA-event time): trees built, edges relaxed, relaxations that improved a label, nodes expanded. */
int drift_tree_stats(drift_engine* e, double* build_ms, int64_t* launches, int64_t* jobs, int64_t* relaxed,
                     int64_t* improved, int64_t* expanded, int64_t* iters, int64_t* far_visits, int32_t reset);
/* Per-launch time (CUDA events on the engine stream, average of `iters` back-to-back launches after a warm-up) of
 * the two auxiliary link passes on the engine's current state: the per-link volume recount (zero-fill +
 * histogram of cur_link, traffic_manager.py:122-135) and the BPR travel-time pass (traffic_manager.py:92-107
 * applied to every link).  For the roofline report of bench.py. */
int drift_time_passes(drift_engine* e, int32_t iters, double* recount_us, double* travel_times_us);
/* Landmark distances for goal-directed bounded searches in the batched router (ALT): K landmarks,
 * dist_from[K][V] = free-flow travel time landmark -> node, dist_to[K][V] = node -> landmark (+inf where
 * unreachable).  Only used while routing on the static travel times.  K = 0 switches it off. */
int drift_set_landmarks(drift_engine* e, int32_t K, const double* dist_from, const double* dist_to);
/* Contraction hierarchy + hub labels for the STATIC travel times (the reference never updates `travel_time`,
 * lib/graph_loader.py:759-766, so the graph can be preprocessed once): replaces the search behind
 * nx.shortest_path (lib/agent.py:94) by label lookups.  drift_hier_build runs on the host (no device is touched)
 * from the forward CSR of drift_graph_desc; drift_hier_export copies the result into caller-owned arrays (sizes
 * from drift_hier_info; e.g. to cache it on disk); drift_set_hierarchy uploads a hierarchy and builds the hub
 * labels of every node on the device.  From then on DRIFT_ROUTER_BATCHED answers every query on the static
 * travel times from the labels (queries on congested times keep the tree / bounded-search router).  Paths are
 * shortest paths; they equal NetworkX's wherever the shortest path is unique (ties within ~1e-13 relative may
 * resolve differently: shortcut weights are fp64 sums in contraction order).  Fails with DRIFT_EOVERFLOW /
 * DRIFT_ENOMEM / DRIFT_ESTATE (time limit) on graphs without a usable hierarchy (e.g. uniform grids): the
 * engine then keeps its other routers. */
typedef struct drift_hierarchy drift_hierarchy;
typedef struct drift_hier_desc {
  int32_t V;
  int32_t max_level;          /* arcs lead from a lower to a strictly higher level */
  int32_t max_label;          /* largest label seen on a sample of nodes */
  int32_t pad0;
  double mean_label;          /* mean label size (hubs per node and side) on that sample */
  int64_t n_up, n_down;
  const int32_t* level;       /* [V] */
  const int32_t* up_rowptr;   /* [V+1] arcs v -> w with w contracted later, stored at v */
  const int32_t* up_other;    /* [n_up] w */
  const double* up_w;
  const int32_t* up_a;        /* original link id when up_b < 0, else position (down side) of the first half */
  const int32_t* up_b;        /* -1, or position (up side) of the second half */
  const int32_t* up_hops;     /* original links below the arc */
  const int32_t* down_rowptr; /* [V+1] arcs u -> v with u contracted later, stored at v */
  const int32_t* down_other;  /* [n_down] u */
  const double* down_w;
  const int32_t* down_a;
  const int32_t* down_b;
  const int32_t* down_hops;
} drift_hier_desc;
int drift_hier_build(int32_t V, int32_t Ef, const int32_t* fwd_rowptr, const int32_t* fwd_col, const double* fwd_w,
                     const int32_t* fwd_link, int32_t witness_settled, double time_limit_s, drift_hierarchy** out);
int drift_hier_info(const drift_hierarchy* h, int64_t* n_up, int64_t* n_down, int64_t* shortcuts, int32_t* max_level,
                    int32_t* max_label, double* mean_label, double* build_seconds);
int drift_hier_export(const drift_hierarchy* h, drift_hier_desc* d);
int drift_hier_free(drift_hierarchy* h);
int drift_set_hierarchy(drift_engine* e, const drift_hier_desc* d);
/* Hub-label router since the last reset: label entries held, device time of the label build, time (needs option
 * "time_plan") / launches / queries of the query kernel and label entries it read. */
int drift_label_stats(drift_engine* e, int64_t* entries, double* build_ms, double* query_ms, int64_t* launches,
                      int64_t* queries, int64_t* entries_read, int32_t reset);
/* Host-side helper of the vectorised s-t selectors (drift_b200/fast_selection.py; replaces the Python loops
 * of lib/st_selection.py:327-363 and :408-423): out[i] = pow(x[i], y) through the C library, i.e. exactly
 * what CPython's `x ** y` computes.  No device is touched. */
int drift_host_pow(const double* x, double y, double* out, int64_t n);
/* Options (name, value):
 *   route_window        ticks between two route-planning batches (default 300)
 *   plan_lookahead      1 = plan only the departures the coming window's draws allow (default 0 = keep the
 *                       routes of every agent's next two trips ready)
 *   plan_arrival_hops   lookahead: a moving agent with at most this many links to go may arrive within the
 *                       window (-1 = window * fastest link / shortest link)
 *   persistent          0 = one launch per phase and tick instead of the persistent tick kernel
 *   coop_blocks_per_sm  blocks per SM of the persistent tick kernel (default 6)
 *   tree_cache_bytes    HBM budget of the tree cache (default min(48 GB, 30 % of the free memory))
 *   tree_scratch_bytes  HBM budget of the search scratch, 36 B per node and resident search (default: what full
 *                       residency needs, at most 85 % of the free memory less the tree cache)
 *   tree_sides          1 = trees towards destinations only, 2 = also out of sources (default)
 *   bounded_max         a destination wanted by at most this many queries of a batch is answered by a bounded
 *                       search that keeps no tree (0..8; default 8, or 0 when every tree fits the cache)
 *   tree_pop_threshold  a destination asked for more often than this over ~64 batches gets a cached tree anyway
 *   tree_delta_scale    multiplies the near-far bucket width; tree_threads: 64 / 128 / 256 threads per search
 *   arena_compact_frac  the live routes are compacted once this fraction of the route arena is used (0.75)
 *   time_plan           1 = time the planning batches with CUDA events (drift_phase_times / drift_tree_stats)
 *   one_wave_tiles      0 = plain 1024-agent tiles in the persistent tick kernel even when a slightly larger tile
 *                       (1024 + up to 256 agents) would let all tiles run in one wave of resident blocks (default 1)
 *   router_chunk        queries per launch of the batched tree router (default 0 = as many as are guaranteed a cache slot
 *                       should every key be hot; cold demand -- few trips share an end -- can take far larger chunks,
 *                       an overflow of the cache is reported as DRIFT_EOVERFLOW)
 *   detect_ties         1 = the batched tree router re-runs possible ties through the NetworkX-exact router, 0 = it
 *                       returns A shortest path; default -1 = on for graphs of at most 20 000 nodes (the reference's
 *                       scale, where ties are the rule and an exact query is cheap), off above (an exact query on a
 *                       200 k-node graph takes ~0.2 s on its one thread)
 *   tree_wide_regs      1 / 0 = force the 64-register (no spills, at most 4 searches of 256 threads per SM) / the
 *                       32-register build of the search kernel; default -1 = 64, 40 or 32 registers, whichever the
 *                       number of resident searches the scratch budget allows still fits
 *   pot_cache           0 = goal-directed searches fetch the landmark line of a node at every relaxation instead of
 *                       caching the node's potential in its search record (default 1; timing comparisons)
 *   congested_landmarks 1 = keep the free-flow landmark potentials when routing on congested travel times (valid
 *                       lower bounds, no measured gain; default 0)
 *   use_labels          0 = ignore an uploaded hierarchy (DRIFT_ROUTER_BATCHED falls back to trees / bounded searches)
 *   debug_mask          timing probes only, results are then NOT the reference's: 1 = no departure draws,
 *                       2 = crossers stay on their link, 4 = events are not emitted (tools/tick_probe.py)
 * The tree_* budgets, tree_sides and tree_threads must be set before the first batched route. */
int drift_set_option(drift_engine* e, const char* name, double value);
/* Raw device pointers for PyTorch interop (tensor views of engine state). name in
 * {"occ","volume_recount","state","dist","speed","link_len","cur_link","events","travel_time"}. */
int drift_device_ptr(drift_engine* e, const char* name, void** ptr, int64_t* n_elems);

#ifdef __cplusplus
}
#endif
#endif /* DRIFT_B200_H */

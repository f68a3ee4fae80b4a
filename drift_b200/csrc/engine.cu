// drift_b200 -- engine + C ABI (include/drift_b200.h).  sm_100a only; no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "hierarchy.hpp"
#include "label_kernels.cuh"
#include "route_kernels.cuh"
#include "tick_kernels.cuh"

using namespace drift;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(_e == cudaErrorMemoryAllocation ? DRIFT_ENOMEM : DRIFT_ECUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(_e), __FILE__, __LINE__);                                     \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  int64_t n = 0;
  cudaError_t alloc(int64_t count, bool zero = true) {
    release();
    n = count;
    if (count <= 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, sizeof(T) * (size_t)count);
    if (e != cudaSuccess) {
      p = nullptr;
      n = 0;
      return e;
    }
    if (zero) e = cudaMemset(p, 0, sizeof(T) * (size_t)count);
    return e;
  }
  cudaError_t upload(const T* h, int64_t count) {
    cudaError_t e = alloc(count, false);
    if (e != cudaSuccess || count <= 0) return e;
    return cudaMemcpy(p, h, sizeof(T) * (size_t)count, cudaMemcpyHostToDevice);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

inline bool has_duplicates(const int32_t* a, int64_t n) {
  std::vector<int32_t> v(a, a + n);
  std::sort(v.begin(), v.end());
  return std::adjacent_find(v.begin(), v.end()) != v.end();
}

inline int grid_for(int64_t n, int threads, int max_blocks) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace

struct drift_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  // graph
  int32_t V = 0, L = 0, Ef = 0, Eb = 0;
  DevBuf<int32_t> fwd_rowptr, fwd_col, fwd_link, bwd_rowptr, bwd_col, bwd_link, link_cap, link_u, link_v, link_cls;
  DevBuf<double> fwd_w, bwd_w, link_len, link_v0, link_t0, link_tt, factor, fwd_wc, bwd_wc;
  DevBuf<int64_t> cls_off;
  DevBuf<LinkRec> link_rec;
  DevBuf<int32_t> occ, hist;
  DevBuf<unsigned long long> head;
  int64_t total_capacity = 0;
  double alpha = 0.15, beta = 4.0;
  bool have_bpr = false;
  // agents
  int64_t N = 0, Npad = 0;
  uint32_t agent_base = 0;
  DevBuf<uint8_t> state;
  DevBuf<double> dist, speed, elen;
  DevBuf<int32_t> cur_link, next_link, route_idx, route_len, trip_count, cur_node, chain_dst, next_len, next_status;
  DevBuf<uint8_t> chain_head, chain_cnt;
  DevBuf<int64_t> next_off;
  int32_t chain_cap = 4;
  int32_t route_window = 300;
  int32_t plan_lookahead = 0;   // 1 = plan routes only for departures that can happen in the coming window
  int32_t plan_arrival_hops = -1;  // moving agents with at most this many links left may arrive in the window (-1 = from vmax / min length)
  double vmax_kph = 0.0, min_len = 0.0;
  DevBuf<unsigned long long> plan_thresh;
  DevBuf<int32_t> trip_stage, trip_pool;
  int64_t pool_blocks = 0, pool_k = 0;
  DevBuf<int64_t> route_off;
  DevBuf<int32_t> arena, arena2;  // live arena + spare used by the compaction
  int64_t arena_cap = 0;
  int64_t compactions = 0;
  double arena_compact_frac = 0.75;  // live routes are compacted once this fraction of the arena is used
  // events
  DevBuf<Event> events, send;  // kEvSub sub-buffers of ev_sub_cap records; contiguous copy for the rank exchange
  int64_t ev_cap = 0, ev_sub_cap = 0;
  DevBuf<int32_t> send_n;
  // in-tick fallback router scratch (device-driven demand)
  DevBuf<double> f_seen;
  DevBuf<int32_t> f_plink;
  DevBuf<uint32_t> f_mark, f_epoch;
  DevBuf<HeapItem> f_heap;
  int32_t f_workers = 0;
  int64_t f_heap_cap = 0;
  // persistent tick loop
  bool persistent = true;
  int coop_blocks_per_sm = 6;
  int occ_tick_loop = 0, occ_tick_loop_p2p = 0;  // resident blocks per SM of the two persistent kernels (queried once)
  bool one_wave = true;  // option "one_wave_tiles"
  int num_sms = kSMs;
  // pending departures + staged route batch
  DevBuf<int32_t> pend_agent, pend_len;
  DevBuf<int64_t> pend_off;
  int64_t pend_cap = 0;
  DevBuf<int32_t> q_src, q_dst, q_agent, q_status, q_len;
  DevBuf<int64_t> q_off;
  DevBuf<double> q_cost;
  int64_t q_cap = 0;
  int64_t last_Q = 0;
  unsigned long long last_arena_begin = 0, last_arena_end = 0;
  // router scratch
  DevBuf<double> r_seen;
  DevBuf<int32_t> r_plink;
  DevBuf<uint32_t> r_mark, r_epoch;
  DevBuf<HeapItem> r_heap;
  int32_t r_workers = 0;
  int64_t r_heap_cap = 0;
  // tree cache + tree builder scratch (batched router)
  DevBuf<int32_t> t_slot_of, t_job_dst, t_job_slot, t_ctl;  // t_ctl: n_slots, n_jobs, job_cursor
  DevBuf<int2> t_next;  // cached trees: {link, next node} per node
  DevBuf<NodeRec> t_rec;   // per-block search scratch
  int32_t t_sides = 2;       // 2 = trees out of hot sources as well as into hot destinations
  DevBuf<int32_t> t_cnt, t_pop;
  int32_t t_pop_thresh = 16;  // destinations asked for more often than this over ~64 batches get a cached tree
  DevBuf<uint8_t> t_q_side;
  DevBuf<uint32_t> t_iter, t_slot_stamp;
  DevBuf<int32_t> t_slot_owner, t_slot_job, t_job_qhead, t_job_cnt, t_q_next, t_tie_list;
  // option "detect_ties": possible ties of the batched router go to the NetworkX-exact one.  -1 = automatic: on for
  // graphs of at most kTieAutoNodes nodes (the reference's scale: the bundled graphs tie on 56-98 % of the queries and
  // an exact query costs microseconds), off above (one exact query on a 200 k-node graph takes ~0.2 s on one thread:
  // 16 flagged queries cost the rgg_hub workload 1.7 s per step; ties are then measure-zero accidents of the
  // synthetic weights, e.g. links clipped to the same minimum length)
  int32_t detect_ties = -1;
  DevBuf<uint32_t> t_slot_job_batch;
  int32_t tree_wide_regs = -1;  // option "tree_wide_regs": -1 automatic, see launch_router
  int32_t pot_cache = 1;  // option "pot_cache": goal-directed searches keep a node's potential in its record
  int32_t congested_landmarks = 0;  // option "congested_landmarks": experiment, see launch_router
  int64_t router_chunk_opt = 0;  // option "router_chunk": queries per batched-router launch (0 = from the cache slots)
  int32_t t_bounded_max = -1;  // -1 = automatic: bounded search only when the cache cannot hold every destination
  int32_t t_threads = 256;
  int64_t tree_scratch_bytes = (int64_t)64 << 30;
  bool tree_budget_set = false, tree_scratch_set = false;
  uint32_t t_batch = 0;
  DevBuf<int32_t> t_lists;
  DevBuf<double> lm_tab;  // [V][2K] landmark distances interleaved per node
  int32_t lm_K = 0;
  int32_t t_max_slots = 0, t_blocks = 0;
  double t_delta = 1.0;
  int64_t tree_budget_bytes = (int64_t)48 << 30;
  // contraction hierarchy + hub labels (static travel times)
  DevBuf<int32_t> h_level, h_order;
  DevBuf<int32_t> hu_rowptr, hu_other, hu_a, hu_b, hu_hops, hd_rowptr, hd_other, hd_a, hd_b, hd_hops;
  DevBuf<double> hu_w, hd_w;
  DevBuf<long long> lab_off;        // [2][V]
  DevBuf<int32_t> lab_cnt;          // [2][V]
  DevBuf<int32_t> lab_hub, lab_par, lab_flag;
  DevBuf<uint32_t> lab_pnext;
  DevBuf<uint2> q_ent;
  DevBuf<int4> hu_rec, hd_rec;
  DevBuf<double> lab_dist;
  DevBuf<unsigned long long> lab_cursor;
  bool labels_ready = false;
  bool labels_enabled = true;   // option "use_labels"
  int64_t label_entries = 0, label_cap = 0;
  double label_build_ms = 0, label_ms = 0;
  int64_t label_launches = 0, label_queries_host = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> label_timers;
  // rings
  DevBuf<int32_t> arr_agent, arr_tick;
  int64_t arr_cap = 0;
  unsigned long long arr_drained = 0;
  DevBuf<int32_t> ent_tick, ent_agent, ent_link;
  DevBuf<double> ent_speed;
  int64_t ent_cap = 0;
  unsigned long long ent_drained = 0;
  DevBuf<int32_t> dl_agent, dl_tick, dl_src, dl_dst, dl_status, dl_len, dl_epoch;
  DevBuf<long long> dl_off;
  int64_t dl_cap = 0;
  unsigned long long dl_drained = 0;
  // demand
  bool have_demand = false;
  uint64_t seed = 0;
  double hourly[24];
  DevBuf<int32_t> dep_req;
  int32_t demand_router = DRIFT_ROUTER_NX_EXACT;
  // peer-memory exchange (k_tick_loop_p2p)
  int32_t debug_mask = 0;
  bool p2p_on = false;
  P2PArgs p2p;
  DevBuf<Event> p2p_inbox;
  DevBuf<unsigned long long> p2p_flags;
  DevBuf<int32_t> p2p_out_cnt;
  std::vector<void*> p2p_opened;
  // counters
  DevBuf<Counters> ctr;
  Counters* h_ctr = nullptr;  // pinned
  int64_t tick = 0;
  double last_delta = 1.0;
  double plan_ms = 0, tree_ms = 0;
  int64_t tree_launches = 0;
  int64_t loop_ticks = 0;
  bool time_plan = false;
  bool use_congested = false;
  int32_t ticks_since_plan = 0;
  int64_t routes_planned = 0, reroutes_total = 0;
  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tree_timers;
  std::vector<std::pair<int, int>> adv_marks;  // (begin,end) event indices of advance launches
  int tick_ev0 = -1, tick_ev1 = -1;
  double adv_ms = 0, total_ms = 0;
  int64_t adv_launches = 0, total_launches = 0;

  AgentArrays agents() {
    AgentArrays a;
    a.state = state.p;
    a.dist = dist.p;
    a.speed = speed.p;
    a.elen = elen.p;
    a.cur_link = cur_link.p;
    a.next_link = next_link.p;
    a.route_idx = route_idx.p;
    a.route_len = route_len.p;
    a.route_off = route_off.p;
    a.trip_count = trip_count.p;
    a.cur_node = cur_node.p;
    a.chain_dst = chain_dst.p;
    a.chain_head = chain_head.p;
    a.chain_cnt = chain_cnt.p;
    a.chain_cap = chain_cap;
    a.next_off = next_off.p;
    a.next_len = next_len.p;
    a.next_status = next_status.p;
    a.npad = Npad;
    a.n = N;
    a.base = agent_base;
    return a;
  }
  LinkArrays links() {
    LinkArrays l;
    l.len = link_len.p;
    l.v0 = link_v0.p;
    l.t0 = link_t0.p;
    l.cap = link_cap.p;
    l.u = link_u.p;
    l.v = link_v.p;
    l.occ = occ.p;
    l.head = head.p;
    l.tt = link_tt.p;
    l.cls = link_cls.p;
    l.factor = factor.p;
    l.cls_off = cls_off.p;
    l.rec = link_rec.p;
    l.L = L;
    return l;
  }
  GraphArrays graph() {
    GraphArrays g;
    g.fwd_rowptr = fwd_rowptr.p;
    g.fwd_col = fwd_col.p;
    g.fwd_w = fwd_w.p;
    g.fwd_link = fwd_link.p;
    g.bwd_rowptr = bwd_rowptr.p;
    g.bwd_col = bwd_col.p;
    g.bwd_w = bwd_w.p;
    g.bwd_link = bwd_link.p;
    g.link_u = link_u.p;
    g.link_v = link_v.p;
    g.V = V;
    g.L = L;
    return g;
  }
  HierSide hier_up() { return HierSide{hu_rowptr.p, hu_other.p, hu_w.p, hu_a.p, hu_b.p, hu_hops.p, hu_rec.p}; }
  HierSide hier_down() { return HierSide{hd_rowptr.p, hd_other.p, hd_w.p, hd_a.p, hd_b.p, hd_hops.p, hd_rec.p}; }
  LabelSet labels(int side) {
    return LabelSet{lab_off.p + (int64_t)side * V, lab_cnt.p + (int64_t)side * V, lab_hub.p, lab_par.p, lab_pnext.p,
                    lab_dist.p};
  }
  Rings rings() {
    Rings r;
    r.arr_agent = arr_agent.p;
    r.arr_tick = arr_tick.p;
    r.arr_cap = (unsigned long long)arr_cap;
    r.arr_drained = arr_drained;
    r.ent_tick = ent_tick.p;
    r.ent_agent = ent_agent.p;
    r.ent_link = ent_link.p;
    r.ent_speed = ent_speed.p;
    r.ent_cap = (unsigned long long)ent_cap;
    r.ent_drained = ent_drained;
    return r;
  }
};

namespace {

constexpr double kKphToMps = 1000.0 / 3600.0;  // config.py:186
constexpr int32_t kTieAutoNodes = 20000;       // "detect_ties" automatic: on up to this many nodes

int check_overflow(drift_engine* e, const Counters& c) {
  if (!c.overflow) return DRIFT_OK;
  std::string what;
  if (c.overflow & kOvfEvents) what += " events";
  if (c.overflow & kOvfArrivals) what += " arrivals";
  if (c.overflow & kOvfArena) what += " route-arena";
  if (c.overflow & kOvfHeap) what += " router-heap";
  if (c.overflow & kOvfDepLog) what += " departure-log";
  if (c.overflow & kOvfEntries) what += " entry-trace";
  if (c.overflow & kOvfStage) what += " tree-cache";
  if (c.overflow & kOvfExchange) what += " event-exchange (raise the exchange capacity)";
  if (c.overflow & kOvfPending) what += " pending-departures (more departures committed than agents)";
  return fail(DRIFT_EOVERFLOW, "device buffer overflow:%s", what.c_str());
}

int fetch_counters(drift_engine* e, Counters* out) {
  CU(cudaMemcpyAsync(e->h_ctr, e->ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  *out = *e->h_ctr;
  return DRIFT_OK;
}

// Copy the live routes into the spare arena when more than three quarters of the arena are used.  Pending
// (committed, not yet injected) departures and a staged drift_route_od batch must not exist: called at
// the start of a routing batch only.
int maybe_compact_arena(drift_engine* e, unsigned long long cursor) {
  if ((double)cursor < e->arena_compact_frac * (double)e->arena_cap) return DRIFT_OK;
  if (!e->arena2.p) CU(e->arena2.alloc(e->arena_cap, false));
  Counters* c = e->ctr.p;
  CU(cudaMemsetAsync(&c->arena_cursor, 0, sizeof(unsigned long long), e->stream));
  k_compact_routes<<<grid_for(e->N * 32, 256, kSMs * 8), 256, 0, e->stream>>>(
      e->agents(), e->arena.p, e->arena2.p, &c->arena_cursor, (unsigned long long)e->arena_cap, c);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));
  std::swap(e->arena.p, e->arena2.p);
  e->compactions += 1;
  e->last_Q = 0;  // offsets of a staged batch are stale now
  return DRIFT_OK;
}

int ensure_router(drift_engine* e, int64_t Q) {
  // one worker (thread) per concurrent query; scratch is 16 B/node/direction/worker
  int64_t want = std::min<int64_t>(std::max<int64_t>(Q, 1), 148 * 64);
  const int64_t budget = (int64_t)6 << 30;
  const int64_t per_worker = 2 * (int64_t)e->V * 16 + 2 * (int64_t)(e->Ef + 2) * (int64_t)sizeof(HeapItem);
  want = std::max<int64_t>(1, std::min<int64_t>(want, budget / std::max<int64_t>(per_worker, 1)));
  if (e->r_workers >= want) return DRIFT_OK;
  const int64_t W = std::max<int64_t>(want, std::min<int64_t>(2 * (int64_t)e->r_workers, budget / per_worker));
  CU(e->r_seen.alloc(W * 2 * e->V, false));
  CU(e->r_plink.alloc(W * 2 * e->V, false));
  CU(e->r_mark.alloc(W * 2 * e->V, true));
  CU(e->r_epoch.alloc(W, true));
  e->r_heap_cap = (int64_t)e->Ef + 2;
  CU(e->r_heap.alloc(W * 2 * e->r_heap_cap, false));
  e->r_workers = (int32_t)W;
  return DRIFT_OK;
}

int ensure_qcap(drift_engine* e, int64_t Q) {
  if (Q <= e->q_cap) return DRIFT_OK;
  int64_t cap = std::max<int64_t>(Q, std::max<int64_t>(1024, 2 * e->q_cap));
  CU(e->q_src.alloc(cap));
  CU(e->q_dst.alloc(cap));
  CU(e->q_agent.alloc(cap));
  CU(e->q_status.alloc(cap));
  CU(e->q_len.alloc(cap));
  CU(e->q_off.alloc(cap));
  CU(e->q_cost.alloc(cap));
  e->q_cap = cap;
  return DRIFT_OK;
}

int ensure_trees(drift_engine* e) {
  if (e->t_max_slots) return DRIFT_OK;
  const int64_t V = e->V;
  const int64_t search_bytes = 36 * V;  // a 16-byte record and five 4-byte list slots per node
  {  // budgets follow the free HBM unless set by option: up to 48 GB of cached trees; the search scratch takes what
     // full residency needs, up to 85 % of the free memory less the cache (on a 4 M-node graph a search owns 144 MB
     // and the throughput follows the number of resident searches: 148 / 296 / 444 / 591 / 854 searches took
     // 102 / 57 / 43 / 36 / 31 us per search, profiles/r2_router_sweep.md)
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (!e->tree_budget_set) e->tree_budget_bytes = std::min<int64_t>((int64_t)48 << 30, (int64_t)(free_b * 0.30));
    if (!e->tree_scratch_set) {
      const int64_t full = (int64_t)e->num_sms * std::min(32, 2048 / e->t_threads) * search_bytes;
      const int64_t room = (int64_t)(free_b * 0.85) - std::min<int64_t>(e->tree_budget_bytes, e->t_sides * V * 8 * V);
      e->tree_scratch_bytes = std::max<int64_t>(std::min(full, room), (int64_t)(free_b * 0.25));
    }
  }
  int64_t slots = std::min<int64_t>(e->t_sides * V, e->tree_budget_bytes / (8 * V));
  slots = std::max<int64_t>(slots, 1);
  int64_t blocks = std::min<int64_t>((int64_t)e->num_sms * std::min(32, 2048 / e->t_threads), e->tree_scratch_bytes / search_bytes);
  blocks = std::max<int64_t>(blocks, 1);
  CU(e->t_slot_of.alloc(2 * V, false));  // keys: destination t, or V + source s
  CU(cudaMemset(e->t_slot_of.p, 0xff, sizeof(int32_t) * 2 * V));
  CU(e->t_cnt.alloc(2 * V, true));
  CU(e->t_pop.alloc(V, true));
  CU(e->t_next.alloc(slots * V, false));
  CU(e->t_slot_owner.alloc(slots, false));
  CU(cudaMemset(e->t_slot_owner.p, 0xff, sizeof(int32_t) * slots));
  CU(e->t_slot_stamp.alloc(slots, true));
  CU(e->t_slot_job.alloc(slots));
  CU(e->t_slot_job_batch.alloc(slots, true));
  CU(e->t_job_qhead.alloc(slots));
  CU(e->t_job_cnt.alloc(slots));
  CU(e->t_job_dst.alloc(slots));
  CU(e->t_job_slot.alloc(slots));
  CU(e->t_ctl.alloc(4));
  CU(e->t_rec.alloc(blocks * V, false));
  k_init_noderecs<<<kSMs * 8, 256, 0, e->stream>>>(e->t_rec.p, blocks * V);  // jobs leave it at +inf / 0
  CU(cudaGetLastError());
  CU(e->t_iter.alloc(blocks, true));
  CU(e->t_lists.alloc(blocks * 5 * V, false));
  e->t_max_slots = (int32_t)slots;
  e->t_blocks = (int32_t)blocks;
  return DRIFT_OK;
}

int invalidate_trees(drift_engine* e) {
  if (!e->t_max_slots) return DRIFT_OK;
  CU(cudaMemsetAsync(e->t_slot_of.p, 0xff, sizeof(int32_t) * 2 * e->V, e->stream));
  CU(cudaMemsetAsync(e->t_slot_owner.p, 0xff, sizeof(int32_t) * e->t_max_slots, e->stream));
  CU(cudaMemsetAsync(e->t_ctl.p, 0, 4 * sizeof(int32_t), e->stream));
  return DRIFT_OK;
}

// Largest batch the tree cache can take in one go: every hot key needs a slot, and a key is hot only
// if more than bounded_max queries of the batch share it.
int64_t router_chunk(drift_engine* e, int64_t Q, bool congested) {
  if (e->router_chunk_opt > 0) return e->router_chunk_opt;  // the caller knows its demand is cold (few shared ends)
  const int32_t bm = e->t_bounded_max >= 0 ? e->t_bounded_max : ((e->t_max_slots >= e->V && !congested) ? 0 : kMaxBounded);
  if (bm > 0) return std::max<int64_t>(1, (int64_t)e->t_max_slots * (bm + 1));
  return e->t_max_slots >= (int64_t)e->t_sides * e->V ? std::max<int64_t>(Q, 1) : std::max<int64_t>(1, e->t_max_slots);
}

// launches the router over the staged q_src/q_dst (device).  Q is the host-known batch size or an
// upper bound; when n_dev != nullptr the kernels read the actual size from device memory.
int launch_router(drift_engine* e, int32_t router, int64_t Q, int32_t use_congested, const int32_t* n_dev = nullptr,
                  int64_t q0 = 0) {
  if (Q <= 0) return DRIFT_OK;
  RouteOut out;
  out.status = e->q_status.p + q0;
  out.n_links = e->q_len.p + q0;
  out.off = e->q_off.p + q0;
  out.cost = e->q_cost.p + q0;
  const int32_t* qs = e->q_src.p + q0;
  const int32_t* qd = e->q_dst.p + q0;
  const double* wf = use_congested ? e->fwd_wc.p : e->fwd_w.p;
  const double* wb = use_congested ? e->bwd_wc.p : e->bwd_w.p;
  const double* lc = use_congested ? e->link_tt.p : e->link_t0.p;
  if (router == DRIFT_ROUTER_BATCHED && e->labels_ready && e->labels_enabled && !use_congested) {
    // hub labels: every query is the intersection of two sorted labels + a walk through the parent arcs
    if (e->q_ent.n < e->q_cap) {
      CU(cudaStreamSynchronize(e->stream));
      CU(e->q_ent.alloc(e->q_cap, false));
    }
    cudaEvent_t tb0 = nullptr, tb1 = nullptr;
    if (e->time_plan) {
      cudaEventCreate(&tb0);
      cudaEventCreate(&tb1);
      cudaEventRecord(tb0, e->stream);
    }
    const int gw = grid_for(Q * 32, 256, kSMs * 8);
    k_label_query<<<gw, 256, 0, e->stream>>>(Q, n_dev, qs, qd, e->labels(0), e->labels(1), e->V, e->q_ent.p + q0, out,
                                             &e->ctr.p->label_entries);
    if (tb0) {
      cudaEventRecord(tb1, e->stream);
      e->label_timers.push_back({tb0, tb1});
    }
    k_label_paths<<<grid_for(Q * 32, kPathWarps * 32, kSMs * 16), kPathWarps * 32, 0, e->stream>>>(
        Q, n_dev, e->labels(0), e->labels(1), e->hier_up(), e->hier_down(), e->q_ent.p + q0, e->arena.p,
        (unsigned long long)e->arena_cap, e->ctr.p, out, lc);
    CU(cudaGetLastError());
    e->total_launches += 2;
    e->label_queries_host += Q;
    return DRIFT_OK;
  }
  if (router == DRIFT_ROUTER_BATCHED) {
    int rc = ensure_trees(e);
    if (rc) return rc;
    TreeCache tc;
    tc.slot_of = e->t_slot_of.p;
    tc.next_link = e->t_next.p;
    tc.ring_cursor = e->t_ctl.p + 0;
    tc.slot_owner = e->t_slot_owner.p;
    tc.slot_stamp = e->t_slot_stamp.p;
    tc.batch = ++e->t_batch;
    if (e->t_q_next.n < Q + q0) {
      CU(cudaStreamSynchronize(e->stream));
      const int64_t need = std::max<int64_t>(e->q_cap, Q + q0);
      CU(e->t_q_next.alloc(need, false));
      CU(e->t_q_side.alloc(need, false));
      CU(e->t_tie_list.alloc(need, false));
      // per-batch job table: at most one job per query (bounded searches take no cache slot)
      CU(e->t_job_qhead.alloc(need + e->t_max_slots));
      CU(e->t_job_cnt.alloc(need + e->t_max_slots));
      CU(e->t_job_dst.alloc(need + e->t_max_slots));
      CU(e->t_job_slot.alloc(need + e->t_max_slots));
    }
    tc.cnt = e->t_cnt.p;
    tc.pop = e->t_pop.p;
    tc.pop_thresh = e->t_pop_thresh;
    if ((tc.batch & 63u) == 0) k_pop_decay<<<kSMs * 4, 256, 0, e->stream>>>(e->t_pop.p, e->V);
    tc.q_side = e->t_q_side.p;
    tc.V = e->V;
    tc.slot_job = e->t_slot_job.p;
    tc.slot_job_batch = e->t_slot_job_batch.p;
    tc.job_qhead = e->t_job_qhead.p;
    tc.job_cnt = e->t_job_cnt.p;
    tc.q_next = e->t_q_next.p;  // indexed by the query number inside this call
    // automatic: full trees only pay off when they can all stay cached and the weights do not change
    tc.bounded_max = e->t_bounded_max >= 0 ? e->t_bounded_max
                                           : ((e->t_max_slots >= e->V && !use_congested) ? 0 : kMaxBounded);
    // with full trees only, every key of a batch needs a slot: source-side keys only if all 2V fit
    tc.sides = (tc.bounded_max == 0 && e->t_max_slots < 2 * e->V) ? 1 : e->t_sides;
    tc.detect_ties = e->detect_ties < 0 ? (e->V <= kTieAutoNodes ? 1 : 0) : e->detect_ties;
    CU(cudaMemsetAsync(out.status, 0xff, sizeof(int32_t) * Q, e->stream));  // -1 = not answered yet
    tc.n_jobs = e->t_ctl.p + 1;
    tc.job_cursor = e->t_ctl.p + 2;
    tc.max_slots = e->t_max_slots;
    tc.job_dst = e->t_job_dst.p;
    tc.job_slot = e->t_job_slot.p;
    TreeScratch ts;
    ts.rec = e->t_rec.p;
    ts.lists = e->t_lists.p;
    ts.iter = e->t_iter.p;
    CU(cudaMemsetAsync(e->t_ctl.p + 1, 0, 3 * sizeof(int32_t), e->stream));
    const int gq = grid_for(Q, 256, kSMs * 4);
    k_tree_count<<<gq, 256, 0, e->stream>>>(Q, qs, qd, tc, n_dev);
    k_tree_hits<<<gq, 256, 0, e->stream>>>(Q, qs, qd, tc, n_dev);
    k_tree_requests<<<gq, 256, 0, e->stream>>>(Q, qs, qd, tc, e->ctr.p, n_dev);
    Landmarks lm;
    lm.tab = e->lm_tab.p;
    // The landmark distances are free-flow distances: valid lower bounds on congested times too (tt = t0 / factor >=
    // t0).  Measured on the reroute workloads they change nothing (rgg_hub 105 vs 109 ms per step), so congested
    // routing runs without them unless option "congested_landmarks" asks.
    lm.K = (use_congested && !e->congested_landmarks) ? 0 : e->lm_K;
    k_link_queries<<<gq, 256, 0, e->stream>>>(Q, qs, qd, tc, n_dev);
    cudaEvent_t tb0 = nullptr, tb1 = nullptr;
    if (e->time_plan) {
      cudaEventCreate(&tb0);
      cudaEventCreate(&tb1);
      cudaEventRecord(tb0, e->stream);
    }
    // registers follow the residency the scratch budget allows: at most 4 / 6 / 8 searches of 256 threads per SM ->
    // the 64-register (no spills) / 40-register / 32-register build
    const int64_t per_sm_threads = ((int64_t)e->t_blocks * e->t_threads + e->num_sms - 1) / e->num_sms;
    const int regs_class = e->tree_wide_regs == 1 ? 4 : e->tree_wide_regs == 0 ? 8
                           : per_sm_threads <= 1024 ? 4 : per_sm_threads <= 1536 ? 6 : 8;
#define DRIFT_LAUNCH_TREES(BPS)                                                                                         \
  k_build_trees<BPS, 1><<<e->t_blocks, e->t_threads, 0, e->stream>>>(e->graph(), wb, wf, e->t_delta, tc, ts, qs,          \
                                                                     e->arena.p, (unsigned long long)e->arena_cap,        \
                                                                     e->ctr.p, out, lc, lm, e->pot_cache)
    if (regs_class == 4)
      DRIFT_LAUNCH_TREES(4);
    else if (regs_class == 6)
      DRIFT_LAUNCH_TREES(6);
    else
      DRIFT_LAUNCH_TREES(kTreeBlocksPerSM);
#undef DRIFT_LAUNCH_TREES
    if (tb0) {
      cudaEventRecord(tb1, e->stream);
      e->tree_timers.push_back({tb0, tb1});
    }
    k_extract_paths<<<gq, 256, 0, e->stream>>>(e->graph(), Q, n_dev, qs, qd, tc, e->arena.p,
                                                (unsigned long long)e->arena_cap, e->ctr.p, out, lc);
    CU(cudaGetLastError());
    e->total_launches += 6;
    if (tc.detect_ties) {
      // queries whose shortest path may not be unique (exactly or within rounding) go through the NetworkX-exact
      // router, which reproduces the reference's tie-breaking (nx:weighted.py:2387-2459); rare on continuous weights
      k_collect_ties<<<gq, 256, 0, e->stream>>>(Q, n_dev, out.status, e->t_tie_list.p, e->t_ctl.p + 3, e->ctr.p);
      int32_t n_ties = 0;
      CU(cudaMemcpyAsync(&n_ties, e->t_ctl.p + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
      e->total_launches += 1;
      if (n_ties > 0) {
        const int64_t per_worker = 2 * (int64_t)e->V * 16 + 2 * (int64_t)(e->Ef + 2) * (int64_t)sizeof(HeapItem);
        const int64_t want = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_ties, 256), ((int64_t)2 << 30) / per_worker));
        rc = ensure_router(e, want);
        if (rc) return rc;
        RouteScratch sc;
        sc.seen = e->r_seen.p;
        sc.plink = e->r_plink.p;
        sc.mark = e->r_mark.p;
        sc.heap = e->r_heap.p;
        sc.epoch = e->r_epoch.p;
        sc.heap_cap = e->r_heap_cap;
        sc.workers = (int32_t)std::min<int64_t>(e->r_workers, n_ties);
        k_route_nx<<<(sc.workers + 31) / 32, 32, 0, e->stream>>>(e->graph(), wf, wb, n_ties, qs, qd, sc, e->arena.p,
                                                                 (unsigned long long)e->arena_cap, e->ctr.p, out, lc,
                                                                 nullptr, e->t_tie_list.p);
        CU(cudaGetLastError());
        e->total_launches += 1;
      }
    }
    return DRIFT_OK;
  }
  int rc = ensure_router(e, Q);
  if (rc) return rc;
  RouteScratch sc;
  sc.seen = e->r_seen.p;
  sc.plink = e->r_plink.p;
  sc.mark = e->r_mark.p;
  sc.heap = e->r_heap.p;
  sc.epoch = e->r_epoch.p;
  sc.heap_cap = e->r_heap_cap;
  sc.workers = (int32_t)std::min<int64_t>(e->r_workers, Q);
  const int blocks = (sc.workers + 31) / 32;
  k_route_nx<<<blocks, 32, 0, e->stream>>>(e->graph(), wf, wb, Q, qs, qd, sc, e->arena.p,
                                           (unsigned long long)e->arena_cap, e->ctr.p, out, lc, n_dev);
  CU(cudaGetLastError());
  e->total_launches += 1;
  return DRIFT_OK;
}

struct ProfScope {
  drift_engine* e;
  int i0 = -1;
  bool adv;
  ProfScope(drift_engine* e_, bool adv_) : e(e_), adv(adv_) {
    if (!e->profiling) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    i0 = (int)e->ev_pool.size();
    e->ev_pool.push_back(a);
    e->ev_pool.push_back(b);
    cudaEventRecord(a, e->stream);
  }
  ~ProfScope() {
    if (i0 < 0) return;
    cudaEventRecord(e->ev_pool[i0 + 1], e->stream);
    if (adv) e->adv_marks.push_back({i0, i0 + 1});
  }
};

TickArgs make_tick_args(drift_engine* e, double delta, bool demand) {
  TickArgs p;
  memset(&p, 0, sizeof(p));
  p.a = e->agents();
  p.lk = e->links();
  p.g = e->graph();
  p.rg = e->rings();
  if (demand) {
    p.dm.seed = e->seed;
    p.dm.dep_req = e->dep_req.p;
    p.dm.log_agent = e->dl_agent.p;
    p.dm.log_tick = e->dl_tick.p;
    p.dm.log_src = e->dl_src.p;
    p.dm.log_dst = e->dl_dst.p;
    p.dm.log_status = e->dl_status.p;
    p.dm.log_off = e->dl_off.p;
    p.dm.log_len = e->dl_len.p;
    p.dm.log_epoch = e->dl_epoch.p;
    p.dm.epoch = (int32_t)e->compactions;
    p.dm.log_cap = (unsigned long long)e->dl_cap;
    p.dm.log_drained = e->dl_drained;
    p.fsc.seen = e->f_seen.p;
    p.fsc.plink = e->f_plink.p;
    p.fsc.mark = e->f_mark.p;
    p.fsc.heap = e->f_heap.p;
    p.fsc.epoch = e->f_epoch.p;
    p.fsc.heap_cap = e->f_heap_cap;
    p.fsc.workers = e->f_workers;
  }
  p.w_fwd = e->use_congested ? e->fwd_wc.p : e->fwd_w.p;
  p.w_bwd = e->use_congested ? e->bwd_wc.p : e->bwd_w.p;
  p.link_cost = e->use_congested ? e->link_tt.p : e->link_t0.p;
  p.arena = e->arena.p;
  p.arena_cap = (unsigned long long)e->arena_cap;
  p.ctr = e->ctr.p;
  p.events = e->events.p;
  p.ev_cap = (int32_t)e->ev_sub_cap;
  p.pend_agent = e->pend_agent.p;
  p.pend_off = e->pend_off.p;
  p.pend_len = e->pend_len.p;
  p.pend_cap = (int32_t)e->pend_cap;
  p.delta = delta;
  p.kph_to_mps = kKphToMps;
  p.demand = demand ? 1 : 0;
  p.debug = e->debug_mask;
  return p;
}

// one tick as separate kernels, phases A + A2: advance, departures -> local events
int tick_begin(drift_engine* e, double delta, bool demand, double p_dep) {
  e->tick += 1;
  const int32_t tick = (int32_t)e->tick;
  const TickArgs p = make_tick_args(e, delta, demand);
  const int64_t per_block = (int64_t)kAdvThreads * kAdvPerThread;
  const int blocks = grid_for(e->N, (int)per_block, e->num_sms * 8);
  {
    ProfScope ps(e, true);
    if (demand)
      k_advance<true><<<blocks, kAdvThreads, 0, e->stream>>>(p, tick, p_dep);
    else
      k_advance<false><<<blocks, kAdvThreads, 0, e->stream>>>(p, tick, p_dep);
  }
  k_inject<<<std::max(1, std::min(e->num_sms, (int)((e->pend_cap + 255) / 256))), 256, 0, e->stream>>>(p, tick);
  CU(cudaGetLastError());
  e->total_launches += 2;
  e->adv_launches += 1;
  return DRIFT_OK;
}

// phase C as a separate kernel.  gathered = events of all ranks (re-linked first); n_events < 0 = local
int tick_finish(drift_engine* e, Event* events, int64_t n_events, bool gathered, double delta) {
  const TickArgs p = make_tick_args(e, delta, e->have_demand);
  const int32_t tick = (int32_t)e->tick;
  const int64_t hint = n_events < 0 ? 2 * e->N : n_events;
  const int g = grid_for(std::max<int64_t>(hint, 1), 256, e->num_sms * 4);
  if (gathered && n_events > 0) {
    k_link_events<<<g, 256, 0, e->stream>>>(e->links(), events, (int)n_events, tick);
    e->total_launches += 1;
  }
  k_resolve<<<g, 256, 0, e->stream>>>(p, events, (int)n_events, tick);
  k_reset_counters<<<1, 32, 0, e->stream>>>(e->ctr.p, tick);
  CU(cudaGetLastError());
  e->total_launches += 2;
  return DRIFT_OK;
}

// n_ticks ticks in one cooperative launch (k_tick_loop): p_dep and the window plan are constant inside
int run_tick_loop(drift_engine* e, int32_t n_ticks, double delta, bool demand, double p_dep) {
  if (n_ticks <= 0) return DRIFT_OK;
  TickArgs p = make_tick_args(e, delta, demand);
  int32_t tick0 = (int32_t)e->tick + 1;
  if (!e->occ_tick_loop) {  // resident blocks per SM of the two persistent kernels (queried once per engine)
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->occ_tick_loop, k_tick_loop, kAdvThreads, 0));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->occ_tick_loop_p2p, k_tick_loop_p2p, kAdvThreads, 0));
    if (e->occ_tick_loop < 1 || e->occ_tick_loop_p2p < 1) return fail(DRIFT_ECUDA, "k_tick_loop does not fit on an SM");
  }
  const int per_sm = std::min(e->p2p_on ? e->occ_tick_loop_p2p : e->occ_tick_loop, e->coop_blocks_per_sm);
  int blocks = e->num_sms * per_sm;
  // every tile of a phase in ONE wave where a few more agents per tile achieve it (see TickArgs::tile_singles)
  int64_t tiles = (e->N + kAdvTileQuads - 1) / kAdvTileQuads;
  p.tile_singles = 0;
  if (tiles > blocks && e->one_wave)
    for (int s = 32; s <= kAdvMaxSingles; s += 32)
      if ((e->N + kAdvTileQuads + s - 1) / (kAdvTileQuads + s) <= blocks) {
        p.tile_singles = s;
        tiles = (e->N + kAdvTileQuads + s - 1) / (kAdvTileQuads + s);
        break;
      }
  if (tiles < blocks) blocks = (int)std::max<int64_t>(tiles, 1);
  if (e->p2p_on) {
    void* xargs[] = {(void*)&p, (void*)&e->p2p, (void*)&tick0, (void*)&n_ticks, (void*)&p_dep};
    CU(cudaLaunchCooperativeKernel((void*)k_tick_loop_p2p, dim3(blocks), dim3(kAdvThreads), xargs, 0, e->stream));
  } else {
    void* args[] = {(void*)&p, (void*)&tick0, (void*)&n_ticks, (void*)&p_dep};
    CU(cudaLaunchCooperativeKernel((void*)k_tick_loop, dim3(blocks), dim3(kAdvThreads), args, 0, e->stream));
  }
  e->tick += n_ticks;
  e->loop_ticks += n_ticks;
  e->total_launches += 1;
  e->adv_launches += n_ticks;
  return DRIFT_OK;
}

}  // namespace

extern "C" {

const char* drift_last_error(void) { return g_err.c_str(); }
int drift_abi_version(void) { return DRIFT_ABI_VERSION; }

int drift_create(drift_engine** out, int device, const drift_graph_desc* g) {
  if (!out || !g) return fail(DRIFT_EINVAL, "drift_create: null argument");
  if (g->V <= 0 || g->L < 0 || g->Ef < 0 || g->Eb != g->Ef) return fail(DRIFT_EINVAL, "drift_create: bad sizes");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(DRIFT_ECUDA, "drift_create: no CUDA device (%s); this library has no CPU path",
                cudaGetErrorString(ce));
  CU(cudaSetDevice(device));
  drift_engine* e = new drift_engine();
  e->device = device;
  *out = nullptr;
#define CUX(call)                  \
  do {                             \
    cudaError_t _e2 = (call);      \
    if (_e2 != cudaSuccess) {      \
      delete e;                    \
      CU(_e2);                     \
    }                              \
  } while (0)
  CUX(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  e->V = g->V;
  e->L = g->L;
  e->Ef = g->Ef;
  e->Eb = g->Eb;
  CUX(e->fwd_rowptr.upload(g->fwd_rowptr, g->V + 1));
  CUX(e->fwd_col.upload(g->fwd_col, g->Ef));
  CUX(e->fwd_w.upload(g->fwd_w, g->Ef));
  CUX(e->fwd_link.upload(g->fwd_link, g->Ef));
  CUX(e->bwd_rowptr.upload(g->bwd_rowptr, g->V + 1));
  CUX(e->bwd_col.upload(g->bwd_col, g->Eb));
  CUX(e->bwd_w.upload(g->bwd_w, g->Eb));
  CUX(e->bwd_link.upload(g->bwd_link, g->Eb));
  CUX(e->link_len.upload(g->link_len, g->L));
  CUX(e->link_v0.upload(g->link_v0, g->L));
  for (int32_t l = 0; l < g->L; ++l) {
    e->vmax_kph = std::max(e->vmax_kph, g->link_v0[l]);
    e->min_len = l == 0 ? g->link_len[l] : std::min(e->min_len, g->link_len[l]);
  }
  CUX(e->link_t0.upload(g->link_t0, g->L));
  CUX(e->link_tt.upload(g->link_t0, g->L));
  CUX(e->link_cap.upload(g->link_cap, g->L));
  CUX(e->link_u.upload(g->link_u, g->L));
  CUX(e->link_v.upload(g->link_v, g->L));
  CUX(e->fwd_wc.upload(g->fwd_w, g->Ef));
  CUX(e->bwd_wc.upload(g->bwd_w, g->Eb));
  CUX(e->occ.alloc(std::max(1, g->L)));
  CUX(e->head.alloc(std::max(1, g->L)));
  {
    cudaDeviceProp prop;
    CUX(cudaGetDeviceProperties(&prop, device));
    e->num_sms = prop.multiProcessorCount;
  }
  CUX(e->hist.alloc(std::max(1, g->L)));
  CUX(e->ctr.alloc(1));
  CUX(cudaMallocHost(&e->h_ctr, sizeof(Counters)));
  e->total_capacity = 0;
  for (int32_t l = 0; l < g->L; ++l) e->total_capacity += g->link_cap[l];
  {  // bucket width of the near-far tree builder: a few average hops
    double sum = 0;
    for (int32_t k = 0; k < g->Ef; ++k) sum += g->fwd_w[k];
    e->t_delta = g->Ef ? 8.0 * sum / g->Ef : 1.0;
    if (!(e->t_delta > 0)) e->t_delta = 1.0;
  }
  for (int h = 0; h < 24; ++h) e->hourly[h] = 0.05;
#undef CUX
  *out = e;
  return DRIFT_OK;
}

int drift_destroy(drift_engine* e) {
  if (!e) return DRIFT_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void* q : e->p2p_opened) cudaIpcCloseMemHandle(q);
  e->p2p_opened.clear();
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  if (e->h_ctr) cudaFreeHost(e->h_ctr);
  if (e->stream && e->own_stream) cudaStreamDestroy(e->stream);
  delete e;
  return DRIFT_OK;
}

int drift_sync(drift_engine* e) {
  if (!e) return fail(DRIFT_EINVAL, "null engine");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  return DRIFT_OK;
}

int drift_set_bpr(drift_engine* e, double alpha, double beta, const double* factor_table, const int64_t* class_off,
                  int32_t n_class, const int32_t* link_class) {
  if (!e || !factor_table || !class_off || !link_class || n_class <= 0) return fail(DRIFT_EINVAL, "drift_set_bpr");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  e->alpha = alpha;
  e->beta = beta;
  CU(e->factor.upload(factor_table, class_off[n_class]));
  CU(e->cls_off.upload(class_off, n_class + 1));
  CU(e->link_cls.upload(link_class, e->L));
  if (class_off[n_class] >= ((int64_t)1 << 31)) return fail(DRIFT_EINVAL, "drift_set_bpr: factor table too large");
  CU(e->link_rec.alloc(e->L, false));
  k_pack_linkrec<<<grid_for(e->L, 256, kSMs * 8), 256, 0, e->stream>>>(e->links(), e->link_rec.p);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));
  e->have_bpr = true;
  return DRIFT_OK;
}

int drift_add_agents(drift_engine* e, int64_t n_agents, int64_t agent_base, int64_t route_arena_links) {
  if (!e || n_agents <= 0 || agent_base < 0) return fail(DRIFT_EINVAL, "drift_add_agents: bad argument");
  if (e->N) return fail(DRIFT_ESTATE, "drift_add_agents: agents already created");
  if (agent_base + n_agents >= ((int64_t)1 << 32)) return fail(DRIFT_EINVAL, "agent ids must fit 32 bits");
  CU(cudaSetDevice(e->device));
  // padded to whole tiles, with room for the widest tile shape of the persistent kernel (TickArgs::tile_singles)
  const int64_t tile = (int64_t)kAdvThreads * kAdvPerThread;
  const int64_t Np = (n_agents + kAdvTileQuads + kAdvMaxSingles + tile - 1) / tile * tile;
  CU(e->state.alloc(Np));
  CU(e->dist.alloc(Np));
  CU(e->speed.alloc(Np));
  CU(e->elen.alloc(Np));
  CU(e->cur_link.alloc(Np, false));
  CU(cudaMemset(e->cur_link.p, 0xff, sizeof(int32_t) * Np));
  CU(e->next_link.alloc(Np, false));
  CU(cudaMemset(e->next_link.p, 0xff, sizeof(int32_t) * Np));
  CU(e->route_idx.alloc(Np));
  CU(e->route_len.alloc(Np));
  CU(e->route_off.alloc(Np));
  CU(e->trip_count.alloc(Np));
  CU(e->cur_node.alloc(Np));
  CU(e->chain_head.alloc(Np));
  CU(e->chain_cnt.alloc(Np));
  CU(e->next_off.alloc(2 * Np));
  CU(e->next_len.alloc(2 * Np));
  CU(e->next_status.alloc(2 * Np));
  if (route_arena_links <= 0) route_arena_links = std::max<int64_t>((int64_t)1 << 20, n_agents * 64);
  CU(e->arena.alloc(route_arena_links, false));
  // the spare arena of the compaction is allocated up front: a first cudaMalloc of several GB inside a
  // planning window costs ~100 ms (observed as a one-off spike of the step time)
  if (route_arena_links >= ((int64_t)1 << 26)) CU(e->arena2.alloc(route_arena_links, false));
  e->arena_cap = route_arena_links;
  // every sub-buffer can hold all events of a tick in the worst case only if the warps spread evenly;
  // sized with a 4x margin over the even share and checked at run time (overflow flag)
  e->ev_sub_cap = std::min<int64_t>(2 * n_agents + 1024, (2 * n_agents * 4) / kEvSub + 4096);
  e->ev_cap = e->ev_sub_cap * kEvSub;
  if (e->ev_cap > 0x7fffffff) return fail(DRIFT_EINVAL, "too many agents for one engine");
  CU(e->events.alloc(e->ev_cap, false));
  CU(e->send_n.alloc(1));
  e->pend_cap = n_agents;
  CU(e->pend_agent.alloc(n_agents));
  CU(e->pend_off.alloc(n_agents));
  CU(e->pend_len.alloc(n_agents));
  e->arr_cap = n_agents + n_agents / 8 + 4096;
  CU(e->arr_agent.alloc(e->arr_cap));
  CU(e->arr_tick.alloc(e->arr_cap));
  e->dl_cap = n_agents + n_agents / 8 + 4096;
  CU(e->dl_agent.alloc(e->dl_cap));
  CU(e->dl_tick.alloc(e->dl_cap));
  CU(e->dl_src.alloc(e->dl_cap));
  CU(e->dl_dst.alloc(e->dl_cap));
  CU(e->dl_status.alloc(e->dl_cap));
  CU(e->dl_off.alloc(e->dl_cap));
  CU(e->dl_len.alloc(e->dl_cap));
  CU(e->dl_epoch.alloc(e->dl_cap));
  CU(e->dep_req.alloc(n_agents));
  e->N = n_agents;
  e->Npad = Np;
  e->agent_base = (uint32_t)agent_base;
  return DRIFT_OK;
}

int drift_route_od(drift_engine* e, int32_t router, int64_t Q, const int32_t* src, const int32_t* dst,
                   int32_t use_congested, int32_t* status, int32_t* n_links) {
  if (!e || Q < 0 || (Q && (!src || !dst))) return fail(DRIFT_EINVAL, "drift_route_od: bad argument");
  if (!e->N) return fail(DRIFT_ESTATE, "drift_route_od: call drift_add_agents first (route arena)");
  CU(cudaSetDevice(e->device));
  e->last_Q = Q;
  if (Q == 0) return DRIFT_OK;
  int rc = ensure_qcap(e, Q);
  if (rc) return rc;
  CU(cudaMemcpyAsync(e->q_src.p, src, sizeof(int32_t) * Q, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->q_dst.p, dst, sizeof(int32_t) * Q, cudaMemcpyHostToDevice, e->stream));
  Counters c0;
  rc = fetch_counters(e, &c0);
  if (rc) return rc;
  if (c0.n_pending == 0) {
    rc = maybe_compact_arena(e, c0.arena_cursor);
    if (rc) return rc;
    if (e->compactions && (double)c0.arena_cursor >= e->arena_compact_frac * (double)e->arena_cap) {
      rc = fetch_counters(e, &c0);
      if (rc) return rc;
    }
  }
  e->last_Q = Q;
  e->last_arena_begin = c0.arena_cursor;
  rc = launch_router(e, router, Q, use_congested);
  if (rc) return rc;
  Counters c1;
  rc = fetch_counters(e, &c1);
  if (rc) return rc;
  e->last_arena_end = std::min<unsigned long long>(c1.arena_cursor, (unsigned long long)e->arena_cap);
  if (status) CU(cudaMemcpy(status, e->q_status.p, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost));
  if (n_links) CU(cudaMemcpy(n_links, e->q_len.p, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost));
  return check_overflow(e, c1);
}

int drift_fetch_routes(drift_engine* e, int32_t* links_out, int64_t cap) {
  if (!e || !links_out) return fail(DRIFT_EINVAL, "drift_fetch_routes");
  CU(cudaSetDevice(e->device));
  const int64_t Q = e->last_Q;
  if (Q == 0) return DRIFT_OK;
  std::vector<int64_t> off(Q);
  std::vector<int32_t> len(Q);
  CU(cudaMemcpy(off.data(), e->q_off.p, sizeof(int64_t) * Q, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(len.data(), e->q_len.p, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost));
  const int64_t span = (int64_t)(e->last_arena_end - e->last_arena_begin);
  std::vector<int32_t> buf((size_t)std::max<int64_t>(span, 1));
  if (span > 0)
    CU(cudaMemcpy(buf.data(), e->arena.p + e->last_arena_begin, sizeof(int32_t) * span, cudaMemcpyDeviceToHost));
  int64_t w = 0;
  for (int64_t q = 0; q < Q; ++q) {
    if (len[q] <= 0) continue;
    if (w + len[q] > cap) return fail(DRIFT_EINVAL, "drift_fetch_routes: output buffer too small");
    const int64_t rel = off[q] - (int64_t)e->last_arena_begin;
    if (rel < 0 || rel + len[q] > span) return fail(DRIFT_ESTATE, "drift_fetch_routes: stale batch");
    memcpy(links_out + w, buf.data() + rel, sizeof(int32_t) * len[q]);
    w += len[q];
  }
  return DRIFT_OK;
}

int drift_fetch_route_costs(drift_engine* e, double* cost_out, int64_t Q) {
  if (!e || !cost_out || Q > e->last_Q) return fail(DRIFT_EINVAL, "drift_fetch_route_costs");
  CU(cudaSetDevice(e->device));
  if (Q > 0) CU(cudaMemcpy(cost_out, e->q_cost.p, sizeof(double) * Q, cudaMemcpyDeviceToHost));
  return DRIFT_OK;
}

int drift_commit_departures(drift_engine* e, int64_t Q, const int32_t* agent) {
  if (!e || Q < 0 || Q > e->last_Q || (Q && !agent)) return fail(DRIFT_EINVAL, "drift_commit_departures");
  CU(cudaSetDevice(e->device));
  if (Q == 0) return DRIFT_OK;
  for (int64_t q = 0; q < Q; ++q)
    if (agent[q] < 0 || agent[q] >= e->N) return fail(DRIFT_EINVAL, "drift_commit_departures: agent out of range");
  if (has_duplicates(agent, Q)) return fail(DRIFT_EINVAL, "drift_commit_departures: an agent is listed twice");
  CU(cudaMemcpyAsync(e->q_agent.p, agent, sizeof(int32_t) * Q, cudaMemcpyHostToDevice, e->stream));
  k_commit<<<grid_for(Q, 256, kSMs), 256, 0, e->stream>>>(Q, e->q_agent.p, e->q_off.p, e->q_len.p, e->ctr.p,
                                                           e->pend_agent.p, e->pend_off.p, e->pend_len.p,
                                                           (int32_t)e->pend_cap, nullptr);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));  // `agent` is a caller buffer
  e->last_Q = 0;  // a staged batch is committed once: a second commit would inject its agents twice
  return DRIFT_OK;
}

int drift_submit_departures(drift_engine* e, int64_t n, const int32_t* agent, const int64_t* route_off,
                            const int32_t* route_links) {
  if (!e || n < 0 || (n && (!agent || !route_off))) return fail(DRIFT_EINVAL, "drift_submit_departures");
  if (!e->N) return fail(DRIFT_ESTATE, "drift_submit_departures: no agents");
  CU(cudaSetDevice(e->device));
  if (n == 0) return DRIFT_OK;
  const int64_t total = route_off[n];
  for (int64_t q = 0; q < n; ++q) {
    if (agent[q] < 0 || agent[q] >= e->N) return fail(DRIFT_EINVAL, "drift_submit_departures: agent out of range");
    if (route_off[q + 1] < route_off[q]) return fail(DRIFT_EINVAL, "drift_submit_departures: bad offsets");
  }
  if (has_duplicates(agent, n)) return fail(DRIFT_EINVAL, "drift_submit_departures: an agent is listed twice");
  for (int64_t k = 0; k < total; ++k)
    if (route_links[k] < 0 || route_links[k] >= e->L) return fail(DRIFT_EINVAL, "drift_submit_departures: bad link id");
  int rc = ensure_qcap(e, n);
  if (rc) return rc;
  DevBuf<int64_t> d_off;
  DevBuf<int32_t> d_links;
  CU(d_off.upload(route_off, n + 1));
  CU(d_links.upload(route_links, std::max<int64_t>(total, 1)));
  CU(cudaMemcpyAsync(e->q_agent.p, agent, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
  k_import_routes<<<grid_for(n, 128, kSMs), 128, 0, e->stream>>>(n, d_off.p, d_links.p, e->arena.p,
                                                                 (unsigned long long)e->arena_cap, e->ctr.p,
                                                                 e->q_off.p, e->q_len.p);
  k_commit<<<grid_for(n, 256, kSMs), 256, 0, e->stream>>>(n, e->q_agent.p, e->q_off.p, e->q_len.p, e->ctr.p,
                                                           e->pend_agent.p, e->pend_off.p, e->pend_len.p,
                                                           (int32_t)e->pend_cap, nullptr);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));
  e->last_Q = 0;
  return DRIFT_OK;
}

int drift_step(drift_engine* e, int32_t n_ticks, double delta) {
  if (!e || n_ticks < 0) return fail(DRIFT_EINVAL, "drift_step");
  if (!e->N) return fail(DRIFT_ESTATE, "drift_step: no agents");
  if (!e->have_bpr) return fail(DRIFT_ESTATE, "drift_step: call drift_set_bpr first");
  CU(cudaSetDevice(e->device));
  if (e->persistent && !e->profiling) return run_tick_loop(e, n_ticks, delta, false, 0.0);
  for (int32_t k = 0; k < n_ticks; ++k) {
    int rc = tick_begin(e, delta, false, 0.0);
    if (rc) return rc;
    rc = tick_finish(e, e->events.p, -1, false, delta);
    if (rc) return rc;
  }
  return DRIFT_OK;
}

int drift_set_demand(drift_engine* e, uint64_t seed, const double hourly24[24], const int32_t* start_node,
                     int32_t chain_capacity, int32_t router) {
  if (!e || !hourly24 || !start_node) return fail(DRIFT_EINVAL, "drift_set_demand");
  if (!e->N) return fail(DRIFT_ESTATE, "drift_set_demand: no agents");
  if (chain_capacity < 2 || chain_capacity > 64) return fail(DRIFT_EINVAL, "chain_capacity must be in [2, 64]");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  e->seed = seed;
  for (int h = 0; h < 24; ++h) e->hourly[h] = hourly24[h];
  for (int64_t i = 0; i < e->N; ++i)
    if (start_node[i] < 0 || start_node[i] >= e->V) return fail(DRIFT_EINVAL, "drift_set_demand: bad start node");
  CU(cudaMemcpy(e->cur_node.p, start_node, sizeof(int32_t) * e->N, cudaMemcpyHostToDevice));
  e->chain_cap = chain_capacity;
  CU(e->chain_dst.alloc(e->N * (int64_t)chain_capacity));
  CU(cudaMemset(e->chain_head.p, 0, e->Npad));
  CU(cudaMemset(e->chain_cnt.p, 0, e->Npad));
  CU(cudaMemset(e->next_status.p, 0, sizeof(int32_t) * 2 * e->Npad));
  e->demand_router = router;
  int rc = ensure_qcap(e, 2 * e->N);
  if (rc) return rc;
  {  // scratch of the in-tick fallback router: a few NetworkX-exact workers
    const int64_t per_worker = 2 * (int64_t)e->V * 16 + 2 * (int64_t)(e->Ef + 2) * (int64_t)sizeof(HeapItem);
    int64_t F = std::min<int64_t>(64, ((int64_t)2 << 30) / std::max<int64_t>(per_worker, 1));
    F = std::max<int64_t>(F, 1);
    CU(e->f_seen.alloc(F * 2 * e->V, false));
    CU(e->f_plink.alloc(F * 2 * e->V, false));
    CU(e->f_mark.alloc(F * 2 * e->V, true));
    CU(e->f_epoch.alloc(F, true));
    e->f_heap_cap = (int64_t)e->Ef + 2;
    CU(e->f_heap.alloc(F * 2 * e->f_heap_cap, false));
    e->f_workers = (int32_t)F;
  }
  e->have_demand = true;
  return DRIFT_OK;
}

static int refill_from(drift_engine* e, const int32_t* cand_dev, int32_t k) {
  k_refill_chains<<<grid_for(e->N, 256, kSMs * 8), 256, 0, e->stream>>>(e->agents(), cand_dev, k);
  CU(cudaGetLastError());
  e->total_launches += 1;
  return DRIFT_OK;
}

int drift_push_trips(drift_engine* e, const int32_t* candidates, int32_t k) {
  if (!e || !candidates || k <= 0) return fail(DRIFT_EINVAL, "drift_push_trips");
  if (!e->have_demand) return fail(DRIFT_ESTATE, "drift_push_trips: call drift_set_demand first");
  CU(cudaSetDevice(e->device));
  const int64_t total = e->N * (int64_t)k;
  if (e->trip_stage.n < total) {
    CU(cudaStreamSynchronize(e->stream));
    CU(e->trip_stage.alloc(total, false));
  }
  CU(cudaMemcpyAsync(e->trip_stage.p, candidates, sizeof(int32_t) * total, cudaMemcpyHostToDevice, e->stream));
  return refill_from(e, e->trip_stage.p, k);
}

int drift_store_trip_pool(drift_engine* e, const int32_t* candidates, int32_t n_blocks, int32_t k) {
  if (!e || !candidates || k <= 0 || n_blocks <= 0) return fail(DRIFT_EINVAL, "drift_store_trip_pool");
  if (!e->have_demand) return fail(DRIFT_ESTATE, "drift_store_trip_pool: call drift_set_demand first");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  CU(e->trip_pool.upload(candidates, (int64_t)n_blocks * e->N * k));
  e->pool_blocks = n_blocks;
  e->pool_k = k;
  return DRIFT_OK;
}

int drift_refill_from_pool(drift_engine* e, int32_t block) {
  if (!e || block < 0 || block >= e->pool_blocks) return fail(DRIFT_EINVAL, "drift_refill_from_pool");
  CU(cudaSetDevice(e->device));
  return refill_from(e, e->trip_pool.p + (int64_t)block * e->N * e->pool_k, (int32_t)e->pool_k);
}

// Window planning: route, as one large batch, every missing prefetched route (next two trips of
// every agent).  One host sync per window to learn the batch size.
static double departure_probability(drift_engine* e, double t, double delta);

// t_first = clock of the first tick of the window (lookahead planning evaluates the window's draws)
static int plan_window(drift_engine* e, double t_first, double delta) {
  Counters* c = e->ctr.p;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e->time_plan) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, e->stream);
  }
  CU(cudaMemsetAsync(&c->n_plan, 0, sizeof(int32_t), e->stream));
  if (e->plan_lookahead) {
    const int32_t W = e->route_window;
    std::vector<unsigned long long> th(W);
    double t = t_first;
    for (int32_t k = 0; k < W; ++k) {  // same thresholds as the tick kernel's departure_threshold()
      const double pk = departure_probability(e, t, delta);
      const double sc = std::ceil(pk * 9007199254740992.0);
      th[k] = !(pk > 0.0) ? 0ULL : (sc >= 18446744073709551615.0 ? ~0ULL : (unsigned long long)sc);
      t += delta;
    }
    if (e->plan_thresh.n < W) CU(e->plan_thresh.alloc(W, false));
    CU(cudaMemcpyAsync(e->plan_thresh.p, th.data(), sizeof(unsigned long long) * W, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));  // th is a stack-lifetime buffer
    int32_t hops = e->plan_arrival_hops;
    if (hops < 0) {
      const double reach = (double)W * delta * e->vmax_kph * kKphToMps;  // metres an agent can cover in the window
      hops = (int32_t)std::min(2.0e9, std::ceil(reach / std::max(e->min_len, 1e-3)) + 1.0);
    }
    k_plan_lookahead<<<grid_for((e->N + 1) / 2, 256, kSMs * 8), 256, 0, e->stream>>>(
        e->agents(), c, e->q_src.p, e->q_dst.p, e->q_agent.p, (int32_t)e->q_cap, e->seed, (int32_t)e->tick + 1, W,
        e->plan_thresh.p, hops);
  } else {
    k_plan_requests<<<grid_for(e->N, 256, kSMs * 8), 256, 0, e->stream>>>(e->agents(), c, e->q_src.p, e->q_dst.p,
                                                                           e->q_agent.p, (int32_t)e->q_cap);
  }
  CU(cudaGetLastError());
  e->total_launches += 1;
  Counters hc;
  int rc = fetch_counters(e, &hc);
  if (rc) return rc;
  const int64_t Q = std::min<int64_t>(hc.n_plan, e->q_cap);
  e->routes_planned += Q;
  if (hc.n_pending == 0) {  // committed, not yet injected departures point into the live arena (see maybe_compact_arena)
    rc = maybe_compact_arena(e, hc.arena_cursor);
    if (rc) return rc;
  }
  int64_t chunk = std::max<int64_t>(Q, 1);
  const bool by_labels = e->labels_ready && e->labels_enabled && !e->use_congested;
  if (e->demand_router == DRIFT_ROUTER_BATCHED && !by_labels) {
    rc = ensure_trees(e);
    if (rc) return rc;
    chunk = router_chunk(e, Q, e->use_congested);
  }
  for (int64_t q0 = 0; q0 < Q; q0 += chunk) {
    const int64_t n = std::min(chunk, Q - q0);
    rc = launch_router(e, e->demand_router, n, e->use_congested ? 1 : 0, nullptr, q0);
    if (rc) return rc;
  }
  if (Q > 0) {
    k_store_prefetch<<<grid_for(Q, 256, kSMs * 8), 256, 0, e->stream>>>(Q, e->q_agent.p, e->q_off.p, e->q_len.p,
                                                                       e->q_status.p, e->agents());
    CU(cudaGetLastError());
    e->total_launches += 1;
  }
  if (ev0) {
    cudaEventRecord(ev1, e->stream);
    cudaEventSynchronize(ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    e->plan_ms += ms;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  }
  return DRIFT_OK;
}

static double departure_probability(drift_engine* e, double t, double delta) {
  const int hour = (int)std::fmod((t + 21600.0) / 3600.0, 24.0);  // lib/agent.py:251-252
  return e->hourly[hour] * 4.0 * delta / 3600.0;                 // lib/agent.py:259,180
}

// one tick of device-driven demand up to (not including) the event resolution
static int demand_tick_begin(drift_engine* e, double delta, double t) {
  if (e->ticks_since_plan <= 0) {
    int rc = plan_window(e, t, delta);
    if (rc) return rc;
    e->ticks_since_plan = e->route_window;
  }
  e->ticks_since_plan -= 1;
  return tick_begin(e, delta, true, departure_probability(e, t, delta));
}

int drift_run(drift_engine* e, int32_t n_ticks, double delta, double t0) {
  if (!e || n_ticks < 0) return fail(DRIFT_EINVAL, "drift_run");
  if (!e->have_demand) return fail(DRIFT_ESTATE, "drift_run: call drift_set_demand first");
  if (!e->have_bpr) return fail(DRIFT_ESTATE, "drift_run: call drift_set_bpr first");
  CU(cudaSetDevice(e->device));
  double t = t0;
  if ((e->persistent && !e->profiling) || e->p2p_on) {
    // split into runs of ticks that share the departure probability and do not cross a planning boundary
    int32_t k = 0;
    while (k < n_ticks) {
      if (e->ticks_since_plan <= 0) {
        int rc = plan_window(e, t + delta, delta);
        if (rc) return rc;
        e->ticks_since_plan = e->route_window;
      }
      double tt = t + delta;  // repeated fp64 sum like lib/simulation_thread.py:295
      const double p = departure_probability(e, tt, delta);
      int32_t n = 1;
      while (k + n < n_ticks && n < e->ticks_since_plan) {
        const double t_next = tt + delta;
        if (departure_probability(e, t_next, delta) != p) break;
        tt = t_next;
        ++n;
      }
      int rc = run_tick_loop(e, n, delta, true, p);
      if (rc) return rc;
      e->ticks_since_plan -= n;
      t = tt;
      k += n;
    }
    return DRIFT_OK;
  }
  for (int32_t k = 0; k < n_ticks; ++k) {
    t += delta;
    int rc = demand_tick_begin(e, delta, t);
    if (rc) return rc;
    rc = tick_finish(e, e->events.p, -1, false, delta);
    if (rc) return rc;
  }
  return DRIFT_OK;
}

// local events of the tick -> contiguous send buffer; returns their number
static int compact_for_exchange(drift_engine* e, int64_t* n_events) {
  if (e->send.n < e->ev_cap) CU(e->send.alloc(e->ev_cap, false));
  k_compact_events<<<e->num_sms, 256, 0, e->stream>>>(e->events.p, (int32_t)e->ev_sub_cap, e->ctr.p, (int32_t)e->tick,
                                                     e->send.p, e->send_n.p);
  CU(cudaGetLastError());
  e->total_launches += 1;
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  rc = check_overflow(e, c);
  if (rc) return rc;
  int64_t tot = 0;
  for (int s = 0; s < kEvSub; ++s) tot += c.n_events[e->tick & 1][s];
  if (n_events) *n_events = tot;
  return DRIFT_OK;
}

int drift_tick_begin_demand(drift_engine* e, double delta, double t, int64_t* n_events) {
  if (!e || !e->have_demand || !e->have_bpr) return fail(DRIFT_ESTATE, "drift_tick_begin_demand");
  CU(cudaSetDevice(e->device));
  int rc = demand_tick_begin(e, delta, t);
  if (rc) return rc;
  e->last_delta = delta;
  return compact_for_exchange(e, n_events);
}

static int update_travel_times_from(drift_engine* e, const int32_t* volumes_dev);

int drift_update_travel_times(drift_engine* e) {
  if (!e || !e->have_bpr) return fail(DRIFT_ESTATE, "drift_update_travel_times");
  return update_travel_times_from(e, e->occ.p);
}

int drift_update_travel_times_with(drift_engine* e, const void* volumes_dev) {
  if (!e || !e->have_bpr || !volumes_dev) return fail(DRIFT_ESTATE, "drift_update_travel_times_with");
  return update_travel_times_from(e, (const int32_t*)volumes_dev);
}

static int update_travel_times_from(drift_engine* e, const int32_t* volumes_dev) {
  CU(cudaSetDevice(e->device));
  LinkArrays lk = e->links();
  lk.occ = const_cast<int32_t*>(volumes_dev);  // read only by k_travel_times
  k_travel_times<<<grid_for((e->L + 3) / 4, 256, kSMs * 8), 256, 0, e->stream>>>(lk);
  k_refresh_weights<<<grid_for(e->Ef, 256, kSMs * 8), 256, 0, e->stream>>>(e->Ef, e->link_tt.p, e->fwd_link.p,
                                                                           e->bwd_link.p, e->fwd_wc.p, e->bwd_wc.p);
  CU(cudaGetLastError());
  e->use_congested = true;  // from now on departures and the in-tick fallback route on tt
  return invalidate_trees(e);  // cached trees were built on the old weights
}

int drift_reroute_moving(drift_engine* e, int32_t router, int64_t* n_rerouted) {
  if (!e || !e->N) return fail(DRIFT_EINVAL, "drift_reroute_moving");
  CU(cudaSetDevice(e->device));
  if (n_rerouted) *n_rerouted = 0;
  int rc = ensure_qcap(e, e->N);
  if (rc) return rc;
  Counters* c = e->ctr.p;
  CU(cudaMemsetAsync(&c->n_plan, 0, sizeof(int32_t), e->stream));
  k_reroute_requests<<<grid_for(e->N, 256, kSMs * 8), 256, 0, e->stream>>>(e->agents(), e->link_v.p, e->arena.p, c,
                                                                            e->q_src.p, e->q_dst.p, e->q_agent.p,
                                                                            (int32_t)e->q_cap);
  CU(cudaGetLastError());
  Counters hc;
  rc = fetch_counters(e, &hc);
  if (rc) return rc;
  const int64_t Q = std::min<int64_t>(hc.n_plan, e->q_cap);
  int64_t chunk = std::max<int64_t>(Q, 1);
  if (router == DRIFT_ROUTER_BATCHED) {
    rc = ensure_trees(e);
    if (rc) return rc;
    chunk = router_chunk(e, Q, true);
  }
  for (int64_t q0 = 0; q0 < Q; q0 += chunk) {
    rc = launch_router(e, router, std::min(chunk, Q - q0), 1, nullptr, q0);
    if (rc) return rc;
  }
  if (Q > 0) {
    k_apply_reroutes<<<grid_for(Q, 128, kSMs * 8), 128, 0, e->stream>>>(Q, e->q_agent.p, e->q_off.p, e->q_len.p,
                                                                       e->q_status.p, e->agents(), e->arena.p,
                                                                       (unsigned long long)e->arena_cap, c);
    CU(cudaGetLastError());
  }
  e->total_launches += 2;
  e->reroutes_total += Q;
  if (n_rerouted) *n_rerouted = Q;
  return DRIFT_OK;
}

int drift_invalidate_prefetch(drift_engine* e, double t0, double delta, int32_t n_ticks) {
  if (!e || !e->have_demand || n_ticks <= 0) return fail(DRIFT_EINVAL, "drift_invalidate_prefetch");
  CU(cudaSetDevice(e->device));
  std::vector<double> p(n_ticks);
  double t = t0;
  for (int32_t k = 0; k < n_ticks; ++k) {
    t += delta;
    p[k] = departure_probability(e, t, delta);
  }
  DevBuf<double> dp;
  CU(dp.upload(p.data(), n_ticks));
  k_invalidate_prefetch<<<grid_for(e->N, 256, kSMs * 8), 256, 0, e->stream>>>(e->agents(), e->seed,
                                                                               (int32_t)e->tick + 1, n_ticks, dp.p);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));  // dp is freed on return
  e->ticks_since_plan = 0;              // plan again before the next tick
  return DRIFT_OK;
}

int drift_read_occupancy(drift_engine* e, int32_t* occ_out) {
  if (!e || !occ_out) return fail(DRIFT_EINVAL, "drift_read_occupancy");
  CU(cudaSetDevice(e->device));
  CU(cudaMemcpyAsync(occ_out, e->occ.p, sizeof(int32_t) * e->L, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return DRIFT_OK;
}

int drift_recount_occupancy(drift_engine* e, int32_t* occ_out) {
  if (!e || !occ_out || !e->N) return fail(DRIFT_EINVAL, "drift_recount_occupancy");
  CU(cudaSetDevice(e->device));
  CU(cudaMemsetAsync(e->hist.p, 0, sizeof(int32_t) * e->L, e->stream));
  k_recount<<<grid_for(e->N, kRecThreads * kRecPerThread, kSMs * 8), kRecThreads, 0, e->stream>>>(e->agents(), e->hist.p);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(occ_out, e->hist.p, sizeof(int32_t) * e->L, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return DRIFT_OK;
}

int drift_recount_device(drift_engine* e) {
  if (!e || !e->N) return fail(DRIFT_EINVAL, "drift_recount_device");
  CU(cudaSetDevice(e->device));
  CU(cudaMemsetAsync(e->hist.p, 0, sizeof(int32_t) * e->L, e->stream));
  k_recount<<<grid_for(e->N, kRecThreads * kRecPerThread, kSMs * 8), kRecThreads, 0, e->stream>>>(e->agents(), e->hist.p);
  CU(cudaGetLastError());
  e->total_launches += 1;
  return DRIFT_OK;
}

int drift_read_travel_times(drift_engine* e, double* tt_out) {
  if (!e || !tt_out) return fail(DRIFT_EINVAL, "drift_read_travel_times");
  CU(cudaSetDevice(e->device));
  CU(cudaMemcpyAsync(tt_out, e->link_tt.p, sizeof(double) * e->L, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return DRIFT_OK;
}

int drift_read_agents(drift_engine* e, uint8_t* state, double* dist_on_link, double* speed, double* link_len,
                      int32_t* cur_link, int32_t* route_idx, int32_t* route_len, int32_t* trip_count,
                      int32_t* cur_node) {
  if (!e || !e->N) return fail(DRIFT_EINVAL, "drift_read_agents");
  CU(cudaSetDevice(e->device));
  const int64_t n = e->N;
  cudaStream_t s = e->stream;
  if (state) CU(cudaMemcpyAsync(state, e->state.p, n, cudaMemcpyDeviceToHost, s));
  if (dist_on_link) CU(cudaMemcpyAsync(dist_on_link, e->dist.p, 8 * n, cudaMemcpyDeviceToHost, s));
  if (speed) CU(cudaMemcpyAsync(speed, e->speed.p, 8 * n, cudaMemcpyDeviceToHost, s));
  if (link_len) CU(cudaMemcpyAsync(link_len, e->elen.p, 8 * n, cudaMemcpyDeviceToHost, s));
  if (cur_link) CU(cudaMemcpyAsync(cur_link, e->cur_link.p, 4 * n, cudaMemcpyDeviceToHost, s));
  if (route_idx) CU(cudaMemcpyAsync(route_idx, e->route_idx.p, 4 * n, cudaMemcpyDeviceToHost, s));
  if (route_len) CU(cudaMemcpyAsync(route_len, e->route_len.p, 4 * n, cudaMemcpyDeviceToHost, s));
  if (trip_count) CU(cudaMemcpyAsync(trip_count, e->trip_count.p, 4 * n, cudaMemcpyDeviceToHost, s));
  if (cur_node) CU(cudaMemcpyAsync(cur_node, e->cur_node.p, 4 * n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return DRIFT_OK;
}

int drift_read_agent_route(drift_engine* e, int64_t agent, int32_t* links_out, int64_t cap, int64_t* n) {
  if (!e || agent < 0 || agent >= e->N || !n) return fail(DRIFT_EINVAL, "drift_read_agent_route");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  int64_t off = 0;
  int32_t len = 0;
  CU(cudaMemcpy(&off, e->route_off.p + agent, sizeof(int64_t), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&len, e->route_len.p + agent, sizeof(int32_t), cudaMemcpyDeviceToHost));
  *n = len;
  if (links_out && len > 0) {
    if (len > cap) return fail(DRIFT_EINVAL, "drift_read_agent_route: buffer too small");
    CU(cudaMemcpy(links_out, e->arena.p + off, sizeof(int32_t) * len, cudaMemcpyDeviceToHost));
  }
  return DRIFT_OK;
}

static int drain_ring(drift_engine* e, unsigned long long total, unsigned long long* drained, int64_t ring_cap,
                      int64_t cap, int64_t* n, std::initializer_list<std::pair<void*, const void*>> cols,
                      std::initializer_list<size_t> widths) {
  unsigned long long avail = total - *drained;
  if ((int64_t)avail > ring_cap) return fail(DRIFT_EOVERFLOW, "ring overflow: drain more often");
  int64_t take = std::min<int64_t>((int64_t)avail, cap);
  *n = take;
  if (take == 0) return DRIFT_OK;
  const int64_t start = (int64_t)(*drained % (unsigned long long)ring_cap);
  const int64_t first = std::min<int64_t>(take, ring_cap - start);
  auto w = widths.begin();
  for (auto& c : cols) {
    const size_t sz = *w++;
    if (!c.first) continue;
    CU(cudaMemcpy(c.first, (const char*)c.second + sz * start, sz * first, cudaMemcpyDeviceToHost));
    if (take > first)
      CU(cudaMemcpy((char*)c.first + sz * first, c.second, sz * (take - first), cudaMemcpyDeviceToHost));
  }
  *drained += (unsigned long long)take;
  return DRIFT_OK;
}

int drift_export_routes(drift_engine* e, int64_t n, const int32_t* agents, int32_t* n_links, double* metres,
                        int32_t* links_out, int64_t cap) {
  if (!e || n < 0 || (n && !agents)) return fail(DRIFT_EINVAL, "drift_export_routes");
  CU(cudaSetDevice(e->device));
  if (n == 0) return DRIFT_OK;
  for (int64_t r = 0; r < n; ++r)
    if (agents[r] < 0 || agents[r] >= e->N) return fail(DRIFT_EINVAL, "drift_export_routes: agent out of range");
  DevBuf<int32_t> d_agents, d_len, d_links;
  DevBuf<int64_t> d_off;
  DevBuf<double> d_m;
  CU(d_agents.upload(agents, n));
  CU(d_len.alloc(n));
  CU(d_m.alloc(n));
  CU(cudaStreamSynchronize(e->stream));
  // pass 1: lengths (and metres)
  k_gather_routes<<<grid_for(n, 128, kSMs * 8), 128, 0, e->stream>>>(n, d_agents.p, e->agents(), e->arena.p,
                                                                     e->link_len.p, nullptr, nullptr, d_m.p, d_len.p);
  CU(cudaGetLastError());
  std::vector<int32_t> len(n);
  CU(cudaMemcpyAsync(len.data(), d_len.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
  if (metres) CU(cudaMemcpyAsync(metres, d_m.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (n_links) memcpy(n_links, len.data(), sizeof(int32_t) * n);
  if (links_out) {
    std::vector<int64_t> off(n + 1, 0);
    for (int64_t r = 0; r < n; ++r) off[r + 1] = off[r] + len[r];
    if (off[n] > cap) return fail(DRIFT_EINVAL, "drift_export_routes: output buffer too small (%lld needed)", (long long)off[n]);
    if (off[n] > 0) {
      CU(d_off.upload(off.data(), n + 1));
      CU(d_links.alloc(off[n], false));
      k_gather_routes<<<grid_for(n, 128, kSMs * 8), 128, 0, e->stream>>>(n, d_agents.p, e->agents(), e->arena.p,
                                                                         e->link_len.p, d_off.p, d_links.p, d_m.p,
                                                                         nullptr);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(links_out, d_links.p, sizeof(int32_t) * off[n], cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
    }
  }
  return DRIFT_OK;
}

int drift_drain_arrivals(drift_engine* e, int32_t* agent, int32_t* tick, int64_t cap, int64_t* n) {
  if (!e || !n || cap < 0) return fail(DRIFT_EINVAL, "drift_drain_arrivals");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  rc = check_overflow(e, c);
  if (rc) return rc;
  return drain_ring(e, c.n_arrivals, &e->arr_drained, e->arr_cap, cap, n,
                    {{agent, e->arr_agent.p}, {tick, e->arr_tick.p}}, {4, 4});
}

int drift_drain_departures(drift_engine* e, int32_t* agent, int32_t* tick, int32_t* src, int32_t* dst,
                           int32_t* status, int64_t cap, int64_t* n) {
  if (!e || !n || cap < 0) return fail(DRIFT_EINVAL, "drift_drain_departures");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  rc = check_overflow(e, c);
  if (rc) return rc;
  return drain_ring(e, c.n_departures, &e->dl_drained, e->dl_cap, cap, n,
                    {{agent, e->dl_agent.p}, {tick, e->dl_tick.p}, {src, e->dl_src.p}, {dst, e->dl_dst.p},
                     {status, e->dl_status.p}},
                    {4, 4, 4, 4, 4});
}

int drift_drain_departures_routes(drift_engine* e, int32_t* agent, int32_t* tick, int32_t* src, int32_t* dst,
                                  int32_t* status, int64_t* route_off, int32_t* route_len, int32_t* route_epoch,
                                  int64_t cap, int64_t* n) {
  if (!e || !n || cap < 0) return fail(DRIFT_EINVAL, "drift_drain_departures_routes");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  rc = check_overflow(e, c);
  if (rc) return rc;
  return drain_ring(e, c.n_departures, &e->dl_drained, e->dl_cap, cap, n,
                    {{agent, e->dl_agent.p}, {tick, e->dl_tick.p}, {src, e->dl_src.p}, {dst, e->dl_dst.p},
                     {status, e->dl_status.p}, {route_off, e->dl_off.p}, {route_len, e->dl_len.p},
                     {route_epoch, e->dl_epoch.p}},
                    {4, 4, 4, 4, 4, 8, 4, 4});
}

int drift_export_logged_routes(drift_engine* e, int64_t n, const int64_t* route_off, const int32_t* route_len,
                               const int32_t* route_epoch, double* metres, int32_t* links_out, int64_t cap) {
  if (!e || n < 0 || (n && (!route_off || !route_len || !route_epoch || !metres)))
    return fail(DRIFT_EINVAL, "drift_export_logged_routes");
  CU(cudaSetDevice(e->device));
  if (n == 0) return DRIFT_OK;
  std::vector<int64_t> off(n + 1, 0);
  for (int64_t r = 0; r < n; ++r) {
    if (route_epoch[r] != (int32_t)e->compactions)
      return fail(DRIFT_ESTATE, "drift_export_logged_routes: the route arena was compacted since departure %lld was "
                                "logged; drain and export after every block of at most route_window ticks", (long long)r);
    if (route_len[r] < 0 || route_off[r] < 0 || route_off[r] + route_len[r] > e->arena_cap)
      return fail(DRIFT_EINVAL, "drift_export_logged_routes: bad arena range");
    off[r + 1] = off[r] + route_len[r];
  }
  if (links_out && off[n] > cap) return fail(DRIFT_EINVAL, "drift_export_logged_routes: output buffer too small");
  DevBuf<long long> d_off;
  DevBuf<int32_t> d_len, d_links;
  DevBuf<int64_t> d_out;
  DevBuf<double> d_m;
  CU(d_off.upload((const long long*)route_off, n));
  CU(d_len.upload(route_len, n));
  CU(d_m.alloc(n));
  if (links_out && off[n] > 0) {
    CU(d_out.upload(off.data(), n + 1));
    CU(d_links.alloc(off[n], false));
  }
  CU(cudaStreamSynchronize(e->stream));
  k_gather_logged<<<grid_for(n, 128, kSMs * 8), 128, 0, e->stream>>>(n, d_off.p, d_len.p, e->arena.p, e->link_len.p,
                                                                     d_out.p, d_links.p, d_m.p);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(metres, d_m.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
  if (links_out && off[n] > 0)
    CU(cudaMemcpyAsync(links_out, d_links.p, sizeof(int32_t) * off[n], cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return DRIFT_OK;
}

int drift_enable_entry_trace(drift_engine* e, int64_t cap) {
  if (!e || cap < 0) return fail(DRIFT_EINVAL, "drift_enable_entry_trace");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  CU(e->ent_tick.alloc(cap));
  CU(e->ent_agent.alloc(cap));
  CU(e->ent_link.alloc(cap));
  CU(e->ent_speed.alloc(cap));
  e->ent_cap = cap;
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  e->ent_drained = c.n_entries;
  return DRIFT_OK;
}

int drift_drain_entries(drift_engine* e, int32_t* tick, int32_t* agent, int32_t* link, double* speed, int64_t cap,
                        int64_t* n) {
  if (!e || !n || cap < 0) return fail(DRIFT_EINVAL, "drift_drain_entries");
  if (!e->ent_cap) return fail(DRIFT_ESTATE, "entry trace is off");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  rc = check_overflow(e, c);
  if (rc) return rc;
  return drain_ring(e, c.n_entries, &e->ent_drained, e->ent_cap, cap, n,
                    {{tick, e->ent_tick.p}, {agent, e->ent_agent.p}, {link, e->ent_link.p}, {speed, e->ent_speed.p}},
                    {4, 4, 4, 8});
}

int drift_stats(drift_engine* e, drift_stats_out* out) {
  if (!e || !out) return fail(DRIFT_EINVAL, "drift_stats");
  CU(cudaSetDevice(e->device));
  Counters* c = e->ctr.p;
  CU(cudaMemsetAsync(&c->moving, 0, 3 * sizeof(long long), e->stream));
  if (e->N) {
    k_stats<<<kSMs * 4, 256, 0, e->stream>>>(e->agents(), e->links(), c);
    CU(cudaGetLastError());
  }
  Counters h;
  int rc = fetch_counters(e, &h);
  if (rc) return rc;
  out->moving_agents = h.moving;
  out->total_agents_on_roads = h.on_roads;
  out->links_with_traffic = h.links_with_traffic;
  out->total_capacity = e->total_capacity;
  out->arrivals_total = (int64_t)h.n_arrivals;
  out->departures_total = (int64_t)h.n_departures;
  out->route_arena_used = (int64_t)h.arena_cursor;
  out->arena_compactions = e->compactions;
  out->fallback_routes = (int64_t)h.n_fallback;
  out->tie_fallbacks = (int64_t)h.n_tie_fallback;
  out->events_last_tick = 0;
  for (int sb = 0; sb < kEvSub; ++sb) out->events_last_tick += h.n_events[e->tick & 1][sb];
  return check_overflow(e, h);
}

int drift_tick_index(drift_engine* e, int64_t* tick) {
  if (!e || !tick) return fail(DRIFT_EINVAL, "drift_tick_index");
  *tick = e->tick;
  return DRIFT_OK;
}

int drift_tick_begin(drift_engine* e, double delta, int64_t* n_events) {
  if (!e || !e->N || !e->have_bpr) return fail(DRIFT_ESTATE, "drift_tick_begin");
  CU(cudaSetDevice(e->device));
  int rc = tick_begin(e, delta, false, 0.0);
  if (rc) return rc;
  e->last_delta = delta;
  return compact_for_exchange(e, n_events);
}

int drift_events_ptr(drift_engine* e, void** dev_ptr, int64_t* capacity) {
  if (!e || !dev_ptr) return fail(DRIFT_EINVAL, "drift_events_ptr");
  if (e->send.n < e->ev_cap) CU(e->send.alloc(e->ev_cap, false));
  *dev_ptr = e->send.p;
  if (capacity) *capacity = e->ev_cap;
  return DRIFT_OK;
}

int drift_tick_finish(drift_engine* e, const void* events_dev, int64_t n_events) {
  if (!e || !events_dev || n_events < 0) return fail(DRIFT_EINVAL, "drift_tick_finish");
  if (n_events > 0x7fffffff) return fail(DRIFT_EINVAL, "drift_tick_finish: too many events");
  CU(cudaSetDevice(e->device));
  return tick_finish(e, (Event*)events_dev, n_events, true, e->last_delta);
}

int drift_set_stream(drift_engine* e, void* cuda_stream) {
  if (!e) return fail(DRIFT_EINVAL, "drift_set_stream");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  if (e->own_stream) cudaStreamDestroy(e->stream);
  e->stream = (cudaStream_t)cuda_stream;
  e->own_stream = false;
  return DRIFT_OK;
}

int drift_tick_begin_async(drift_engine* e, double delta, double t, void* send_dev, int32_t cap) {
  if (!e || !e->N || !e->have_bpr || !send_dev || cap <= 0) return fail(DRIFT_EINVAL, "drift_tick_begin_async");
  CU(cudaSetDevice(e->device));
  int rc = e->have_demand ? demand_tick_begin(e, delta, t) : tick_begin(e, delta, false, 0.0);
  if (rc) return rc;
  e->last_delta = delta;
  k_pack_events<<<e->num_sms, 256, 0, e->stream>>>(e->events.p, (int32_t)e->ev_sub_cap, e->ctr.p, (int32_t)e->tick,
                                                  (Event*)send_dev, cap);
  CU(cudaGetLastError());
  e->total_launches += 1;
  return DRIFT_OK;
}

int drift_tick_finish_gathered(drift_engine* e, void* recv_dev, int32_t world, int32_t cap) {
  if (!e || !recv_dev || world <= 0 || world > 16 || cap <= 0) return fail(DRIFT_EINVAL, "drift_tick_finish_gathered");
  CU(cudaSetDevice(e->device));
  const TickArgs p = make_tick_args(e, e->last_delta, e->have_demand);
  const int32_t tick = (int32_t)e->tick;
  const int g = e->num_sms * 4;
  k_link_gathered<<<g, 256, 0, e->stream>>>(e->links(), (Event*)recv_dev, world, cap, tick);
  k_resolve_gathered<<<g, 256, 0, e->stream>>>(p, (const Event*)recv_dev, world, cap, tick);
  k_reset_counters<<<1, 32, 0, e->stream>>>(e->ctr.p, tick);
  CU(cudaGetLastError());
  e->total_launches += 3;
  return DRIFT_OK;
}

int drift_p2p_export(drift_engine* e, int32_t rank, int32_t world, int64_t n_total, int32_t cap, void* handles_out) {
  if (!e || !handles_out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || cap <= 0 || !e->N)
    return fail(DRIFT_EINVAL, "drift_p2p_export: bad arguments (call drift_add_agents first; world <= %d)", kMaxRanks);
  if (e->tick != 0) return fail(DRIFT_ESTATE, "drift_p2p_export: the exchange mode is fixed before the first tick");
  CU(cudaSetDevice(e->device));
  const int64_t base = n_total / world, rem = n_total % world;
  if (e->N != base + (rank < rem ? 1 : 0) || (int64_t)e->agent_base != rank * base + std::min<int64_t>(rank, rem))
    return fail(DRIFT_EINVAL, "drift_p2p_export: this engine does not hold rank %d's contiguous block of %lld agents",
                rank, (long long)n_total);
  CU(e->p2p_inbox.alloc((int64_t)2 * world * cap, true));  // double-buffered by tick parity
  CU(e->p2p_flags.alloc(std::max<int64_t>(3 * kMaxRanks, (2 << 20) / 8), true));
  CU(e->p2p_out_cnt.alloc(kMaxRanks, true));
  memset(&e->p2p, 0, sizeof(e->p2p));
  e->p2p.rank = rank;
  e->p2p.world = world;
  e->p2p.cap = cap;
  e->p2p.shard_base = (uint32_t)base;
  e->p2p.shard_rem = (uint32_t)rem;
  e->p2p.out_cnt = e->p2p_out_cnt.p;
  e->p2p.inbox[rank] = e->p2p_inbox.p;
  e->p2p.flags[rank] = e->p2p_flags.p;
  e->p2p.speed[rank] = e->speed.p;
  e->p2p.elen[rank] = e->elen.p;
  cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)handles_out;
  CU(cudaIpcGetMemHandle(&h[0], e->p2p_inbox.p));
  CU(cudaIpcGetMemHandle(&h[1], e->p2p_flags.p));
  CU(cudaIpcGetMemHandle(&h[2], e->speed.p));
  CU(cudaIpcGetMemHandle(&h[3], e->elen.p));
  return DRIFT_OK;
}

int drift_p2p_connect(drift_engine* e, const void* all_handles) {
  if (!e || !all_handles || e->p2p.world < 1 || !e->p2p_inbox.p) return fail(DRIFT_ESTATE, "drift_p2p_connect: export first");
  CU(cudaSetDevice(e->device));
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
  for (int r = 0; r < e->p2p.world; ++r) {
    if (r == e->p2p.rank) continue;
    void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; ++k) {
      CU(cudaIpcOpenMemHandle(&ptr[k], h[r * 4 + k], cudaIpcMemLazyEnablePeerAccess));
      e->p2p_opened.push_back(ptr[k]);
    }
    e->p2p.inbox[r] = (Event*)ptr[0];
    e->p2p.flags[r] = (unsigned long long*)ptr[1];
    e->p2p.speed[r] = (double*)ptr[2];
    e->p2p.elen[r] = (double*)ptr[3];
  }
  e->p2p_on = true;
  return DRIFT_OK;
}

int drift_max_events(drift_engine* e, int64_t* max_events, int32_t reset) {
  if (!e || !max_events) return fail(DRIFT_EINVAL, "drift_max_events");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  *max_events = c.max_events;
  if (reset) CU(cudaMemsetAsync(&e->ctr.p->max_events, 0, sizeof(int32_t), e->stream));
  return check_overflow(e, c);
}

int drift_set_profiling(drift_engine* e, int32_t on) {
  if (!e) return fail(DRIFT_EINVAL, "drift_set_profiling");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  e->ev_pool.clear();
  e->adv_marks.clear();
  e->tick_ev0 = e->tick_ev1 = -1;
  e->profiling = on != 0;
  e->adv_launches = e->total_launches = 0;
  return DRIFT_OK;
}

int drift_kernel_times(drift_engine* e, double* advance_ms, int64_t* advance_launches, double* total_ms,
                       int64_t* total_launches) {
  if (!e) return fail(DRIFT_EINVAL, "drift_kernel_times");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  double adv = 0;
  for (auto& m : e->adv_marks) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e->ev_pool[m.first], e->ev_pool[m.second]));
    adv += ms;
  }
  if (advance_ms) *advance_ms = adv;
  if (advance_launches) *advance_launches = (int64_t)e->adv_marks.size();
  if (total_ms) *total_ms = 0;
  if (total_launches) *total_launches = e->total_launches;
  return DRIFT_OK;
}

int drift_tree_stats(drift_engine* e, double* build_ms, int64_t* launches, int64_t* jobs, int64_t* relaxed,
                     int64_t* improved, int64_t* expanded, int64_t* iters, int64_t* far_visits, int32_t reset) {
  if (!e) return fail(DRIFT_EINVAL, "drift_tree_stats");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  for (auto& t : e->tree_timers) {
    float ms = 0;
    cudaEventElapsedTime(&ms, t.first, t.second);
    e->tree_ms += ms;
    e->tree_launches += 1;
    cudaEventDestroy(t.first);
    cudaEventDestroy(t.second);
  }
  e->tree_timers.clear();
  if (build_ms) *build_ms = e->tree_ms;
  if (launches) *launches = e->tree_launches;
  if (jobs) *jobs = (int64_t)c.tree_jobs;
  if (relaxed) *relaxed = (int64_t)c.tree_relaxed;
  if (improved) *improved = (int64_t)c.tree_improved;
  if (expanded) *expanded = (int64_t)c.tree_expanded;
  if (iters) *iters = (int64_t)c.tree_iters;
  if (far_visits) *far_visits = (int64_t)c.tree_far_visits;
  if (reset) {
    CU(cudaMemsetAsync(&e->ctr.p->tree_relaxed, 0, 6 * sizeof(unsigned long long), e->stream));
    e->tree_ms = 0;
    e->tree_launches = 0;
  }
  return DRIFT_OK;
}

int drift_phase_times(drift_engine* e, double* advance_ns, double* inject_ns, double* resolve_ns, double* plan_ms,
                      int64_t* ticks, int32_t reset) {
  if (!e) return fail(DRIFT_EINVAL, "drift_phase_times");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  if (advance_ns) *advance_ns = (double)c.phase_ns[0];
  if (inject_ns) *inject_ns = (double)c.phase_ns[1];
  if (resolve_ns) *resolve_ns = (double)c.phase_ns[2];
  if (plan_ms) *plan_ms = e->plan_ms;
  if (ticks) *ticks = e->loop_ticks;
  if (reset) {
    CU(cudaMemsetAsync(&e->ctr.p->phase_ns[0], 0, 4 * sizeof(unsigned long long), e->stream));
    e->plan_ms = 0;
    e->loop_ticks = 0;
  }
  return DRIFT_OK;
}

int drift_p2p_phase_detail(drift_engine* e, double ns_out[8], int32_t reset) {
  if (!e || !ns_out) return fail(DRIFT_EINVAL, "drift_p2p_phase_detail");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  for (int k = 0; k < 8; ++k) ns_out[k] = (double)c.p2p_ns[k];
  if (reset) CU(cudaMemsetAsync(&e->ctr.p->p2p_ns[0], 0, 8 * sizeof(unsigned long long), e->stream));
  return DRIFT_OK;
}

int drift_set_landmarks(drift_engine* e, int32_t K, const double* dist_from, const double* dist_to) {
  if (!e || K < 0 || (K && (!dist_from || !dist_to))) return fail(DRIFT_EINVAL, "drift_set_landmarks");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  e->lm_K = 0;
  if (K == 0) return DRIFT_OK;
  if (K > kMaxLandmarks) return fail(DRIFT_EINVAL, "drift_set_landmarks: at most %d landmarks", kMaxLandmarks);
  const int64_t V = e->V;
  std::vector<double> tab((size_t)V * 2 * K);  // one row of 2K bounds per node: a potential reads one line
  for (int32_t k = 0; k < K; ++k)
    for (int64_t v = 0; v < V; ++v) {
      tab[(size_t)v * 2 * K + k] = dist_from[(size_t)k * V + v];
      tab[(size_t)v * 2 * K + K + k] = dist_to[(size_t)k * V + v];
    }
  CU(e->lm_tab.upload(tab.data(), (int64_t)tab.size()));
  e->lm_K = K;
  return DRIFT_OK;
}

// Host utility (no device work): out[i] = pow(x[i], y) with the C library's pow -- the function CPython's
// float ** float reaches (floatobject.c: float_pow -> pow), which NumPy's vectorised power does not
// reproduce bit for bit.  Used by the vectorised s-t selectors (drift_b200/fast_selection.py) to keep
// lib/st_selection.py:416 `(S_i * A_j) / (d_ij ** self.beta)` and :352 `(dx*dx + dy*dy)**0.5` exact.
int drift_host_pow(const double* x, double y, double* out, int64_t n) {
  if (n < 0 || (n && (!x || !out))) return fail(DRIFT_EINVAL, "drift_host_pow");
  for (int64_t i = 0; i < n; ++i) out[i] = std::pow(x[i], y);
  return DRIFT_OK;
}

// ---- contraction hierarchy + hub labels ------------------------------------------------------------
struct drift_hierarchy {
  Hierarchy h;
  int32_t max_label = 0;
  double mean_label = 0;
  double build_s = 0;
};

int drift_hier_build(int32_t V, int32_t Ef, const int32_t* fwd_rowptr, const int32_t* fwd_col, const double* fwd_w,
                     const int32_t* fwd_link, int32_t witness_settled, double time_limit_s, drift_hierarchy** out) {
  if (!out || V <= 0 || Ef < 0 || !fwd_rowptr || (Ef && (!fwd_col || !fwd_w || !fwd_link)))
    return fail(DRIFT_EINVAL, "drift_hier_build: bad argument");
  *out = nullptr;
  if (fwd_rowptr[V] != Ef) return fail(DRIFT_EINVAL, "drift_hier_build: rowptr[V] != Ef");
  for (int32_t k = 0; k < Ef; ++k)
    if (fwd_col[k] < 0 || fwd_col[k] >= V || !(fwd_w[k] > 0.0) || !(fwd_w[k] < 1e300))
      return fail(DRIFT_EINVAL, "drift_hier_build: travel times must be positive and finite");
  drift_hierarchy* hh = new drift_hierarchy();
  HierParams prm;
  if (witness_settled > 0) prm.witness_settled = witness_settled;
  prm.time_limit_s = time_limit_s;
  const auto t0 = std::chrono::steady_clock::now();
  if (!build_hierarchy(V, fwd_rowptr, fwd_col, fwd_w, fwd_link, prm, &hh->h)) {
    delete hh;
    return fail(DRIFT_ESTATE, "drift_hier_build: time limit of %.0f s hit (the graph has no hierarchy worth finding)",
                time_limit_s);
  }
  hh->build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // label sizes on a sample (the device sizes its tables and the label arena from these)
  std::vector<uint32_t> stamp(V, 0u);
  std::vector<int32_t> stack;
  uint32_t epoch = 0;
  const int32_t samples = std::min<int32_t>(V, 4096);
  uint64_t x = 0x9E3779B97F4A7C15ULL;
  double sum = 0;
  for (int32_t k = 0; k < samples; ++k) {
    x ^= x << 13, x ^= x >> 7, x ^= x << 17;
    const int32_t v = samples == V ? k : (int32_t)(x % (uint64_t)V);
    const int32_t nf = closure_size(hh->h.up, v, stamp, ++epoch, stack);
    const int32_t nb = closure_size(hh->h.down, v, stamp, ++epoch, stack);
    hh->max_label = std::max(hh->max_label, std::max(nf, nb));
    sum += nf + nb;
  }
  hh->mean_label = sum / (2.0 * samples);
  *out = hh;
  return DRIFT_OK;
}

int drift_hier_info(const drift_hierarchy* h, int64_t* n_up, int64_t* n_down, int64_t* shortcuts, int32_t* max_level,
                    int32_t* max_label, double* mean_label, double* build_seconds) {
  if (!h) return fail(DRIFT_EINVAL, "drift_hier_info");
  if (n_up) *n_up = (int64_t)h->h.up.other.size();
  if (n_down) *n_down = (int64_t)h->h.down.other.size();
  if (shortcuts) *shortcuts = h->h.shortcuts;
  if (max_level) *max_level = h->h.max_level;
  if (max_label) *max_label = h->max_label;
  if (mean_label) *mean_label = h->mean_label;
  if (build_seconds) *build_seconds = h->build_s;
  return DRIFT_OK;
}

int drift_hier_export(const drift_hierarchy* h, drift_hier_desc* d) {
  if (!h || !d) return fail(DRIFT_EINVAL, "drift_hier_export");
  const Hierarchy& H = h->h;
  auto put = [](const auto& v, auto* dst) {
    if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  put(H.level, const_cast<int32_t*>(d->level));
  put(H.up.rowptr, const_cast<int32_t*>(d->up_rowptr));
  put(H.up.other, const_cast<int32_t*>(d->up_other));
  put(H.up.w, const_cast<double*>(d->up_w));
  put(H.up.a, const_cast<int32_t*>(d->up_a));
  put(H.up.b, const_cast<int32_t*>(d->up_b));
  put(H.up.hops, const_cast<int32_t*>(d->up_hops));
  put(H.down.rowptr, const_cast<int32_t*>(d->down_rowptr));
  put(H.down.other, const_cast<int32_t*>(d->down_other));
  put(H.down.w, const_cast<double*>(d->down_w));
  put(H.down.a, const_cast<int32_t*>(d->down_a));
  put(H.down.b, const_cast<int32_t*>(d->down_b));
  put(H.down.hops, const_cast<int32_t*>(d->down_hops));
  d->V = H.V;
  d->n_up = (int64_t)H.up.other.size();
  d->n_down = (int64_t)H.down.other.size();
  d->max_level = H.max_level;
  d->max_label = h->max_label;
  d->mean_label = h->mean_label;
  return DRIFT_OK;
}

int drift_hier_free(drift_hierarchy* h) {
  delete h;
  return DRIFT_OK;
}

int drift_set_hierarchy(drift_engine* e, const drift_hier_desc* d) {
  if (!e || !d || d->V != e->V || !d->level || !d->up_rowptr || !d->down_rowptr)
    return fail(DRIFT_EINVAL, "drift_set_hierarchy: bad argument (hierarchy of another graph?)");
  if (d->n_up >= ((int64_t)1 << 30) || d->n_down >= ((int64_t)1 << 30))
    return fail(DRIFT_EINVAL, "drift_set_hierarchy: too many arcs");
  if (d->max_level + 2 > kPathItems / 2)
    return fail(DRIFT_EINVAL, "drift_set_hierarchy: %d levels, the route unpacker handles climbs of %d arcs",
                d->max_level + 1, kPathItems / 2);
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  e->labels_ready = false;
  const int32_t V = e->V;
  // table size of the label builder: a power of two with head room over the largest sampled label
  int32_t n_max = 512;
  while (n_max < (int32_t)(1.5 * d->max_label) + 64) n_max <<= 1;
  if (n_max > 4096)
    return fail(DRIFT_EOVERFLOW, "drift_set_hierarchy: labels of up to ~%d hubs do not fit the builder's shared-memory "
                                 "table (4096): this graph keeps the tree / bounded-search router", d->max_label);
  int32_t log2H = 0;
  while ((1 << log2H) < 2 * n_max) ++log2H;
  const size_t smem = (size_t)48 * n_max;
  CU(cudaFuncSetAttribute(k_build_labels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(e->h_level.upload(d->level, V));
  CU(e->hu_rowptr.upload(d->up_rowptr, V + 1));
  CU(e->hu_other.upload(d->up_other, d->n_up));
  CU(e->hu_w.upload(d->up_w, d->n_up));
  CU(e->hu_a.upload(d->up_a, d->n_up));
  CU(e->hu_b.upload(d->up_b, d->n_up));
  CU(e->hu_hops.upload(d->up_hops, d->n_up));
  CU(e->hd_rowptr.upload(d->down_rowptr, V + 1));
  CU(e->hd_other.upload(d->down_other, d->n_down));
  CU(e->hd_w.upload(d->down_w, d->n_down));
  CU(e->hd_a.upload(d->down_a, d->n_down));
  CU(e->hd_b.upload(d->down_b, d->n_down));
  CU(e->hd_hops.upload(d->down_hops, d->n_down));
  {  // what unpacking needs about an arc in one record: {a, b, hops, hops of the first half (a down arc)}
    std::vector<int4> rec((size_t)std::max<int64_t>(std::max(d->n_up, d->n_down), 1));
    for (int side = 0; side < 2; ++side) {
      const int64_t n = side == 0 ? d->n_up : d->n_down;
      const int32_t* a = side == 0 ? d->up_a : d->down_a;
      const int32_t* b = side == 0 ? d->up_b : d->down_b;
      const int32_t* h = side == 0 ? d->up_hops : d->down_hops;
      for (int64_t k = 0; k < n; ++k) {
        if (b[k] >= 0 && (a[k] < 0 || a[k] >= d->n_down || b[k] >= d->n_up))
          return fail(DRIFT_EINVAL, "drift_set_hierarchy: bad child arc");
        rec[k] = make_int4(a[k], b[k], h[k], b[k] < 0 ? 0 : d->down_hops[a[k]]);
      }
      CU((side == 0 ? e->hu_rec : e->hd_rec).upload(rec.data(), n));
    }
  }
  // nodes by level, highest first
  const int32_t nl = d->max_level + 1;
  std::vector<int32_t> start(nl + 1, 0), order(V);
  for (int32_t v = 0; v < V; ++v) {
    if (d->level[v] < 0 || d->level[v] > d->max_level) return fail(DRIFT_EINVAL, "drift_set_hierarchy: bad level");
    start[nl - 1 - d->level[v] + 1] += 1;
  }
  for (int32_t k = 0; k < nl; ++k) start[k + 1] += start[k];
  {
    std::vector<int32_t> fill(start.begin(), start.end() - 1);
    for (int32_t v = 0; v < V; ++v) order[fill[nl - 1 - d->level[v]]++] = v;
  }
  CU(e->h_order.upload(order.data(), V));
  CU(e->lab_off.alloc((int64_t)2 * V));
  CU(e->lab_cnt.alloc((int64_t)2 * V));
  CU(e->lab_flag.alloc(1));
  CU(e->lab_cursor.alloc(1));
  int64_t cap = (int64_t)(2.0 * d->mean_label * V * 1.25) + (1 << 20);
  for (int attempt = 0; attempt < 3; ++attempt) {
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if ((double)cap * 20.0 > 0.6 * (double)free_b)
      return fail(DRIFT_ENOMEM, "drift_set_hierarchy: %.1f GB of labels do not fit (%.1f GB free)", cap * 20.0 / 1e9,
                  free_b / 1e9);
    if (cap >= ((int64_t)1 << 32)) return fail(DRIFT_EOVERFLOW, "drift_set_hierarchy: more than 2^32 label entries");
    CU(e->lab_hub.alloc(cap, false));
    CU(e->lab_par.alloc(cap, false));
    CU(e->lab_pnext.alloc(cap, false));
    CU(e->lab_dist.alloc(cap, false));
    CU(cudaMemsetAsync(e->lab_flag.p, 0, sizeof(int32_t), e->stream));
    CU(cudaMemsetAsync(e->lab_cursor.p, 0, sizeof(unsigned long long), e->stream));
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, e->stream);
    for (int side = 0; side < 2; ++side)
      for (int32_t k = 0; k < nl; ++k) {
        const int32_t n = start[k + 1] - start[k];
        if (n <= 0) continue;
        k_build_labels<<<n, kLabelThreads, smem, e->stream>>>(side == 0 ? e->hier_up() : e->hier_down(), e->labels(side),
                                                              e->h_order.p, start[k], n_max, log2H, e->lab_cursor.p,
                                                              (unsigned long long)cap, e->lab_flag.p);
      }
    cudaEventRecord(ev1, e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    e->label_build_ms = ms;
    int32_t flag = 0;
    unsigned long long used = 0;
    CU(cudaMemcpy(&flag, e->lab_flag.p, sizeof(flag), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&used, e->lab_cursor.p, sizeof(used), cudaMemcpyDeviceToHost));
    if (flag & 1) {
      e->lab_hub.release(), e->lab_par.release(), e->lab_pnext.release(), e->lab_dist.release();
      return fail(DRIFT_EOVERFLOW, "drift_set_hierarchy: a label outgrew the builder's table of %d hubs", n_max);
    }
    if (!(flag & 2)) {
      e->label_entries = (int64_t)used;
      e->label_cap = cap;
      e->labels_ready = true;
      return DRIFT_OK;
    }
    cap = (int64_t)((double)used * 1.05) + (1 << 20);  // the arena was too small: now the exact size is known
  }
  return fail(DRIFT_EOVERFLOW, "drift_set_hierarchy: label arena");
}

int drift_label_stats(drift_engine* e, int64_t* entries, double* build_ms, double* query_ms, int64_t* launches,
                      int64_t* queries, int64_t* entries_read, int32_t reset) {
  if (!e) return fail(DRIFT_EINVAL, "drift_label_stats");
  CU(cudaSetDevice(e->device));
  Counters c;
  int rc = fetch_counters(e, &c);
  if (rc) return rc;
  for (auto& t : e->label_timers) {
    float ms = 0;
    cudaEventElapsedTime(&ms, t.first, t.second);
    e->label_ms += ms;
    e->label_launches += 1;
    cudaEventDestroy(t.first);
    cudaEventDestroy(t.second);
  }
  e->label_timers.clear();
  if (entries) *entries = e->labels_ready ? e->label_entries : 0;
  if (build_ms) *build_ms = e->label_build_ms;
  if (query_ms) *query_ms = e->label_ms;
  if (launches) *launches = e->label_launches;
  if (queries) *queries = e->label_queries_host;
  if (entries_read) *entries_read = (int64_t)c.label_entries;
  if (reset) {
    CU(cudaMemsetAsync(&e->ctr.p->label_entries, 0, 2 * sizeof(unsigned long long), e->stream));
    e->label_ms = 0;
    e->label_launches = 0;
    e->label_queries_host = 0;
  }
  return DRIFT_OK;
}

int drift_time_passes(drift_engine* e, int32_t iters, double* recount_us, double* travel_times_us) {
  if (!e || iters <= 0 || !e->N || !e->have_bpr) return fail(DRIFT_EINVAL, "drift_time_passes");
  CU(cudaSetDevice(e->device));
  cudaEvent_t ev[4];
  for (auto& x : ev) CU(cudaEventCreate(&x));
  const int gr = grid_for(e->N, kRecThreads * kRecPerThread, kSMs * 8);
  const int gt = grid_for((e->L + 3) / 4, 256, kSMs * 8);
  for (int w = 0; w < 2; ++w) {  // warm-up
    CU(cudaMemsetAsync(e->hist.p, 0, sizeof(int32_t) * e->L, e->stream));
    k_recount<<<gr, kRecThreads, 0, e->stream>>>(e->agents(), e->hist.p);
    k_travel_times<<<gt, 256, 0, e->stream>>>(e->links());
  }
  CU(cudaEventRecord(ev[0], e->stream));
  for (int i = 0; i < iters; ++i) {
    CU(cudaMemsetAsync(e->hist.p, 0, sizeof(int32_t) * e->L, e->stream));  // the zero-fill is part of the recount
    k_recount<<<gr, kRecThreads, 0, e->stream>>>(e->agents(), e->hist.p);
  }
  CU(cudaEventRecord(ev[1], e->stream));
  CU(cudaEventRecord(ev[2], e->stream));
  for (int i = 0; i < iters; ++i) k_travel_times<<<gt, 256, 0, e->stream>>>(e->links());
  CU(cudaEventRecord(ev[3], e->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(e->stream));
  float a = 0, b = 0;
  CU(cudaEventElapsedTime(&a, ev[0], ev[1]));
  CU(cudaEventElapsedTime(&b, ev[2], ev[3]));
  for (auto& x : ev) cudaEventDestroy(x);
  if (recount_us) *recount_us = 1e3 * a / iters;
  if (travel_times_us) *travel_times_us = 1e3 * b / iters;
  e->total_launches += 2 * iters;
  return DRIFT_OK;
}

int drift_set_option(drift_engine* e, const char* name, double value) {
  if (!e || !name) return fail(DRIFT_EINVAL, "drift_set_option");
  std::string s(name);
  if (s == "tree_cache_bytes") {
    if (e->t_max_slots) return fail(DRIFT_ESTATE, "tree_cache_bytes must be set before the first batched route");
    e->tree_budget_bytes = (int64_t)value;
    e->tree_budget_set = true;
  } else if (s == "time_plan") {
    e->time_plan = value != 0;
  } else if (s == "persistent") {
    e->persistent = value != 0;
  } else if (s == "coop_blocks_per_sm") {
    e->coop_blocks_per_sm = std::max(1, (int)value);
  } else if (s == "plan_lookahead") {
    e->plan_lookahead = value != 0;
  } else if (s == "plan_arrival_hops") {
    e->plan_arrival_hops = (int32_t)value;
  } else if (s == "route_window") {
    if (value < 1) return fail(DRIFT_EINVAL, "route_window must be >= 1");
    e->route_window = (int32_t)value;
  } else if (s == "replan") {
    e->ticks_since_plan = 0;
  } else if (s == "tree_pop_threshold") {
    e->t_pop_thresh = (int32_t)value;
  } else if (s == "arena_compact_frac") {
    e->arena_compact_frac = std::min(0.95, std::max(0.05, value));
  } else if (s == "one_wave_tiles") {
    e->one_wave = value != 0;
  } else if (s == "tree_wide_regs") {
    e->tree_wide_regs = value < 0 ? -1 : (value != 0 ? 1 : 0);
  } else if (s == "pot_cache") {
    e->pot_cache = value != 0;
  } else if (s == "congested_landmarks") {
    e->congested_landmarks = value != 0;
  } else if (s == "router_chunk") {
    e->router_chunk_opt = std::max<int64_t>(0, (int64_t)value);
  } else if (s == "detect_ties") {
    e->detect_ties = value < 0 ? -1 : (value != 0 ? 1 : 0);
  } else if (s == "use_labels") {
    e->labels_enabled = value != 0;
  } else if (s == "debug_mask") {
    e->debug_mask = (int32_t)value;
  } else if (s == "p2p_suspend") {
    // timing probe of bench.py: run the next ticks on this rank's agents only (k_tick_loop) although the peer
    // exchange is connected; the state is NOT the sharded run's afterwards
    if (!e->p2p.world || !e->p2p_inbox.p) return fail(DRIFT_ESTATE, "p2p_suspend: no peer exchange to suspend");
    e->p2p_on = value == 0;
    if (!e->p2p_on) {
      // with the peer exchange a link's volume lives on its owner; a local run needs the volumes of THIS rank's
      // agents on every link: recount them (otherwise leaving agents would drive the local counts negative)
      CU(cudaSetDevice(e->device));
      CU(cudaMemsetAsync(e->hist.p, 0, sizeof(int32_t) * e->L, e->stream));
      k_recount<<<grid_for(e->N, kRecThreads * kRecPerThread, kSMs * 8), kRecThreads, 0, e->stream>>>(e->agents(), e->hist.p);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(e->occ.p, e->hist.p, sizeof(int32_t) * e->L, cudaMemcpyDeviceToDevice, e->stream));
    }
  } else if (s == "tree_sides") {
    if (e->t_max_slots) return fail(DRIFT_ESTATE, "tree_sides must be set before the first batched route");
    e->t_sides = value >= 2 ? 2 : 1;
  } else if (s == "tree_scratch_bytes") {
    if (e->t_max_slots) return fail(DRIFT_ESTATE, "tree_scratch_bytes must be set before the first batched route");
    e->tree_scratch_bytes = (int64_t)value;
    e->tree_scratch_set = true;
  } else if (s == "tree_threads") {
    if (e->t_max_slots) return fail(DRIFT_ESTATE, "tree_threads must be set before the first batched route");
    const int v = (int)value;
    if (v != 64 && v != 128 && v != 256) return fail(DRIFT_EINVAL, "tree_threads: 64, 128 or 256");
    e->t_threads = v;
  } else if (s == "bounded_max") {
    e->t_bounded_max = std::max(0, std::min((int)value, kMaxBounded));
  } else if (s == "tree_delta_scale") {
    e->t_delta *= value;
  } else {
    return fail(DRIFT_EINVAL, "drift_set_option: unknown option '%s'", name);
  }
  return DRIFT_OK;
}

int drift_device_ptr(drift_engine* e, const char* name, void** ptr, int64_t* n_elems) {
  if (!e || !name || !ptr) return fail(DRIFT_EINVAL, "drift_device_ptr");
  std::string s(name);
  int64_t n = 0;
  void* p = nullptr;
  if (s == "occ") p = e->occ.p, n = e->L;
  else if (s == "volume_recount") p = e->hist.p, n = e->L;
  else if (s == "state") p = e->state.p, n = e->N;
  else if (s == "dist") p = e->dist.p, n = e->N;
  else if (s == "speed") p = e->speed.p, n = e->N;
  else if (s == "link_len") p = e->elen.p, n = e->N;
  else if (s == "cur_link") p = e->cur_link.p, n = e->N;
  else if (s == "events") p = e->send.p, n = e->send.n;
  else if (s == "travel_time") p = e->link_tt.p, n = e->L;
  else return fail(DRIFT_EINVAL, "drift_device_ptr: unknown buffer '%s'", name);
  *ptr = p;
  if (n_elems) *n_elems = n;
  return DRIFT_OK;
}

}  // extern "C"

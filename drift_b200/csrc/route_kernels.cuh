// drift_b200 -- routing kernels.
//
// k_route_nx: bit-exact restatement of what DRIFT calls at lib/agent.py:94,
//   nx.shortest_path(G, s, t, weight='travel_time') -> NetworkX 3.6.1 bidirectional_dijkstra
//   (networkx/algorithms/shortest_paths/weighted.py:2387-2459; third-party, algorithm restated,
//   tie-breaking rules in SURVEY.md Appendix B).  The search is inherently sequential per query
//   (one shared push counter orders equal-distance heap entries, strict alternation of the two
//   directions), so the parallelism is across queries: one query per thread, private binary
//   heap ordered by (dist, push counter) and epoch-stamped per-node scratch.
#pragma once
#include "common.cuh"

namespace drift {

struct GraphArrays {
  const int32_t* fwd_rowptr;
  const int32_t* fwd_col;
  const double* fwd_w;
  const int32_t* fwd_link;
  const int32_t* bwd_rowptr;
  const int32_t* bwd_col;
  const double* bwd_w;
  const int32_t* bwd_link;
  const int32_t* link_u;
  const int32_t* link_v;
  int32_t V;
  int32_t L;
};

struct RouteScratch {
  double* seen;      // [W][2][V] tentative distance
  int32_t* plink;    // [W][2][V] link of the relaxing hop (predecessor = its other end)
  uint32_t* mark;    // [W][2][V] (epoch << 1) | finalised
  HeapItem* heap;    // [W][2][heap_cap]
  uint32_t* epoch;   // [W]
  int64_t heap_cap;
  int32_t workers;
};

struct RouteOut {
  int32_t* status;   // [Q]
  int32_t* n_links;  // [Q]
  int64_t* off;      // [Q] start in the arena
  double* cost;      // [Q]
};

__device__ __forceinline__ bool heap_less(const HeapItem& x, const HeapItem& y) {
  return x.dist < y.dist || (x.dist == y.dist && x.cnt < y.cnt);
}

__device__ __forceinline__ void heap_push(HeapItem* h, int& n, HeapItem it) {
  int i = n++;
  while (i > 0) {
    const int p = (i - 1) >> 1;
    if (!heap_less(it, h[p])) break;
    h[i] = h[p];
    i = p;
  }
  h[i] = it;
}

__device__ __forceinline__ HeapItem heap_pop(HeapItem* h, int& n) {
  const HeapItem top = h[0];
  const HeapItem last = h[--n];
  int i = 0;
  for (;;) {
    int c = 2 * i + 1;
    if (c >= n) break;
    if (c + 1 < n && heap_less(h[c + 1], h[c])) ++c;
    if (!heap_less(h[c], last)) break;
    h[i] = h[c];
    i = c;
  }
  if (n > 0) h[i] = last;
  return top;
}

struct RouteResult {
  int32_t status;
  int32_t n_links;
  int64_t off;
  double cost;
};

// One NetworkX-exact query by one thread.  w_fwd / w_bwd: per-CSR-entry weights (static
// travel_time, or congestion-updated).  `worker` selects the private scratch.
__device__ inline RouteResult route_nx_one(const GraphArrays& g, const double* __restrict__ w_fwd,
                                           const double* __restrict__ w_bwd, int32_t s, int32_t t,
                                           const RouteScratch& sc, int worker, uint32_t& epoch,
                                           int32_t* __restrict__ arena, unsigned long long arena_cap, Counters* ctr,
                                           const double* __restrict__ link_cost) {
  RouteResult res;
  res.status = DRIFT_ROUTE_NOPATH;
  res.n_links = 0;
  res.off = 0;
  res.cost = 0.0;
  if (s < 0 || t < 0 || s >= g.V || t >= g.V) return res;  // agent.py:87-91
  if (s == t) {  // weighted.py:2393-2394
    res.status = DRIFT_ROUTE_TRIVIAL;
    return res;
  }
  const int64_t V = g.V;
  double* seen[2] = {sc.seen + ((int64_t)worker * 2 + 0) * V, sc.seen + ((int64_t)worker * 2 + 1) * V};
  int32_t* plink[2] = {sc.plink + ((int64_t)worker * 2 + 0) * V, sc.plink + ((int64_t)worker * 2 + 1) * V};
  uint32_t* mark[2] = {sc.mark + ((int64_t)worker * 2 + 0) * V, sc.mark + ((int64_t)worker * 2 + 1) * V};
  HeapItem* heap[2] = {sc.heap + ((int64_t)worker * 2 + 0) * sc.heap_cap,
                       sc.heap + ((int64_t)worker * 2 + 1) * sc.heap_cap};
  ++epoch;
  const uint32_t seen_tag = epoch << 1;
  int hn[2] = {0, 0};
  uint32_t pushes = 2;
  seen[0][s] = 0.0;
  mark[0][s] = seen_tag;
  plink[0][s] = -1;
  seen[1][t] = 0.0;
  mark[1][t] = seen_tag;
  plink[1][t] = -1;
  heap_push(heap[0], hn[0], HeapItem{0.0, 0u, s});
  heap_push(heap[1], hn[1], HeapItem{0.0, 1u, t});
  bool have_best = false;
  double best = 0.0;
  int32_t meet = -1;
  int dir = 1;
  bool found = false, overflow = false;
  while (hn[0] > 0 && hn[1] > 0) {
    dir = 1 - dir;  // forward first, strict alternation; a stale pop still uses the turn
    const HeapItem it = heap_pop(heap[dir], hn[dir]);
    const int32_t v = it.node;
    if (mark[dir][v] == (seen_tag | 1u)) continue;
    mark[dir][v] = seen_tag | 1u;
    if (mark[1 - dir][v] == (seen_tag | 1u)) {
      found = true;
      break;
    }
    const int32_t* rowptr = dir == 0 ? g.fwd_rowptr : g.bwd_rowptr;
    const int32_t* col = dir == 0 ? g.fwd_col : g.bwd_col;
    const double* wts = dir == 0 ? w_fwd : w_bwd;
    const int32_t* lnk = dir == 0 ? g.fwd_link : g.bwd_link;
    const int32_t e1 = rowptr[v + 1];
    for (int32_t e = rowptr[v]; e < e1; ++e) {
      const int32_t w = col[e];
      const double cand = __dadd_rn(it.dist, wts[e]);
      const uint32_t mw = mark[dir][w];
      if (mw == (seen_tag | 1u)) continue;  // finalised in this direction
      if (mw != seen_tag || cand < seen[dir][w]) {
        seen[dir][w] = cand;
        mark[dir][w] = seen_tag;
        plink[dir][w] = lnk[e];
        if (hn[dir] >= sc.heap_cap) {
          overflow = true;
          break;
        }
        heap_push(heap[dir], hn[dir], HeapItem{cand, pushes++, w});
        if ((mark[1 - dir][w] >> 1) == epoch) {
          const double tot = __dadd_rn(cand, seen[1 - dir][w]);
          if (!have_best || best > tot) {
            have_best = true;
            best = tot;
            meet = w;
          }
        }
      }
    }
    if (overflow) break;
  }
  if (overflow) {
    atomicOr(&ctr->overflow, kOvfHeap);
    return res;
  }
  if (!found) return res;  // weighted.py:2459 NetworkXNoPath
  int nf = 0, nb = 0;
  for (int32_t x = meet; x != s; x = g.link_u[plink[0][x]]) ++nf;
  for (int32_t x = meet; x != t; x = g.link_v[plink[1][x]]) ++nb;
  const int n = nf + nb;
  const unsigned long long off = atomicAdd(&ctr->arena_cursor, (unsigned long long)n);
  if (off + n > arena_cap) {
    atomicOr(&ctr->overflow, kOvfArena);
    return res;
  }
  int k = nf;
  for (int32_t x = meet; x != s;) {
    const int32_t l = plink[0][x];
    arena[off + (--k)] = l;
    x = g.link_u[l];
  }
  k = nf;
  for (int32_t x = meet; x != t;) {
    const int32_t l = plink[1][x];
    arena[off + (k++)] = l;
    x = g.link_v[l];
  }
  double c = 0.0;
  for (int j = 0; j < n; ++j) c = __dadd_rn(c, link_cost[arena[off + j]]);
  res.status = DRIFT_ROUTE_OK;
  res.n_links = n;
  res.off = (int64_t)off;
  res.cost = c;
  return res;
}

__global__ void __launch_bounds__(32) k_route_nx(GraphArrays g, const double* __restrict__ w_fwd,
                                                  const double* __restrict__ w_bwd, int64_t Q,
                                                  const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                  RouteScratch sc, int32_t* __restrict__ arena,
                                                  unsigned long long arena_cap, Counters* ctr, RouteOut out,
                                                  const double* __restrict__ link_cost,
                                                  const int32_t* __restrict__ n_dev,
                                                  const int32_t* __restrict__ list = nullptr) {
  const int worker = blockIdx.x * blockDim.x + threadIdx.x;
  if (worker >= sc.workers) return;
  if (n_dev) Q = *n_dev;  // device-side batch size
  uint32_t epoch = sc.epoch[worker];
  for (int64_t k = worker; k < Q; k += sc.workers) {
    const int64_t q = list ? (int64_t)list[k] : k;  // list: only these queries of the batch (tie fallback)
    const RouteResult r = route_nx_one(g, w_fwd, w_bwd, src[q], dst[q], sc, worker, epoch, arena, arena_cap, ctr,
                                       link_cost);
    out.status[q] = r.status;
    out.n_links[q] = r.n_links;
    out.off[q] = r.off;
    out.cost[q] = r.cost;
  }
  sc.epoch[worker] = epoch;
}

// Host-supplied routes (trace replay): copy into the arena.
__global__ void k_import_routes(int64_t n, const int64_t* __restrict__ route_off,
                                const int32_t* __restrict__ route_links, int32_t* __restrict__ arena,
                                unsigned long long arena_cap, Counters* ctr, int64_t* __restrict__ out_off,
                                int32_t* __restrict__ out_len) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = route_off[r], len = route_off[r + 1] - b;
    out_len[r] = (int32_t)len;
    out_off[r] = 0;
    if (len <= 0) continue;
    const unsigned long long off = atomicAdd(&ctr->arena_cursor, (unsigned long long)len);
    if (off + len > arena_cap) {
      atomicOr(&ctr->overflow, kOvfArena);
      out_len[r] = 0;
      continue;
    }
    for (int64_t j = 0; j < len; ++j) arena[off + j] = route_links[b + j];
    out_off[r] = (int64_t)off;
  }
}

// Pending departures of the next tick: (agent, arena offset, hops) from the staged batch.
__global__ void k_commit(int64_t Q, const int32_t* __restrict__ agent, const int64_t* __restrict__ q_off,
                         const int32_t* __restrict__ q_len, Counters* ctr, int32_t* __restrict__ pend_agent,
                         int64_t* __restrict__ pend_off, int32_t* __restrict__ pend_len, int32_t pend_cap,
                         const int32_t* __restrict__ n_dev) {
  if (n_dev) Q = *n_dev;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    if (q_len[q] <= 0) continue;
    const int p = atomicAdd(&ctr->n_pending, 1);
    if (p < pend_cap) {
      pend_agent[p] = agent[q];
      pend_off[p] = q_off[q];
      pend_len[p] = q_len[q];
    } else {
      atomicOr(&ctr->overflow, kOvfPending);  // the injection clamps n_pending to pend_cap
    }
  }
}

// Window planning: every agent missing a prefetched route for one of its next two trips asks
// for it (source of trip k+1 = destination of trip k, known from the chain).
__global__ void k_plan_requests(AgentArrays a, Counters* ctr, int32_t* __restrict__ q_src,
                                int32_t* __restrict__ q_dst, int32_t* __restrict__ q_agent, int32_t q_cap) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cnt = a.chain_cnt[i], head = a.chain_head[i];
    const int32_t* ring = a.chain_dst + i * a.chain_cap;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k >= cnt) break;
      if (a.next_status[k * a.npad + i] != 0) continue;
      const int32_t s = k == 0 ? a.cur_node[i] : ring[head];
      const int r = atomicAdd(&ctr->n_plan, 1);
      if (r < q_cap) {
        q_src[r] = s;
        q_dst[r] = ring[(head + k) % a.chain_cap];
        q_agent[r] = (int32_t)i * 2 + k;
      }
    }
  }
}

// Lookahead planning: the departure draws are counter-based (seed, agent, tick), so the planner can
// evaluate the coming window's draws and ask for routes only where departures can happen in it:
// agents with one hit need the route of their next trip, agents with two or more hits also the one
// after it (a short first trip can end inside the window).  Moving agents count only if they may
// arrive within the window (at most `max_hops_left` links to go).  A third departure inside one
// window falls back to the in-tick router.  One thread per agent pair (the two agents of a pair share
// a Philox block, common.cuh).
__global__ void k_plan_lookahead(AgentArrays a, Counters* ctr, int32_t* __restrict__ q_src, int32_t* __restrict__ q_dst,
                                 int32_t* __restrict__ q_agent, int32_t q_cap, uint64_t seed, int32_t tick0,
                                 int32_t n_ticks, const unsigned long long* __restrict__ thresh,
                                 int32_t max_hops_left) {
  const int64_t pairs = (a.n + 1) / 2;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < pairs; t += (int64_t)gridDim.x * blockDim.x) {
    bool want[2] = {false, false};
    for (int h = 0; h < 2; ++h) {
      const int64_t i = 2 * t + h;
      if (i >= a.n || a.chain_cnt[i] == 0) continue;
      if (a.next_status[i] != 0 && (a.chain_cnt[i] < 2 || a.next_status[a.npad + i] != 0)) continue;  // nothing missing
      want[h] = a.state[i] != kMoving || (a.route_len[i] - a.route_idx[i] - 1 <= max_hops_left);
    }
    if (!want[0] && !want[1]) continue;
    const uint32_t gid0 = a.base + (uint32_t)(2 * t);
    int hits[2] = {0, 0};
    for (int k = 0; k < n_ticks; ++k) {
      const unsigned long long th = thresh[k];
      if ((gid0 & 1u) == 0) {
        uint32_t o[4];
        philox4x32_10(seed, gid0 >> 1, (uint32_t)(tick0 + k), o);
        hits[0] += bits53(o[0], o[1]) < th;
        hits[1] += bits53(o[2], o[3]) < th;
      } else {
        if (want[0]) hits[0] += philox_bits53(seed, gid0, (uint32_t)(tick0 + k)) < th;
        if (want[1]) hits[1] += philox_bits53(seed, gid0 + 1u, (uint32_t)(tick0 + k)) < th;
      }
      if ((hits[0] >= 2 || !want[0]) && (hits[1] >= 2 || !want[1])) break;
    }
    for (int h = 0; h < 2; ++h) {
      if (!want[h] || hits[h] == 0) continue;
      const int64_t i = 2 * t + h;
      const int head = a.chain_head[i];
      const int32_t* ring = a.chain_dst + i * a.chain_cap;
      for (int k = 0; k < 2; ++k) {
        if (k == 1 && (hits[h] < 2 || a.chain_cnt[i] < 2)) break;
        if (a.next_status[(int64_t)k * a.npad + i] != 0) continue;
        const int r = atomicAdd(&ctr->n_plan, 1);
        if (r < q_cap) {
          // trip k starts where trip k-1 ends; trip 0 at the agent's position / current target
          q_src[r] = k == 0 ? a.cur_node[i] : ring[head];
          q_dst[r] = ring[(head + k) % a.chain_cap];
          q_agent[r] = (int32_t)i * 2 + k;
        }
      }
    }
  }
}

// Extension (congestion-aware rerouting, SURVEY.md 8c last row): every moving agent asks for a new
// route from the head of its current link to its destination.
__global__ void k_reroute_requests(AgentArrays a, const int32_t* __restrict__ link_v,
                                   const int32_t* __restrict__ arena, Counters* ctr,
                                   int32_t* __restrict__ q_src, int32_t* __restrict__ q_dst,
                                   int32_t* __restrict__ q_agent, int32_t q_cap) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    if (a.state[i] != kMoving) continue;
    const int32_t l = a.cur_link[i];
    if (l < 0 || a.next_link[i] < 0) continue;  // on its last link: nothing to re-plan
    const int32_t t = link_v[arena[a.route_off[i] + a.route_len[i] - 1]];  // head of the route's last link
    const int r = atomicAdd(&ctr->n_plan, 1);
    if (r < q_cap) {
      q_src[r] = link_v[l];
      q_dst[r] = t;
      q_agent[r] = (int32_t)i;
    }
  }
}

// New route = current link + new path; the agent stays where it is on the current link.
__global__ void k_apply_reroutes(int64_t Q, const int32_t* __restrict__ q_agent, const int64_t* __restrict__ q_off,
                                 const int32_t* __restrict__ q_len, const int32_t* __restrict__ q_status,
                                 AgentArrays a, int32_t* __restrict__ arena, unsigned long long arena_cap,
                                 Counters* ctr) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    if (q_status[q] != DRIFT_ROUTE_OK || q_len[q] <= 0) continue;
    const int32_t i = q_agent[q];
    const int32_t n = q_len[q] + 1;
    const unsigned long long off = atomicAdd(&ctr->arena_cursor, (unsigned long long)n);
    if (off + n > arena_cap) {
      atomicOr(&ctr->overflow, kOvfArena);
      continue;
    }
    arena[off] = a.cur_link[i];
    for (int k = 1; k < n; ++k) arena[off + k] = arena[q_off[q] + k - 1];
    a.route_off[i] = (int64_t)off;
    a.route_len[i] = n;
    a.route_idx[i] = 0;
    a.next_link[i] = arena[off + 1];
  }
}

// Route arena compaction: copy the live routes (the route of every moving agent and the prefetched
// routes) into a fresh arena and repoint the agents.  One warp per agent: one reservation for its
// (up to three) routes, coalesced copies.
__global__ void k_compact_routes(AgentArrays a, const int32_t* __restrict__ src, int32_t* __restrict__ dst,
                                 unsigned long long* __restrict__ cursor, unsigned long long cap, Counters* ctr) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < a.n; i += warps) {
    int32_t len[3] = {0, 0, 0};
    int64_t off[3] = {0, 0, 0};
    if (a.state[i] == kMoving) {
      len[0] = a.route_len[i];
      off[0] = a.route_off[i];
    }
    for (int k = 0; k < 2; ++k) {
      const int64_t idx = (int64_t)k * a.npad + i;
      if (a.next_status[idx] != 1 + DRIFT_ROUTE_OK) continue;
      len[1 + k] = a.next_len[idx];
      off[1 + k] = a.next_off[idx];
    }
    const unsigned long long total = (unsigned long long)len[0] + len[1] + len[2];
    unsigned long long base = 0;
    if (lane == 0 && total) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const bool fits = base + total <= cap;
    if (!fits && lane == 0) atomicOr(&ctr->overflow, kOvfArena);
    unsigned long long o = base;
    for (int r = 0; r < 3; ++r) {
      if (fits)
        for (int32_t k = lane; k < len[r]; k += 32) dst[o + k] = src[off[r] + k];
      if (lane == 0) {
        if (r == 0) {
          a.route_off[i] = len[0] && fits ? (int64_t)o : 0;
          if (!len[0]) a.route_len[i] = 0;  // the finished route is dropped
        } else if (len[r] && fits) {
          a.next_off[(int64_t)(r - 1) * a.npad + i] = (int64_t)o;
        }
      }
      o += (unsigned long long)len[r];
    }
  }
}

// Trip-record export: length (metres, left-to-right fp64 sum) and, optionally, the links of the
// current route of each listed agent.
__global__ void k_gather_routes(int64_t n, const int32_t* __restrict__ agents, AgentArrays a,
                                const int32_t* __restrict__ arena, const double* __restrict__ link_len,
                                const int64_t* __restrict__ out_off, int32_t* __restrict__ out_links,
                                double* __restrict__ out_metres, int32_t* __restrict__ out_len) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = agents[r];
    const int64_t off = a.route_off[i];
    const int32_t len = a.route_len[i];
    double m = 0.0;
    for (int32_t k = 0; k < len; ++k) {
      const int32_t l = arena[off + k];
      m = __dadd_rn(m, link_len[l]);
      if (out_links) out_links[out_off[r] + k] = l;
    }
    out_metres[r] = m;
    if (out_len) out_len[r] = len;
  }
}

// Trip-record export by arena position (routes logged at departure, drift_drain_departures_routes): metres
// (left-to-right fp64 sum of the link lengths, data_logger.py:103-120) and, optionally, the links.
__global__ void k_gather_logged(int64_t n, const long long* __restrict__ off, const int32_t* __restrict__ len,
                                const int32_t* __restrict__ arena, const double* __restrict__ link_len,
                                const int64_t* __restrict__ out_off, int32_t* __restrict__ out_links,
                                double* __restrict__ out_metres) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double m = 0.0;
    for (int32_t k = 0; k < len[r]; ++k) {
      const int32_t l = arena[off[r] + k];
      m = __dadd_rn(m, link_len[l]);
      if (out_links) out_links[out_off[r] + k] = l;
    }
    out_metres[r] = m;
  }
}

// Waiting agents that will depart within the next n_ticks ticks (same Philox draws as the tick
// kernel) drop their prefetched routes so that the next planning batch recomputes them on the
// current travel times.
__global__ void k_invalidate_prefetch(AgentArrays a, uint64_t seed, int32_t tick0, int32_t n_ticks,
                                      const double* __restrict__ p_dep) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    if (a.state[i] != kWaiting || a.chain_cnt[i] == 0 || a.next_status[i] == 0) continue;
    for (int k = 0; k < n_ticks; ++k) {
      if (philox_uniform53(seed, a.base + (uint32_t)i, (uint32_t)(tick0 + k)) < p_dep[k]) {
        a.next_status[i] = 0;
        a.next_status[a.npad + i] = 0;
        break;
      }
    }
  }
}

__global__ void k_store_prefetch(int64_t Q, const int32_t* __restrict__ q_agent, const int64_t* __restrict__ q_off,
                                 const int32_t* __restrict__ q_len, const int32_t* __restrict__ q_status,
                                 AgentArrays a) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t code = q_agent[q];
    const int64_t idx = (int64_t)(code & 1) * a.npad + (code >> 1);
    a.next_off[idx] = q_off[q];
    a.next_len[idx] = q_len[q];
    a.next_status[idx] = 1 + q_status[q];
  }
}

struct NodeRec;
__global__ void k_init_noderecs(NodeRec* __restrict__ p, int64_t n);
constexpr int32_t kStatusTie = -3;  // batched router: the shortest path may not be unique -> NetworkX-exact fallback

// =====================================================================================
// K5b: batched multi-source SSSP -- one reverse shortest-path tree per destination.
//
// DRIFT's s-t models concentrate trips on few destinations (activity pools, hubs, zones:
// lib/st_selection.py), and the reference's travel_time never changes, so a tree
// "first link of the shortest path from every node to destination t" is built once per
// distinct t and cached in HBM; a route is then a pointer walk.  Trees are built many at a
// time: one thread block per tree (near-far / delta-stepping worklists, block-level
// barriers only), blocks pull jobs from a device queue so no host round trip is needed.
// Distances are accumulated from the destination side (dist[u] = dist[v] + w(u,v)), so the
// path equals NetworkX's wherever the shortest path is unique; cost parity 1e-9 relative.
// =====================================================================================

// 16 bytes per node and resident search: label, near-list stamp, and one word that holds the far-list flag (bit 31) and
// the node's potential for the running goal-directed search as a non-negative float ROUNDED DOWN (bits 0..30).  A
// relaxation used to fetch the 128-byte landmark line of its head node every time (4 in-links = 4 fetches per node on
// a grid); now the first relaxation of a node computes the potential and the later ones find it in the sector they
// read anyway.  Rounding down keeps the potential admissible, which is all the bounded search's stopping rule needs
// (a node still in the far list has label + potential > T, so no path through it reaches a source below T); the
// labels themselves stay fp64 sums.  The throughput of large graphs follows the number of resident searches, i.e.
// the bytes of scratch per node: that is why the record is not wider.
struct __align__(16) NodeRec {
  unsigned long long dist;  // fp64 bit pattern (non-negative doubles order like u64)
  uint32_t stamp;           // near-list dedupe stamp (running wave number, only ever compared for equality)
  uint32_t potf;            // kFarBit | float bits of the cached potential, kPotUnset = not computed in this search
};
constexpr uint32_t kFarBit = 0x80000000u;
constexpr uint32_t kPotUnset = 0x7FC00000u;  // a NaN pattern: no potential has it
__device__ __forceinline__ void reset_noderec(NodeRec* r) {
  *reinterpret_cast<uint4*>(r) = make_uint4(0u, 0x7FF00000u, 0u, kPotUnset);  // +inf, stamp 0, no flag, no potential
}

struct TreeScratch {   // per resident block
  NodeRec* rec;              // [blocks][V]
  int32_t* lists;            // [blocks][5][V]: touched, farA, farB, nearA, nearB
  uint32_t* iter;            // [blocks] running stamp
};

// Landmark lower bounds for goal-directed bounded searches (ALT: A*, landmarks, triangle inequality;
// Goldberg & Harrelson 2005).  d(s,v) >= from[k][v] - from[k][s] and >= to[k][s] - to[k][v].
struct Landmarks {
  // [V][2K] interleaved per node: entries 0..K-1 = distance landmark k -> node, K..2K-1 = node -> landmark k
  // (one 128-byte line per node for K = 8, so that a potential can use ALL bounds at the price of one line)
  const double* tab;
  int32_t K;
};
constexpr int kMaxLandmarks = 16;

struct TreeCache {
  int32_t* slot_of;      // [V] destination node -> slot, -1 = not cached
  int32_t* slot_owner;   // [max_slots] destination held by the slot, -1 = free
  uint32_t* slot_stamp;  // [max_slots] last batch that used the slot (never evicted within that batch)
  int2* next_link;       // [slots][V] {first link towards the root (-1 = none), node at its other end}: one load per hop
  int32_t* ring_cursor;  // FIFO allocation cursor (device counter)
  int32_t max_slots;
  int32_t* job_dst;      // [max_slots] trees to build
  int32_t* job_slot;
  int32_t* n_jobs;
  int32_t* job_cursor;
  uint32_t batch;        // id of this batch
  // queries of the batch hanging off the job that builds their destination (bounded search)
  int32_t* slot_job;        // [max_slots] job that fills the slot ...
  uint32_t* slot_job_batch; // [max_slots] ... in this batch
  int32_t* job_qhead;       // [max_slots] head of the job's query list
  int32_t* job_cnt;         // [max_slots] queries on that list
  int32_t* q_next;          // [Q]
  int32_t bounded_max;      // jobs with <= this many queries search only until their sources are settled (0 = off)
  // Trees hang off either end of a trip: key t in [0, V) = reverse tree towards destination t, key
  // V + s = forward tree out of source s (return legs to scattered targets leave from few places:
  // depots, hubs, activity pools -- lib/st_selection.py).  cnt[key] = queries of this batch per key.
  int32_t* cnt;             // [2V], zero between batches
  // Popularity of a destination across batches (halved every 64 batches): with small batches (lookahead
  // planning) a mildly popular destination never collects bounded_max queries in ONE batch and would pay
  // a bounded search per query for ever; once pop[t] exceeds pop_thresh its tree is built and cached.
  int32_t* pop;             // [V]
  int32_t pop_thresh;
  uint8_t* q_side;          // [Q] 0 = answered from the destination's tree, 1 = from the source's
  int32_t V;
  int32_t sides;            // 1 = destination trees only, 2 = both
  int32_t detect_ties;      // 1 = flag queries whose shortest path may not be unique (kStatusTie)
};

constexpr int kMaxBounded = 8;

constexpr unsigned long long kInfBits = 0x7FF0000000000000ULL;
constexpr int kTreeThreadsMax = 256;
constexpr int kTreeBlocksPerSM = 8;  // 32 registers per thread: searches are latency-bound, residency is what pays

__global__ void k_pop_decay(int32_t* __restrict__ pop, int32_t V) {
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) pop[i] >>= 1;
}

__device__ __forceinline__ bool tree_query_ok(int32_t s, int32_t t, int32_t V) {
  return s >= 0 && t >= 0 && s < V && t < V && s != t;
}

// Pass 0: how many queries of this batch share each destination / each source.
__global__ void k_tree_count(int64_t Q, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, TreeCache tc,
                             const int32_t* __restrict__ n_dev) {
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = src[q], t = dst[q];
    if (!tree_query_ok(s, t, tc.V)) continue;
    atomicAdd(&tc.cnt[t], 1);
    atomicAdd(&tc.pop[t], 1);
    if (tc.sides > 1) atomicAdd(&tc.cnt[tc.V + s], 1);
  }
}

// Pass 1: choose the end of the trip whose tree answers the query -- a cached tree first (it pins its
// slot for this batch), else the end shared by more queries of the batch.
__global__ void k_tree_hits(int64_t Q, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, TreeCache tc,
                            const int32_t* __restrict__ n_dev) {
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = src[q], t = dst[q];
    tc.q_side[q] = 0;
    if (!tree_query_ok(s, t, tc.V)) continue;
    int32_t slot = tc.slot_of[t];
    if (slot >= 0) {
      tc.slot_stamp[slot] = tc.batch;
      continue;
    }
    if (tc.sides < 2) continue;
    slot = tc.slot_of[tc.V + s];
    if (slot >= 0) {
      tc.slot_stamp[slot] = tc.batch;
      tc.q_side[q] = 1;
      continue;
    }
    const int32_t cs = tc.cnt[tc.V + s], cd = tc.cnt[t];
    if (cs > cd && cs > tc.bounded_max) tc.q_side[q] = 1;
  }
}

// Pass 2: every other key gets a slot from the FIFO ring (evicting the oldest tree that is not used
// by this batch) and becomes a build job.
__global__ void k_tree_requests(int64_t Q, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                TreeCache tc, Counters* ctr, const int32_t* __restrict__ n_dev) {
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    if (!tree_query_ok(src[q], dst[q], tc.V)) continue;
    const int32_t t = tc.q_side[q] ? tc.V + src[q] : dst[q];
    if (atomicCAS(&tc.slot_of[t], -1, -2) != -1) continue;  // cached, or another thread is allocating it
    const bool may_bound = t < tc.V && tc.bounded_max > 0 && tc.cnt[t] <= tc.bounded_max;
    int32_t s = -1;
    if (!(may_bound && tc.pop[t] <= tc.pop_thresh)) {  // hot in this batch, or popular over the last batches: keep a tree
      for (int tries = 0; tries < tc.max_slots; ++tries) {
        const int32_t c = (int32_t)((uint32_t)atomicAdd(tc.ring_cursor, 1) % (uint32_t)tc.max_slots);
        if (atomicExch(&tc.slot_stamp[c], tc.batch) != tc.batch) {  // not used by this batch: take it
          s = c;
          break;
        }
      }
    }
    if (s < 0 && may_bound) {
      // cold destination (or no free slot for a merely popular one): a bounded search answers its few
      // queries and keeps no tree -> no slot; slot_of[t] = -(job + 3) until the job is done
      const int32_t j = atomicAdd(tc.n_jobs, 1);
      tc.job_dst[j] = t;
      tc.job_slot[j] = -1;
      tc.job_qhead[j] = -1;
      tc.job_cnt[j] = 0;
      tc.slot_of[t] = -(j + 3);
      continue;
    }
    if (s < 0) {  // more distinct hot keys in one batch than slots
      tc.slot_of[t] = -1;
      atomicOr(&ctr->overflow, kOvfStage);
      continue;
    }
    const int32_t old = tc.slot_owner[s];
    if (old >= 0) tc.slot_of[old] = -1;  // evicted (old != any key of this batch: its slot is stamped)
    tc.slot_owner[s] = t;
    const int32_t j = atomicAdd(tc.n_jobs, 1);
    tc.job_dst[j] = t;
    tc.job_slot[j] = s;
    tc.job_qhead[j] = -1;
    tc.job_cnt[j] = 0;
    tc.slot_job[s] = j;
    tc.slot_job_batch[s] = tc.batch;
    tc.slot_of[t] = s;
  }
}

// Pass 3: hang every query whose tree is being built in this batch on that job's list; leave cnt[]
// zero for the next batch.
__global__ void k_link_queries(int64_t Q, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                               TreeCache tc, const int32_t* __restrict__ n_dev) {
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    if (!tree_query_ok(src[q], dst[q], tc.V)) continue;
    tc.cnt[dst[q]] = 0;
    if (tc.sides > 1) tc.cnt[tc.V + src[q]] = 0;
    const int32_t t = tc.q_side[q] ? tc.V + src[q] : dst[q];
    const int32_t s = tc.slot_of[t];
    int32_t j;
    if (s <= -3)
      j = -(s + 3);  // bounded job of this batch
    else if (s >= 0 && tc.slot_job_batch[s] == tc.batch)
      j = tc.slot_job[s];
    else
      continue;
    tc.q_next[q] = atomicExch(&tc.job_qhead[j], (int32_t)q);
    atomicAdd(&tc.job_cnt[j], 1);
  }
}

__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v, unsigned long long* s_min) {
  for (int o = 16; o; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xffffffffu, v, o);
    v = x < v ? x : v;
  }
  if ((threadIdx.x & 31) == 0) atomicMin(s_min, v);
  __syncthreads();
  return *s_min;
}

// First entry e of [e0, e1) (CSR order) whose hop is tight: rec[col[e]].dist + w[e] == du.  The loads of
// four entries are issued together (col/w, then the four labels), so a hop of the walk below costs
// three dependent memory levels instead of two per entry.  Returns the entry, its label in *d_next.
// A second entry whose hop is tight within kTieTol (relative) means the shortest path may not be unique -- or
// unique only in one summation order: NetworkX sums forward labels from the source and backward labels from
// the target (nx:weighted.py:2440-2452), this builder from the tree's root.  Such queries are re-run by the
// NetworkX-exact router (k_route_nx), which reproduces the reference's tie-breaking.
constexpr double kTieTol = 1e-12;

__device__ __forceinline__ bool near_tight(double s, double du) { return fabs(s - du) <= kTieTol * du; }

// Same, but scans every entry: returns the first exactly tight one and sets *tie when another entry is tight
// within kTieTol.  Loads are issued four entries at a time like find_tight_entry.
__device__ __forceinline__ int32_t find_tight_entry_ties(int32_t e0, int32_t e1, unsigned long long du,
                                                         const NodeRec* __restrict__ rec,
                                                         const int32_t* __restrict__ col, const double* __restrict__ w,
                                                         unsigned long long* d_next, bool* tie) {
  int32_t first = -1;
  int near = 0;
  const double dud = __longlong_as_double((long long)du);
  for (int32_t eb = e0; eb < e1; eb += 4) {
    int32_t c[4];
    double ww[4];
    unsigned long long d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int32_t e = eb + k < e1 ? eb + k : e1 - 1;
      c[k] = col[e];
      ww[k] = w[e];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = rec[c[k]].dist;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (eb + k >= e1 || d[k] == kInfBits) continue;
      const double s = __dadd_rn(__longlong_as_double((long long)d[k]), ww[k]);
      if (first < 0 && (unsigned long long)__double_as_longlong(s) == du) {
        first = eb + k;
        *d_next = d[k];
        ++near;
      } else if (near_tight(s, dud)) {
        ++near;
      }
    }
  }
  if (near > 1) *tie = true;
  return first;
}

__device__ __forceinline__ int32_t find_tight_entry(int32_t e0, int32_t e1, unsigned long long du,
                                                    const NodeRec* __restrict__ rec, const int32_t* __restrict__ col,
                                                    const double* __restrict__ w, unsigned long long* d_next) {
  for (int32_t eb = e0; eb < e1; eb += 4) {
    int32_t c[4];
    double ww[4];
    unsigned long long d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int32_t e = eb + k < e1 ? eb + k : e1 - 1;
      c[k] = col[e];
      ww[k] = w[e];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = rec[c[k]].dist;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (eb + k >= e1 || d[k] == kInfBits) continue;
      const double s = __dadd_rn(__longlong_as_double((long long)d[k]), ww[k]);
      if ((unsigned long long)__double_as_longlong(s) == du) {
        *d_next = d[k];
        return eb + k;
      }
    }
  }
  return -1;
}

// Route of one query of a bounded search, straight from dist[]: at x follow the first forward entry
// (CSR order) that is tight; every node on the way is already final.  One pass: the links go to `tmp`
// (scratch of the finished search) and are copied into the arena once their number is known; routes
// longer than `tmp_cap` are walked a second time instead.
__device__ __forceinline__ void route_from_labels(const GraphArrays& g, const double* __restrict__ w_fwd,
                                                  const NodeRec* __restrict__ rec, int32_t sn, int32_t dest,
                                                  int32_t* __restrict__ tmp, int32_t tmp_cap,
                                                  int32_t* __restrict__ arena, unsigned long long arena_cap,
                                                  Counters* ctr, const double* __restrict__ link_cost, int32_t q,
                                                  const RouteOut& out, bool detect_ties) {
  const int64_t V = g.V;
  int32_t status = DRIFT_ROUTE_NOPATH, hops = 0;
  int64_t off = 0;
  double cost = 0.0;
  bool tie = false;
  if (sn == dest) {
    status = DRIFT_ROUTE_TRIVIAL;
  } else if (sn >= 0 && sn < V && rec[sn].dist != kInfBits) {
    int k = 0;
    int32_t x = sn;
    unsigned long long dx = rec[sn].dist;
    while (x != dest) {
      unsigned long long dn = 0;
      const int32_t e = detect_ties
                            ? find_tight_entry_ties(g.fwd_rowptr[x], g.fwd_rowptr[x + 1], dx, rec, g.fwd_col, w_fwd, &dn, &tie)
                            : find_tight_entry(g.fwd_rowptr[x], g.fwd_rowptr[x + 1], dx, rec, g.fwd_col, w_fwd, &dn);
      if (e < 0 || k > V) {
        k = -1;
        break;
      }
      if (k < tmp_cap) tmp[k] = g.fwd_link[e];
      ++k;
      x = g.fwd_col[e];
      dx = dn;
    }
    if (tie) {
      status = kStatusTie;  // answered by the NetworkX-exact router afterwards
    } else if (k > 0) {
      hops = k;
      const unsigned long long o = atomicAdd(&ctr->arena_cursor, (unsigned long long)hops);
      if (o + hops > arena_cap) {
        atomicOr(&ctr->overflow, kOvfArena);
      } else {
        off = (int64_t)o;
        if (hops <= tmp_cap) {
          for (int j = 0; j < hops; ++j) {
            const int32_t l = tmp[j];
            arena[off + j] = l;
            cost = __dadd_rn(cost, link_cost[l]);
          }
        } else {  // longer than the scratch: walk again
          x = sn;
          dx = rec[sn].dist;
          for (int j = 0; j < hops; ++j) {
            unsigned long long dn = 0;
            const int32_t e = find_tight_entry(g.fwd_rowptr[x], g.fwd_rowptr[x + 1], dx, rec, g.fwd_col, w_fwd, &dn);
            const int32_t l = g.fwd_link[e];
            arena[off + j] = l;
            cost = __dadd_rn(cost, link_cost[l]);
            x = g.fwd_col[e];
            dx = dn;
          }
        }
        status = DRIFT_ROUTE_OK;
      }
    }
  }
  out.status[q] = status;
  out.n_links[q] = status == DRIFT_ROUTE_OK ? hops : 0;
  out.off[q] = status == DRIFT_ROUTE_OK ? off : 0;
  out.cost[q] = cost;
}

// Jobs whose destination is wanted by at most tc.bounded_max queries of the batch ("cold"
// destinations) stop as soon as all their sources are settled, write those routes themselves
// (walking dist[] from the source) and do not keep a tree; the others build and cache the tree.
// kBlocksPerSM = 8 / 6 / 4: 32 / 40 / 64 registers per thread; the engine launches the widest build whose residency the
// scratch budget can fill (large graphs: memory, not the SM, limits the resident searches).  kFly = relaxations a
// thread keeps in flight; 2 was measured on the 4 M-node grid and changes nothing (the waves are short), so 1 is built.
template <int kBlocksPerSM, int kFly>
__global__ void __launch_bounds__(kTreeThreadsMax, kBlocksPerSM) k_build_trees(GraphArrays g, const double* __restrict__ w_bwd,
                                                              const double* __restrict__ w_fwd, double delta,
                                                              TreeCache tc, TreeScratch sc,
                                                              const int32_t* __restrict__ q_src,
                                                              int32_t* __restrict__ arena, unsigned long long arena_cap,
                                                              Counters* ctr, RouteOut out,
                                                              const double* __restrict__ link_cost, Landmarks lm,
                                                              int32_t pot_cache) {
  __shared__ int s_job, s_near_n, s_next_n, s_far_n, s_far2_n, s_done, s_touched, s_lm_on;
  __shared__ double s_ref[2 * kMaxLandmarks];
  __shared__ int s_e0[kTreeThreadsMax], s_off[kTreeThreadsMax], s_wsum[32];
  __shared__ double s_dv[kTreeThreadsMax];
  __shared__ int s_bsrc[kMaxBounded], s_bq[kMaxBounded];
  __shared__ unsigned long long s_min;
  const int64_t V = g.V;
  // per-node scratch packed into one 16-byte record: a relaxation touches one sector, not three
  NodeRec* rec = sc.rec + (int64_t)blockIdx.x * V;
  int32_t* touched = sc.lists + (int64_t)blockIdx.x * 5 * V;
  // the A / B halves of the far and near queues swap by parity (fewer live pointers than swapping them: registers)
  int fpar = 0, npar = 0;
#define FAR_Q(par) (touched + (int64_t)(1 + (par)) * V)
#define NEAR_Q(par) (touched + (int64_t)(3 + (par)) * V)
  uint32_t iter = sc.iter[blockIdx.x];
  const int n_jobs = *tc.n_jobs;
  // work counters: per thread and launch 32 bits are plenty; the three that only thread 0 keeps live in shared memory
  // (registers are what limits the residency of this kernel)
  uint32_t c_improved = 0;
  __shared__ unsigned long long s_c_iters, s_c_far, s_c_jobs, s_c_relaxed, s_c_expanded;
  if (threadIdx.x == 0) s_c_iters = s_c_far = s_c_jobs = s_c_relaxed = s_c_expanded = 0ULL;
  for (;;) {
    if (threadIdx.x == 0) s_job = atomicAdd(tc.job_cursor, 1);
    __syncthreads();
    const int job = s_job;
    __syncthreads();
    if (job >= n_jobs) break;
    const int32_t key = tc.job_dst[job];
    const bool out_tree = key >= g.V;  // forward tree out of a source: expand along forward entries
    const int32_t dest = out_tree ? key - g.V : key;
    const int32_t* __restrict__ x_rowptr = out_tree ? g.fwd_rowptr : g.bwd_rowptr;
    const int32_t* __restrict__ x_col = out_tree ? g.fwd_col : g.bwd_col;
    const double* __restrict__ x_w = out_tree ? w_fwd : w_bwd;
    const int32_t slot = tc.job_slot[job];
    int2* next_link = tc.next_link + (int64_t)(slot < 0 ? 0 : slot) * V;
    const int nb = min(tc.job_cnt[job], kMaxBounded);
    const bool bounded = slot < 0;  // decided when the job was created (k_tree_requests)
    if (threadIdx.x == 0) ++s_c_jobs;
    if (bounded && threadIdx.x == 0) {
      int k = 0;
      for (int32_t q = tc.job_qhead[job]; q >= 0 && k < kMaxBounded; q = tc.q_next[q], ++k) {
        s_bq[k] = q;
        s_bsrc[k] = q_src[q];
      }
    }
    // dist[] is +inf and infar[] 0 for every node here: each job resets what it touched when it ends
    if (threadIdx.x == 0) {
      s_near_n = 1;
      s_next_n = 0;
      s_far_n = 0;
      s_far2_n = 0;
      s_touched = 1;
      NEAR_Q(npar)[0] = dest;
      touched[0] = dest;
      rec[dest].dist = 0ULL;
      // goal direction for a single-source bounded job: ALT potentials, the source's row of the landmark table
      s_lm_on = 0;
      if (bounded && nb == 1 && lm.K > 0) {
        const int32_t sn = s_bsrc[0];
        if (sn >= 0 && sn < V) {
          for (int k = 0; k < 2 * lm.K; ++k) s_ref[k] = lm.tab[(int64_t)sn * (2 * lm.K) + k];
          s_lm_on = 1;
        }
      }
    }
    __syncthreads();
    const bool lm_on = s_lm_on != 0;
    const int lmK = lm.K;
    // feasible potential: lower bound on the remaining distance d(source, v), the maximum over ALL landmark
    // bounds d(s,v) >= from[k][v] - from[k][s] and d(s,v) >= to[k][s] - to[k][v] (a maximum of consistent
    // potentials is consistent).  With the two bounds that are best for the pair itself the search settles
    // about twice as many nodes (tools/search_space_study.py).  0 without landmarks.
    auto pot = [&](int32_t v) -> double {
      double h = 0.0;
      if (lm_on) {
        const double* __restrict__ row = lm.tab + (int64_t)v * (2 * lmK);
#pragma unroll 2
        for (int k = 0; k < lmK; ++k) {
          const double b0 = row[k] - s_ref[k];
          const double b1 = s_ref[lmK + k] - row[lmK + k];
          if (b0 > h && b0 < 1e300) h = b0;  // inf / NaN (unreachable landmark) never pass
          if (b1 > h && b1 < 1e300) h = b1;
        }
      }
      return h;
    };
    // potential of a node that has been relaxed in this search (far-list scans): from its record when cached
    const bool cached = lm_on && pot_cache != 0;
    auto pot_seen = [&](int32_t v) -> double {
      return cached ? (double)__uint_as_float(__ldcg(&rec[v].potf) & ~kFarBit) : pot(v);
    };
    // goal-directed keys cluster just below d(source, dest): finer buckets keep the search inside the corridor
    const double jdelta = lm_on ? delta * 0.25 : delta;
    double T_old = -1.0, T = pot(dest) + jdelta;
    __syncthreads();
    for (;;) {
      const int n = s_near_n;
      if (n == 0) {
        const int nf = s_far_n;
        if (nf == 0) break;
        if (bounded) {  // every node with dist <= T is final: stop once all sources are
          if (threadIdx.x == 0) {
            int done = 1;
            for (int k = 0; k < nb; ++k) {
              const int32_t sn = s_bsrc[k];
              if (sn >= 0 && sn < V && __longlong_as_double((long long)rec[sn].dist) > T) done = 0;
            }
            s_done = done;
          }
          __syncthreads();
          const int done = s_done;
          __syncthreads();
          if (done) break;
        }
        // next threshold: skip empty buckets
        if (threadIdx.x == 0) s_min = kInfBits;
        __syncthreads();
        unsigned long long m = kInfBits;
        const int32_t* __restrict__ far_in = FAR_Q(fpar);
        int32_t* __restrict__ far_out = FAR_Q(fpar ^ 1);
        int32_t* __restrict__ near_out = NEAR_Q(npar);
        for (int i = threadIdx.x; i < nf; i += blockDim.x) {
          const int32_t u = far_in[i];
          const double key = __dadd_rn(__longlong_as_double((long long)rec[u].dist), pot_seen(u));
          const unsigned long long kb = (unsigned long long)__double_as_longlong(key);
          m = kb < m ? kb : m;
        }
        const double minfar = __longlong_as_double((long long)block_min_u64(m, &s_min));
        if (threadIdx.x == 0) s_c_far += (unsigned long long)nf;
        T_old = T;
        T = (minfar > T ? minfar : T) + jdelta;
        for (int i = threadIdx.x; i < nf; i += blockDim.x) {
          const int32_t u = far_in[i];
          const double d = __dadd_rn(__longlong_as_double((long long)rec[u].dist), pot_seen(u));
          if (d <= T_old) {
            rec[u].potf &= ~kFarBit;  // already expanded as a near node (one thread per far entry, no relaxation runs)
          } else if (d <= T) {
            rec[u].potf &= ~kFarBit;
            near_out[atomicAdd(&s_near_n, 1)] = u;
          } else {
            far_out[atomicAdd(&s_far2_n, 1)] = u;  // <= nf entries
          }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          s_far_n = s_far2_n;
          s_far2_n = 0;
        }
        fpar ^= 1;
        __syncthreads();
        continue;
      }
      if (++iter == 0u) iter = 1u;  // 0 is the stamp of a reset record
      if (threadIdx.x == 0) ++s_c_iters;
      // Edge-parallel expansion: the block loads a chunk of frontier nodes, scans their degrees and
      // then every thread relaxes ONE edge, so the dependent-load chain of a wave does not grow with
      // the degree (a wave of a road network holds few nodes: latency, not bandwidth, is the cost).
      const int32_t* __restrict__ near_in = NEAR_Q(npar);
      for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int deg = 0;
        if (i < n) {
          const int32_t v = near_in[i];
          const int32_t e0 = x_rowptr[v];
          deg = x_rowptr[v + 1] - e0;
          s_e0[threadIdx.x] = e0;
          s_dv[threadIdx.x] = __longlong_as_double((long long)rec[v].dist);
        }
        // block-wide exclusive scan of the degrees
        int incl = deg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int x = __shfl_up_sync(0xffffffffu, incl, o);
          if ((threadIdx.x & 31) >= o) incl += x;
        }
        if ((threadIdx.x & 31) == 31) s_wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
          const int nw = (blockDim.x + 31) >> 5;
          int wv = threadIdx.x < nw ? s_wsum[threadIdx.x] : 0;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, wv, o);
            if (threadIdx.x >= o) wv += x;
          }
          s_wsum[threadIdx.x] = wv;  // inclusive warp totals
        }
        __syncthreads();
        const int wbase = (threadIdx.x >> 5) ? s_wsum[(threadIdx.x >> 5) - 1] : 0;
        s_off[threadIdx.x] = wbase + incl - deg;
        const int total = s_wsum[((blockDim.x + 31) >> 5) - 1];
        const int cnt = min((int)blockDim.x, n - base);
        if (threadIdx.x == 0) {  // the block's totals of this chunk: nodes expanded, links relaxed
          s_c_expanded += (unsigned long long)cnt;
          s_c_relaxed += (unsigned long long)total;
        }
        __syncthreads();
        // kFly relaxations per thread in flight, in stages, so that their load chains (link -> head record ->
        // landmark line -> atomics) overlap
        for (int k0 = threadIdx.x; k0 < total; k0 += kFly * blockDim.x) {
          bool on[kFly];
          int32_t u[kFly];
          double cand[kFly], pu[kFly];
          unsigned long long cb[kFly], old[kFly];
#pragma unroll
          for (int j = 0; j < kFly; ++j) {  // stage 1: owner of the link, head node, candidate label
            const int k = k0 + j * (int)blockDim.x;
            on[j] = k < total;
            const int kk = on[j] ? k : k0;
            int lo = 0, hi = cnt - 1;  // owner: largest index with s_off[] <= kk
            while (lo < hi) {
              const int mid = (lo + hi + 1) >> 1;
              if (s_off[mid] <= kk) lo = mid; else hi = mid - 1;
            }
            const int32_t e = s_e0[lo] + (kk - s_off[lo]);
            u[j] = x_col[e];
            cand[j] = __dadd_rn(s_dv[lo], x_w[e]);
            cb[j] = (unsigned long long)__double_as_longlong(cand[j]);
          }
          if (cached) {
            // the node's sector first (L2, never a stale L1 line: the records are reset between jobs): label and
            // cached potential; the landmark line only on the node's first relaxation, the atomic only if it can win
#pragma unroll
            for (int j = 0; j < kFly; ++j) {  // stage 2: head records
              const ulonglong2 head = __ldcg(reinterpret_cast<const ulonglong2*>(&rec[u[j]]));
              old[j] = head.x;
              const uint32_t pf = (uint32_t)(head.y >> 32) & ~kFarBit;
              pu[j] = pf == kPotUnset ? -1.0 : (double)__uint_as_float(pf);
            }
#pragma unroll
            for (int j = 0; j < kFly; ++j) {  // stage 3: first relaxation of a node: its potential
              if (pu[j] < 0.0) {
                const float pf = __double2float_rd(pot(u[j]));
                pu[j] = (double)pf;
                // only the first store lands (the far flag is set after it, by a thread that has seen a potential)
                atomicCAS(&rec[u[j]].potf, kPotUnset, __float_as_uint(pf));
              }
            }
#pragma unroll
            for (int j = 0; j < kFly; ++j)  // stage 4: labels
              if (on[j] && cb[j] < old[j]) old[j] = atomicMin(&rec[u[j]].dist, cb[j]);
          } else {
#pragma unroll
            for (int j = 0; j < kFly; ++j) pu[j] = pot(u[j]);  // issued before the atomic: one dependent level less
#pragma unroll
            for (int j = 0; j < kFly; ++j) old[j] = on[j] ? atomicMin(&rec[u[j]].dist, cb[j]) : 0ULL;
          }
#pragma unroll
          for (int j = 0; j < kFly; ++j) {  // stage 5: queues
            if (on[j] && cb[j] < old[j]) {
              ++c_improved;
              if (old[j] == kInfBits) touched[atomicAdd(&s_touched, 1)] = u[j];  // exactly one thread sees the first label
              if (__dadd_rn(cand[j], pu[j]) <= T) {
                if (atomicExch(&rec[u[j]].stamp, iter) != iter) NEAR_Q(npar ^ 1)[atomicAdd(&s_next_n, 1)] = u[j];
              } else if ((atomicOr(&rec[u[j]].potf, kFarBit) & kFarBit) == 0u) {
                FAR_Q(fpar)[atomicAdd(&s_far_n, 1)] = u[j];  // at most one far entry per node: <= V entries
              }
            }
          }
        }
        __syncthreads();  // s_e0 / s_dv / s_off are reused by the next chunk
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        s_near_n = s_next_n;
        s_next_n = 0;
      }
      npar ^= 1;
      __syncthreads();
    }
    if (bounded) {
      // routes of the few sources, straight from dist[] (route_from_labels); the far list of the finished
      // search is the scratch the links are collected in
      if (threadIdx.x < nb) {
        const int32_t tcap = (int32_t)(V / kMaxBounded);
        route_from_labels(g, w_fwd, rec, s_bsrc[threadIdx.x], dest, FAR_Q(fpar ^ 1) + (int64_t)threadIdx.x * tcap, tcap, arena,
                          arena_cap, ctr, link_cost, s_bq[threadIdx.x], out, tc.detect_ties != 0);
      }
      if (threadIdx.x == 0) tc.slot_of[dest] = -1;  // no tree was kept
      __syncthreads();
      for (int i = threadIdx.x; i < s_touched; i += blockDim.x) {  // leave the scratch clean for the next job
        reset_noderec(&rec[touched[i]]);
      }
      __syncthreads();
      continue;
    }
    // reverse tree: first link of the shortest path u -> dest = first forward entry (CSR order) that
    // attains rec[u].dist; forward tree: last link of the shortest path dest -> u = first backward entry
    {
      const int32_t* __restrict__ y_rowptr = out_tree ? g.bwd_rowptr : g.fwd_rowptr;
      const int32_t* __restrict__ y_col = out_tree ? g.bwd_col : g.fwd_col;
      const double* __restrict__ y_w = out_tree ? w_bwd : w_fwd;
      const int32_t* __restrict__ y_link = out_tree ? g.bwd_link : g.fwd_link;
      for (int64_t u = threadIdx.x; u < V; u += blockDim.x) {
        int2 nl = make_int2(-1, -1);
        const unsigned long long du = rec[u].dist;
        if (u != dest && du != kInfBits) {
          unsigned long long dn = 0;
          bool tie = false;
          const int32_t e = tc.detect_ties
                                ? find_tight_entry_ties(y_rowptr[u], y_rowptr[u + 1], du, rec, y_col, y_w, &dn, &tie)
                                : find_tight_entry(y_rowptr[u], y_rowptr[u + 1], du, rec, y_col, y_w, &dn);
          // bit 31 of the node field: more than one (nearly) tight hop leaves this node
          if (e >= 0) nl = make_int2(y_link[e], y_col[e] | (tie ? (int32_t)0x80000000 : 0));
        }
        next_link[u] = nl;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < s_touched; i += blockDim.x) {
      reset_noderec(&rec[touched[i]]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) sc.iter[blockIdx.x] = iter;
  unsigned long long w_improved = c_improved;
  for (int o = 16; o; o >>= 1) w_improved += __shfl_down_sync(0xffffffffu, w_improved, o);
  if ((threadIdx.x & 31) == 0 && w_improved) atomicAdd(&ctr->tree_improved, w_improved);
  if (threadIdx.x == 0 && s_c_expanded) {
    atomicAdd(&ctr->tree_relaxed, s_c_relaxed);
    atomicAdd(&ctr->tree_expanded, s_c_expanded);
  }
  if (threadIdx.x == 0 && s_c_jobs) atomicAdd(&ctr->tree_jobs, s_c_jobs);
  if (threadIdx.x == 0 && s_c_iters) {
    atomicAdd(&ctr->tree_iters, s_c_iters);
    atomicAdd(&ctr->tree_far_visits, s_c_far);
  }
}

#undef FAR_Q
#undef NEAR_Q

__global__ void k_init_noderecs(NodeRec* __restrict__ p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    reset_noderec(&p[i]);
  }
}

// Pointer walk over the cached trees: route of every query into the arena.  A reverse tree is walked
// from the source (links come out in travel order), a forward tree from the destination back to the
// source (links are written back to front); the cost is always summed in travel order.
__global__ void k_extract_paths(GraphArrays g, int64_t Q, const int32_t* __restrict__ n_dev,
                                const int32_t* __restrict__ src, const int32_t* __restrict__ dst, TreeCache tc,
                                int32_t* __restrict__ arena, unsigned long long arena_cap, Counters* ctr,
                                RouteOut out, const double* __restrict__ link_cost) {
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = src[q], t = dst[q];
    if (out.status[q] >= 0 || out.status[q] == kStatusTie) continue;  // answered (or left to the exact router) by a bounded search
    out.off[q] = 0;
    out.cost[q] = 0.0;
    out.n_links[q] = 0;
    if (s < 0 || t < 0 || s >= g.V || t >= g.V) {
      out.status[q] = DRIFT_ROUTE_NOPATH;
      continue;
    }
    if (s == t) {
      out.status[q] = DRIFT_ROUTE_TRIVIAL;
      continue;
    }
    const bool out_tree = tc.q_side[q] != 0;
    const int32_t slot = tc.slot_of[out_tree ? g.V + s : t];
    if (slot < 0) {
      out.status[q] = DRIFT_ROUTE_NOPATH;
      continue;
    }
    const int2* next_link = tc.next_link + (int64_t)slot * g.V;
    const int32_t from = out_tree ? t : s, to = out_tree ? s : t;
    int hops = 0;
    int32_t x = from;
    int32_t tie = 0;
    while (x != to) {
      const int2 en = next_link[x];
      if (en.x < 0 || hops > g.V) {
        hops = -1;
        break;
      }
      ++hops;
      tie |= en.y;
      x = en.y & 0x7fffffff;
    }
    if (hops < 0) {
      out.status[q] = DRIFT_ROUTE_NOPATH;
      continue;
    }
    if (tie < 0 && tc.detect_ties) {  // some node on the way has several (nearly) tight hops: the NetworkX-exact router decides
      out.status[q] = kStatusTie;
      continue;
    }
    const unsigned long long off = atomicAdd(&ctr->arena_cursor, (unsigned long long)hops);
    if (off + hops > arena_cap) {
      atomicOr(&ctr->overflow, kOvfArena);
      out.status[q] = DRIFT_ROUTE_NOPATH;
      continue;
    }
    double c = 0.0;
    x = from;
    if (!out_tree) {
      for (int k = 0; k < hops; ++k) {
        const int2 en = next_link[x];
        arena[off + k] = en.x;
        c = __dadd_rn(c, link_cost[en.x]);
        x = en.y & 0x7fffffff;
      }
    } else {
      for (int k = hops - 1; k >= 0; --k) {
        const int2 en = next_link[x];
        arena[off + k] = en.x;
        x = en.y & 0x7fffffff;
      }
      for (int k = 0; k < hops; ++k) c = __dadd_rn(c, link_cost[arena[off + k]]);
    }
    out.status[q] = DRIFT_ROUTE_OK;
    out.n_links[q] = hops;
    out.off[q] = (int64_t)off;
    out.cost[q] = c;
  }
}

// Queries the batched router left to the NetworkX-exact one (kStatusTie): their indices, compacted.
__global__ void k_collect_ties(int64_t Q, const int32_t* __restrict__ n_dev, const int32_t* __restrict__ status,
                               int32_t* __restrict__ list, int32_t* __restrict__ n_list, Counters* ctr) {
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    if (status[q] == kStatusTie) {
      list[atomicAdd(n_list, 1)] = (int32_t)q;
      atomicAdd(&ctr->n_tie_fallback, 1ULL);
    }
}

}  // namespace drift

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

// w_fwd / w_bwd: per-CSR-entry weights (static travel_time, or congestion-updated).
__global__ void __launch_bounds__(32) k_route_nx(GraphArrays g, const double* __restrict__ w_fwd,
                                                  const double* __restrict__ w_bwd, int64_t Q,
                                                  const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                  RouteScratch sc, int32_t* __restrict__ arena,
                                                  unsigned long long arena_cap, Counters* ctr, RouteOut out,
                                                  const double* __restrict__ link_cost) {
  const int worker = blockIdx.x * blockDim.x + threadIdx.x;
  if (worker >= sc.workers) return;
  const int64_t V = g.V;
  double* seen[2] = {sc.seen + ((int64_t)worker * 2 + 0) * V, sc.seen + ((int64_t)worker * 2 + 1) * V};
  int32_t* plink[2] = {sc.plink + ((int64_t)worker * 2 + 0) * V, sc.plink + ((int64_t)worker * 2 + 1) * V};
  uint32_t* mark[2] = {sc.mark + ((int64_t)worker * 2 + 0) * V, sc.mark + ((int64_t)worker * 2 + 1) * V};
  HeapItem* heap[2] = {sc.heap + ((int64_t)worker * 2 + 0) * sc.heap_cap,
                       sc.heap + ((int64_t)worker * 2 + 1) * sc.heap_cap};
  uint32_t epoch = sc.epoch[worker];

  for (int64_t q = worker; q < Q; q += sc.workers) {
    const int32_t s = src[q], t = dst[q];
    out.off[q] = 0;
    out.cost[q] = 0.0;
    out.n_links[q] = 0;
    if (s < 0 || t < 0 || s >= g.V || t >= g.V) {  // agent.py:87-91
      out.status[q] = DRIFT_ROUTE_NOPATH;
      continue;
    }
    if (s == t) {  // weighted.py:2393-2394
      out.status[q] = DRIFT_ROUTE_TRIVIAL;
      continue;
    }
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
      out.status[q] = DRIFT_ROUTE_NOPATH;
      continue;
    }
    if (!found) {  // weighted.py:2459 NetworkXNoPath
      out.status[q] = DRIFT_ROUTE_NOPATH;
      continue;
    }
    int nf = 0, nb = 0;
    for (int32_t x = meet; x != s; x = g.link_u[plink[0][x]]) ++nf;
    for (int32_t x = meet; x != t; x = g.link_v[plink[1][x]]) ++nb;
    const int n = nf + nb;
    const unsigned long long off = atomicAdd(&ctr->arena_cursor, (unsigned long long)n);
    if (off + n > arena_cap) {
      atomicOr(&ctr->overflow, kOvfArena);
      out.status[q] = DRIFT_ROUTE_NOPATH;
      continue;
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
    out.status[q] = DRIFT_ROUTE_OK;
    out.n_links[q] = n;
    out.off[q] = (int64_t)off;
    out.cost[q] = c;
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
                         int64_t* __restrict__ pend_off, int32_t* __restrict__ pend_len, int32_t pend_cap) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    if (q_len[q] <= 0) continue;
    const int p = atomicAdd(&ctr->n_pending, 1);
    if (p < pend_cap) {
      pend_agent[p] = agent[q];
      pend_off[p] = q_off[q];
      pend_len[p] = q_len[q];
    }
  }
}

// Device-driven demand: turn this tick's departure requests into (src, dst) queries
// (Agent.start_new_journey, agent.py:46-83: source = previous target, trip_count += 1).
__global__ void k_prepare_requests(AgentArrays a, Counters* ctr, const int32_t* __restrict__ dep_req,
                                   const int64_t* __restrict__ trip_off, const int32_t* __restrict__ trip_dst,
                                   int32_t* __restrict__ q_src, int32_t* __restrict__ q_dst,
                                   int32_t* __restrict__ q_agent) {
  const int n = ctr->n_dep_req;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int32_t i = dep_req[r];
    const int32_t tc = a.trip_count[i];
    const int32_t s = a.cur_node[i];
    const int32_t t = trip_dst[trip_off[i] + tc];
    a.trip_count[i] = tc + 1;
    a.cur_node[i] = t;  // agent.py:51: target is overwritten even if routing fails
    q_src[r] = s;
    q_dst[r] = t;
    q_agent[r] = i;
  }
}

__global__ void k_log_departures(Counters* ctr, int32_t tick, const int32_t* __restrict__ q_agent,
                                 const int32_t* __restrict__ q_src, const int32_t* __restrict__ q_dst,
                                 const int32_t* __restrict__ q_status, int32_t* __restrict__ log_agent,
                                 int32_t* __restrict__ log_tick, int32_t* __restrict__ log_src,
                                 int32_t* __restrict__ log_dst, int32_t* __restrict__ log_status,
                                 unsigned long long cap, unsigned long long drained) {
  const int n = ctr->n_dep_req;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const unsigned long long slot = atomicAdd(&ctr->n_departures, 1ULL);
    if (slot - drained < cap) {
      const unsigned long long k = slot % cap;
      log_agent[k] = q_agent[r];
      log_tick[k] = tick;
      log_src[k] = q_src[r];
      log_dst[k] = q_dst[r];
      log_status[k] = q_status[r];
    } else {
      atomicOr(&ctr->overflow, kOvfDepLog);
    }
  }
}

}  // namespace drift

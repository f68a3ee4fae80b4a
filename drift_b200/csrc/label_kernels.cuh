// drift_b200 -- hub labels over the contraction hierarchy: routing as a streaming lookup.
//
// Replaces, for static travel times, the per-departure search of lib/agent.py:94
// (nx.shortest_path(G, s, t, weight='travel_time')).  hierarchy.hpp orders the nodes; the label of a node is
// everything a search climbing from it can reach, with its distance:
//   forward  label F(v) = {(h, d(v -> h))  : h reachable from v over up arcs}
//   backward label B(v) = {(h, d(h -> v))  : h reaches v over down arcs}
// Every shortest path is an up-down path in the hierarchy, so d(s, t) = min over common hubs of F(s) and B(t).
// Labels are built on the device top-down, one level per launch (the label of v is the min-merge of the labels
// of its up neighbours, which sit on higher levels and are finished), kept sorted by hub id in HBM
// (16 B per entry) and never change: a query is the intersection of two sorted arrays -- coalesced streaming
// reads instead of a latency-bound wavefront search -- and the route is recovered by following the parent arcs
// through the labels and unpacking the shortcuts into link ids.  The route's cost is re-summed over the links
// in travel order, so links and cost equal NetworkX's wherever the shortest path is unique.
#pragma once
#include "common.cuh"
#include "route_kernels.cuh"

namespace drift {

struct HierSide {  // one side of the hierarchy (hierarchy.hpp: HierArcs)
  const int32_t* rowptr;
  const int32_t* other;
  const double* w;
  const int32_t* a;
  const int32_t* b;
  const int32_t* hops;
};

struct LabelSet {  // labels of one side: the label of v is entries [off[v], off[v] + cnt[v]) sorted by hub
  long long* off;
  int32_t* cnt;
  int32_t* hub;    // shared entry arrays (both sides allocate from one cursor)
  int32_t* par;    // first arc of the best climb towards the hub (index into the side's arc arrays), -1 = the node itself
  double* dist;
};

constexpr int kLabelThreads = 256;
constexpr int kUnpackStack = 1024;  // shortcut nesting is bounded by the number of hierarchy levels
constexpr int32_t kStatusPending = -2;

__device__ __forceinline__ int label_hash_slot(int32_t* keys, int H, int shift, int32_t hub, int* s_n) {
  int i = (int)(((uint32_t)hub * 2654435761u) >> shift);
  for (;;) {
    const int32_t k = keys[i];
    if (k == hub) return i;
    if (k == -1) {
      const int32_t old = atomicCAS(&keys[i], -1, hub);
      if (old == -1) {
        atomicAdd(s_n, 1);
        return i;
      }
      if (old == hub) return i;
    }
    i = (i + 1) & (H - 1);
  }
}

// Label of every node of one level: block per node.  Shared memory: open-addressing table of H = 2 * n_max
// slots (hub, fp64 distance as ordered bits, parent arc) + n_max sort keys.
__global__ void __launch_bounds__(kLabelThreads) k_build_labels(HierSide sd, LabelSet ls, const int32_t* __restrict__ order,
                                                              int32_t first, int32_t n_max, int32_t log2H,
                                                              unsigned long long* cursor, unsigned long long cap,
                                                              int32_t* flag) {
  extern __shared__ unsigned long long smem_u64[];
  const int H = 1 << log2H;
  const int shift = 32 - log2H;
  unsigned long long* s_dist = smem_u64;         // [H]
  unsigned long long* s_sort = s_dist + H;       // [n_max] (hub << 32 | slot)
  int32_t* s_key = (int32_t*)(s_sort + n_max);   // [H]
  int32_t* s_par = s_key + H;                    // [H]
  __shared__ int s_n, s_m;
  __shared__ unsigned long long s_base;
  const int32_t v = order[first + blockIdx.x];
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    s_key[i] = -1;
    s_dist[i] = kInfBits;
    s_par[i] = 0x7fffffff;
  }
  if (threadIdx.x == 0) {
    s_n = 0;
    s_m = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int slot = label_hash_slot(s_key, H, shift, v, &s_n);
    s_dist[slot] = 0ULL;
    s_par[slot] = -1;
  }
  __syncthreads();
  const int32_t e0 = sd.rowptr[v], e1 = sd.rowptr[v + 1];
  // pass 1: min distance per hub over all (arc, entry of the neighbour's label) candidates
  for (int32_t e = e0; e < e1; ++e) {
    const int32_t nb = sd.other[e];
    const double we = sd.w[e];
    const long long o = ls.off[nb];
    const int32_t n = ls.cnt[nb];
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
      if (*(volatile int*)&s_n > n_max) break;  // label too large for the table: flagged below
      const int32_t hub = ls.hub[o + i];
      const double d = __dadd_rn(we, ls.dist[o + i]);
      const int slot = label_hash_slot(s_key, H, shift, hub, &s_n);
      atomicMin(&s_dist[slot], (unsigned long long)__double_as_longlong(d));
    }
  }
  __syncthreads();
  if (s_n > n_max) {
    if (threadIdx.x == 0) {
      atomicOr(flag, 1);
      ls.off[v] = 0;
      ls.cnt[v] = 0;
    }
    return;
  }
  // pass 2: the parent of a hub is the smallest arc that attains its distance (deterministic)
  for (int32_t e = e0; e < e1; ++e) {
    const int32_t nb = sd.other[e];
    const double we = sd.w[e];
    const long long o = ls.off[nb];
    const int32_t n = ls.cnt[nb];
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const int32_t hub = ls.hub[o + i];
      const double d = __dadd_rn(we, ls.dist[o + i]);
      const int slot = label_hash_slot(s_key, H, shift, hub, &s_n);
      if ((unsigned long long)__double_as_longlong(d) == s_dist[slot]) atomicMin(&s_par[slot], e);
    }
  }
  __syncthreads();
  // sort the entries by hub id
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    if (s_key[i] != -1) s_sort[atomicAdd(&s_m, 1)] = ((unsigned long long)(uint32_t)s_key[i] << 32) | (uint32_t)i;
  __syncthreads();
  const int n = s_m;
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + threadIdx.x; i < P; i += blockDim.x) s_sort[i] = ~0ULL;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = s_sort[i], y = s_sort[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            s_sort[i] = y;
            s_sort[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x == 0) {
    const unsigned long long base = atomicAdd(cursor, (unsigned long long)n);
    s_base = base;
    if (base + n > cap) {
      atomicOr(flag, 2);
      ls.off[v] = 0;
      ls.cnt[v] = 0;
    } else {
      ls.off[v] = (long long)base;
      ls.cnt[v] = n;
    }
  }
  __syncthreads();
  const unsigned long long base = s_base;
  if (base + n > cap) return;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const unsigned long long k = s_sort[j];
    const int slot = (int)(uint32_t)(k & 0xffffffffu);
    ls.hub[base + j] = (int32_t)(k >> 32);
    ls.dist[base + j] = __longlong_as_double((long long)s_dist[slot]);
    ls.par[base + j] = s_par[slot];
  }
}

// position of `hub` in the sorted hub array [0, n), or -1
__device__ __forceinline__ int32_t label_find(const int32_t* __restrict__ hubs, int32_t n, int32_t hub) {
  int32_t lo = 0, hi = n;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (hubs[mid] < hub) lo = mid + 1; else hi = mid;
  }
  return (lo < n && hubs[lo] == hub) ? lo : -1;
}

// d(s, t) = min over the common hubs of F(s) and B(t): one warp per query, the shorter label is streamed and
// looked up in the longer one.  Ties go to the smaller hub id.  Leaves status = kStatusPending and the hub.
__global__ void __launch_bounds__(256) k_label_query(int64_t Q, const int32_t* __restrict__ n_dev,
                                                     const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                     LabelSet lf, LabelSet lb, int32_t V, int32_t* __restrict__ q_hub,
                                                     RouteOut out, unsigned long long* __restrict__ work) {
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long entries = 0;
  for (int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; q < n; q += warps) {
    const int32_t s = src[q], t = dst[q];
    int32_t status = kStatusPending;
    if (s < 0 || t < 0 || s >= V || t >= V) status = DRIFT_ROUTE_NOPATH;  // agent.py:87-91
    else if (s == t) status = DRIFT_ROUTE_TRIVIAL;                         // weighted.py:2393-2394
    unsigned long long best = ~0ULL;
    int32_t bhub = 0x7fffffff;
    if (status == kStatusPending) {
      const int32_t* ha = lf.hub + lf.off[s];
      const double* da = lf.dist + lf.off[s];
      int32_t na = lf.cnt[s];
      const int32_t* hb = lb.hub + lb.off[t];
      const double* db = lb.dist + lb.off[t];
      int32_t nb = lb.cnt[t];
      if (na > nb) {
        const int32_t* th = ha; ha = hb; hb = th;
        const double* td = da; da = db; db = td;
        const int32_t tn = na; na = nb; nb = tn;
      }
      entries += (unsigned long long)(na + nb);
      for (int32_t i = lane; i < na; i += 32) {
        const int32_t hub = ha[i];
        const int32_t j = label_find(hb, nb, hub);
        if (j >= 0) {
          const unsigned long long d = (unsigned long long)__double_as_longlong(__dadd_rn(da[i], db[j]));
          if (d < best || (d == best && hub < bhub)) {
            best = d;
            bhub = hub;
          }
        }
      }
      for (int o = 16; o; o >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int32_t oh = __shfl_xor_sync(0xffffffffu, bhub, o);
        if (ob < best || (ob == best && oh < bhub)) {
          best = ob;
          bhub = oh;
        }
      }
      if (bhub == 0x7fffffff) status = DRIFT_ROUTE_NOPATH;  // weighted.py:2459 NetworkXNoPath
    }
    if (lane == 0) {
      out.status[q] = status;
      out.n_links[q] = 0;
      out.off[q] = 0;
      out.cost[q] = 0.0;
      q_hub[q] = bhub;
    }
  }
  if (work && lane == 0 && entries) atomicAdd(work, entries);
}

// shortcut -> link ids, in travel order, into arena[pos...]; returns the next position (or -1: nesting too deep)
__device__ __forceinline__ long long unpack_arc(const HierSide& up, const HierSide& down, int side_is_up, int32_t arc,
                                                int32_t* __restrict__ arena, long long pos, int32_t* stack) {
  int sp = 0;
  stack[sp++] = arc * 2 + side_is_up;
  while (sp > 0) {
    const int32_t x = stack[--sp];
    const HierSide& sd = (x & 1) ? up : down;
    const int32_t i = x >> 1;
    const int32_t a = sd.a[i], b = sd.b[i];
    if (b < 0) {
      arena[pos++] = a;
    } else {
      if (sp + 2 > kUnpackStack) return -1;
      stack[sp++] = b * 2 + 1;  // second half: an up arc of the middle node ...
      stack[sp++] = a * 2 + 0;  // ... after the first half: a down arc of it
    }
  }
  return pos;
}

// Route of every pending query: climb from s to the hub through F, from t to the hub through B, unpack.
__global__ void __launch_bounds__(128) k_label_paths(int64_t Q, const int32_t* __restrict__ n_dev,
                                                     const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                     LabelSet lf, LabelSet lb, HierSide up, HierSide down,
                                                     const int32_t* __restrict__ q_hub, int32_t* __restrict__ arena,
                                                     unsigned long long arena_cap, Counters* ctr, RouteOut out,
                                                     const double* __restrict__ link_cost) {
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  int32_t stack[kUnpackStack];
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    if (out.status[q] != kStatusPending) continue;
    const int32_t s = src[q], t = dst[q], hub = q_hub[q];
    int32_t status = DRIFT_ROUTE_NOPATH;
    int hops = 0;
    bool ok = true;
    for (int32_t x = s; x != hub;) {
      const long long o = lf.off[x];
      const int32_t j = label_find(lf.hub + o, lf.cnt[x], hub);
      const int32_t a = j >= 0 ? lf.par[o + j] : -1;
      if (a < 0) {
        ok = false;
        break;
      }
      hops += up.hops[a];
      x = up.other[a];
    }
    for (int32_t y = t; ok && y != hub;) {
      const long long o = lb.off[y];
      const int32_t j = label_find(lb.hub + o, lb.cnt[y], hub);
      const int32_t a = j >= 0 ? lb.par[o + j] : -1;
      if (a < 0) {
        ok = false;
        break;
      }
      hops += down.hops[a];
      y = down.other[a];
    }
    long long off = 0;
    double cost = 0.0;
    if (ok && hops > 0) {
      const unsigned long long o0 = atomicAdd(&ctr->arena_cursor, (unsigned long long)hops);
      if (o0 + hops > arena_cap) {
        atomicOr(&ctr->overflow, kOvfArena);
      } else {
        off = (long long)o0;
        long long pos = off;
        for (int32_t x = s; x != hub && pos >= 0;) {
          const long long o = lf.off[x];
          const int32_t a = lf.par[o + label_find(lf.hub + o, lf.cnt[x], hub)];
          pos = unpack_arc(up, down, 1, a, arena, pos, stack);
          x = up.other[a];
        }
        long long end = off + hops;
        for (int32_t y = t; y != hub && pos >= 0;) {
          const long long o = lb.off[y];
          const int32_t a = lb.par[o + label_find(lb.hub + o, lb.cnt[y], hub)];
          end -= down.hops[a];
          if (unpack_arc(up, down, 0, a, arena, end, stack) < 0) pos = -1;
          y = down.other[a];
        }
        if (pos < 0) {
          atomicOr(&ctr->overflow, kOvfStage);
        } else {
          for (int k = 0; k < hops; ++k) cost = __dadd_rn(cost, link_cost[arena[off + k]]);  // travel order
          status = DRIFT_ROUTE_OK;
        }
      }
    }
    out.status[q] = status;
    out.n_links[q] = status == DRIFT_ROUTE_OK ? hops : 0;
    out.off[q] = status == DRIFT_ROUTE_OK ? off : 0;
    out.cost[q] = cost;
  }
}

}  // namespace drift

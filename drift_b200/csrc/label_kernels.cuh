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
  const int4* rec;  // {a, b, hops, hops of the first half}: what unpacking needs about an arc, in one 16-byte load
};

struct LabelSet {  // labels of one side: the label of v is entries [off[v], off[v] + cnt[v]) sorted by hub
  long long* off;
  int32_t* cnt;
  int32_t* hub;    // shared entry arrays (both sides allocate from one cursor)
  int32_t* par;    // first arc of the best climb towards the hub (index into the side's arc arrays), -1 = the node itself
  uint32_t* pnext; // entry of the same hub in the label of that arc's other end: the climb is a chain of entries
  double* dist;
};

constexpr int kLabelThreads = 256;
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
// slots (hub, fp64 distance as ordered bits, parent arc + parent entry) + n_max sort keys: 48 B x n_max.
__global__ void __launch_bounds__(kLabelThreads) k_build_labels(HierSide sd, LabelSet ls, const int32_t* __restrict__ order,
                                                              int32_t first, int32_t n_max, int32_t log2H,
                                                              unsigned long long* cursor, unsigned long long cap,
                                                              int32_t* flag) {
  extern __shared__ unsigned long long smem_u64[];
  const int H = 1 << log2H;
  const int shift = 32 - log2H;
  unsigned long long* s_dist = smem_u64;         // [H]
  unsigned long long* s_sort = s_dist + H;       // [n_max] (hub << 32 | slot)
  unsigned long long* s_par = s_sort + n_max;    // [H] (parent arc << 32 | entry of the hub in the neighbour's label)
  int32_t* s_key = (int32_t*)(s_par + H);        // [H]
  __shared__ int s_n, s_m;
  __shared__ unsigned long long s_base;
  const int32_t v = order[first + blockIdx.x];
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    s_key[i] = -1;
    s_dist[i] = kInfBits;
    s_par[i] = ~0ULL;
  }
  if (threadIdx.x == 0) {
    s_n = 0;
    s_m = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int slot = label_hash_slot(s_key, H, shift, v, &s_n);
    s_dist[slot] = 0ULL;  // the node itself: no parent (s_par stays ~0: par = -1)
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
      if ((unsigned long long)__double_as_longlong(d) == s_dist[slot])
        atomicMin(&s_par[slot], ((unsigned long long)(uint32_t)e << 32) | (unsigned long long)(uint32_t)(o + i));
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
    ls.par[base + j] = (int32_t)(s_par[slot] >> 32);  // ~0 -> -1 for the node itself
    ls.pnext[base + j] = (uint32_t)(s_par[slot] & 0xffffffffu);
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
                                                     LabelSet lf, LabelSet lb, int32_t V, uint2* __restrict__ q_ent,
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
    uint32_t ea = 0, eb = 0;  // entries of the best hub in the two labels
    bool swapped = false;
    if (status == kStatusPending) {
      long long oa = lf.off[s], ob = lb.off[t];
      int32_t na = lf.cnt[s], nb = lb.cnt[t];
      if (na > nb) {  // stream the shorter label, search the longer one
        const long long to = oa; oa = ob; ob = to;
        const int32_t tn = na; na = nb; nb = tn;
        swapped = true;
      }
      const int32_t* ha = lf.hub + oa;
      const double* da = lf.dist + oa;
      const int32_t* hb = lf.hub + ob;
      const double* db = lf.dist + ob;
      entries += (unsigned long long)(na + nb);
      for (int32_t i = lane; i < na; i += 32) {
        const int32_t hub = ha[i];
        const int32_t j = label_find(hb, nb, hub);
        if (j >= 0) {
          const unsigned long long d = (unsigned long long)__double_as_longlong(__dadd_rn(da[i], db[j]));
          if (d < best || (d == best && hub < bhub)) {
            best = d;
            bhub = hub;
            ea = (uint32_t)(oa + i);
            eb = (uint32_t)(ob + j);
          }
        }
      }
      for (int o = 16; o; o >>= 1) {
        const unsigned long long xb = __shfl_xor_sync(0xffffffffu, best, o);
        const int32_t xh = __shfl_xor_sync(0xffffffffu, bhub, o);
        const uint32_t xa = __shfl_xor_sync(0xffffffffu, ea, o), xe = __shfl_xor_sync(0xffffffffu, eb, o);
        if (xb < best || (xb == best && xh < bhub)) {
          best = xb;
          bhub = xh;
          ea = xa;
          eb = xe;
        }
      }
      if (bhub == 0x7fffffff) status = DRIFT_ROUTE_NOPATH;  // weighted.py:2459 NetworkXNoPath
    }
    if (lane == 0) {
      out.status[q] = status;
      out.n_links[q] = 0;
      out.off[q] = 0;
      out.cost[q] = 0.0;
      q_ent[q] = swapped ? make_uint2(eb, ea) : make_uint2(ea, eb);  // (entry in F(s), entry in B(t))
    }
  }
  if (work && lane == 0 && entries) atomicAdd(work, entries);
}

// shortcut -> link ids, in travel order, into arena[pos...] by ONE thread (explicit stack); returns the next
// position, or -1 when the nesting is deeper than the stack
__device__ __forceinline__ long long unpack_arc(const HierSide& up, const HierSide& down, int32_t code,
                                                int32_t* __restrict__ arena, long long pos, int32_t* stack, int cap) {
  int sp = 0;
  stack[sp++] = code;
  while (sp > 0) {
    const int32_t x = stack[--sp];
    const int4 r = ((x & 1) ? up : down).rec[x >> 1];
    if (r.y < 0) {
      arena[pos++] = r.x;
    } else {
      if (sp + 2 > cap) return -1;
      stack[sp++] = r.y * 2 + 1;  // second half: an up arc of the middle node ...
      stack[sp++] = r.x * 2 + 0;  // ... after the first half: a down arc of it
    }
  }
  return pos;
}

constexpr int kPathWarps = 4;      // warps (= queries in flight) per block
constexpr int kPathItems = 1024;   // work items per warp: (arc, offset in the route); a climb has < kPathItems / 2 arcs
constexpr int kPathSerial = 96;    // private stack of the serial fallback

// Route of every pending query, one WARP per query:
//   1. lanes 0 / 1 follow the chain of label entries from F(s) / B(t) up to the hub (one dependent load per climb
//      step: an entry names its arc and the entry of the same hub one node further up) and list the arcs;
//   2. the warp sums the arcs' link counts (prefix sum = where each arc's links go), one lane reserves the arena;
//   3. the arcs are unpacked in parallel: the list is a stack of (arc, offset) items, every round the top 32 items
//      are expanded -- a link is written, a shortcut is replaced by its two halves (the first keeps the offset, the
//      second starts `hops of the first` later: one 16-byte record per arc holds all of that);
//   4. the cost is re-summed over the links in travel order (loads in parallel, adds in order).
__global__ void __launch_bounds__(kPathWarps * 32) k_label_paths(int64_t Q, const int32_t* __restrict__ n_dev,
                                                                 LabelSet lf, LabelSet lb, HierSide up, HierSide down,
                                                                 const uint2* __restrict__ q_ent,
                                                                 int32_t* __restrict__ arena,
                                                                 unsigned long long arena_cap, Counters* ctr,
                                                                 RouteOut out, const double* __restrict__ link_cost) {
  __shared__ int2 s_items[kPathWarps][kPathItems];
  const int64_t n = n_dev ? (int64_t)*n_dev : Q;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int2* items = s_items[w];
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; q < n; q += warps) {
    if (out.status[q] != kStatusPending) continue;  // warp-uniform
    const uint2 ent = q_ent[q];
    // 1. climb: lane 0 the forward chain (arcs in travel order, written upwards from item 0), lane 1 the backward
    //    chain (arcs in reverse travel order, written downwards from the last item: read upwards = travel order)
    int cnt = 0;
    if (lane < 2) {
      const LabelSet& ls = lane == 0 ? lf : lb;
      uint32_t e = lane == 0 ? ent.x : ent.y;
      for (;;) {
        const int32_t a = ls.par[e];
        if (a < 0) break;  // the hub's own entry
        const uint32_t nx = ls.pnext[e];
        if (cnt < kPathItems / 2) items[lane == 0 ? cnt : kPathItems - 1 - cnt] = make_int2(a * 2 + (lane == 0 ? 1 : 0), 0);
        ++cnt;
        e = nx;
      }
    }
    const int nf = __shfl_sync(0xffffffffu, cnt, 0), nb = __shfl_sync(0xffffffffu, cnt, 1);
    __syncwarp();
    int32_t status = DRIFT_ROUTE_NOPATH;
    int hops = 0;
    long long off = 0;
    double cost = 0.0;
    if (nf > kPathItems / 2 || nb > kPathItems / 2 || nf + nb == 0) {
      if (lane == 0 && nf + nb) atomicOr(&ctr->overflow, kOvfStage);  // cannot happen below kPathItems / 2 levels
    } else {
      // 2. compact to [0, total) in travel order (an item never moves to a higher index, chunks go upwards: in place),
      //    with the links before each arc (exclusive prefix sum of the arcs' link counts)
      const int total = nf + nb;
      const int b0 = kPathItems - nb;  // first backward item in travel order
      int run = 0;
      for (int c0 = 0; c0 < total; c0 += 32) {
        const int k = c0 + lane;
        int2 it = make_int2(-1, 0);
        if (k < total) it = k < nf ? items[k] : items[b0 + (k - nf)];
        int h = 0;
        if (it.x >= 0) h = ((it.x & 1) ? up : down).rec[it.x >> 1].z;
        int incl = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += y;
        }
        it.y = run + incl - h;
        run += __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        if (k < total) items[k] = it;
      }
      hops = run;
      __syncwarp();
      unsigned long long o0 = 0;
      if (lane == 0) o0 = atomicAdd(&ctr->arena_cursor, (unsigned long long)hops);
      o0 = __shfl_sync(0xffffffffu, o0, 0);
      if (o0 + hops > arena_cap) {
        if (lane == 0) atomicOr(&ctr->overflow, kOvfArena);
      } else {
        off = (long long)o0;
        __syncwarp();
        // 3. parallel expansion
        int top = total;
        bool bad = false;
        while (top > 0) {
          const int take = min(top, 32);
          int2 it = make_int2(-1, 0);
          if (lane < take) it = items[top - 1 - lane];
          top -= take;
          __syncwarp();
          int4 r = make_int4(0, -1, 0, 0);
          if (it.x >= 0) r = ((it.x & 1) ? up : down).rec[it.x >> 1];
          const bool inner = it.x >= 0 && r.y >= 0;
          if (it.x >= 0 && !inner) arena[off + it.y] = r.x;
          const unsigned m = __ballot_sync(0xffffffffu, inner);
          const int push = 2 * __popc(m);
          if (top + push > kPathItems) {
            // no room for both halves of every shortcut of this round: these lanes finish theirs serially
            if (inner) {
              int32_t st[kPathSerial];
              if (unpack_arc(up, down, it.x, arena, off + it.y, st, kPathSerial) < 0) bad = true;
            }
          } else if (inner) {
            const int slot = top + 2 * __popc(m & ((1u << lane) - 1u));
            items[slot] = make_int2(r.y * 2 + 1, it.y + r.w);  // second half, after the links of the first
            items[slot + 1] = make_int2(r.x * 2 + 0, it.y);    // first half on top
          }
          if (top + push <= kPathItems) top += push;
          __syncwarp();
        }
        if (__any_sync(0xffffffffu, bad)) {
          if (lane == 0) atomicOr(&ctr->overflow, kOvfStage);
        } else {
          // 4. cost in travel order
          for (int k0 = 0; k0 < hops; k0 += 32) {
            const double c = k0 + lane < hops ? link_cost[arena[off + k0 + lane]] : 0.0;
            const int m2 = min(32, hops - k0);
            for (int j = 0; j < m2; ++j) cost = __dadd_rn(cost, __shfl_sync(0xffffffffu, c, j));
          }
          status = DRIFT_ROUTE_OK;
        }
      }
    }
    if (lane == 0) {
      out.status[q] = status;
      out.n_links[q] = status == DRIFT_ROUTE_OK ? hops : 0;
      out.off[q] = status == DRIFT_ROUTE_OK ? off : 0;
      out.cost[q] = status == DRIFT_ROUTE_OK ? cost : 0.0;
    }
  }
}

}  // namespace drift

// drift_b200 -- the tick: advance, link events, ordered occupancy + BPR entry speed.
//
// Reference semantics (file:line in jbaudru/DRIFT):
//   lib/simulation_thread.py:299-303,326-340  agents are updated sequentially in list order
//   lib/agent.py:189-203                      d += speed*dt ; if d >= edge_length: next link, d = 0
//   lib/agent.py:106-160                      leave old link, enter new link, speed from BPR factor
//   lib/managers/traffic_manager.py:63-107    occupancy lists + factor(count/capacity)
// Parallel formulation (validated bit-exact against the sequential order):
//   every crossing/departure becomes a link event; the BPR count an entering agent `a` sees on
//   link l is   occ_start[l] + #{enter events on l with agent <= a} - #{leave events on l with agent < a}.
#pragma once
#include "common.cuh"

namespace drift {

// Block-level slot allocation: every thread asks for `n` consecutive slots of a global
// bump counter; one global atomic per block (single-address L2 atomics serialise at
// ~1 op/cycle chip-wide, so they are issued per block, not per event).
__device__ __forceinline__ int block_alloc(int n, int32_t* global_counter, int* s_cnt, int* s_base) {
  if (threadIdx.x == 0) *s_cnt = 0;
  __syncthreads();
  int my = 0;
  if (n) my = atomicAdd(s_cnt, n);
  __syncthreads();
  if (threadIdx.x == 0) *s_base = (*s_cnt) ? atomicAdd(global_counter, *s_cnt) : 0;
  __syncthreads();
  return *s_base + my;
}

__device__ __forceinline__ double bpr_factor_lookup(const LinkArrays& lk, int32_t link, int32_t count) {
  // host-tabulated max(0.1, 1/(1+alpha*(count/cap)**beta)) (traffic_manager.py:88-107)
  int32_t c = lk.cls[link];
  int64_t lo = lk.cls_off[c], hi = lk.cls_off[c + 1];
  int64_t idx = lo + (int64_t)count;
  if (idx >= hi) idx = hi - 1;  // saturated at the 10 % floor
  return lk.factor[idx];
}

constexpr int kAdvThreads = 256;
constexpr int kAdvPerThread = 4;

// K1+K2: advance every agent by one tick and emit the link events of the crossers.
// DEMAND: waiting agents also draw their Bernoulli departure (agent.py:176-183).
template <bool DEMAND>
__global__ void __launch_bounds__(kAdvThreads) k_advance(AgentArrays a, const int32_t* __restrict__ arena,
                                                          double delta, int32_t tick, Counters* ctr,
                                                          Event* __restrict__ events, int32_t ev_cap, Rings rg,
                                                          DemandArgs dm) {
  __shared__ int s_cnt, s_base;
  const int64_t tiles = (a.n + (int64_t)kAdvThreads * kAdvPerThread - 1) / ((int64_t)kAdvThreads * kAdvPerThread);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t i0 = (tile * kAdvThreads + threadIdx.x) * kAdvPerThread;
    int32_t ev_link[2 * kAdvPerThread];
    int32_t ev_kind_agent[2 * kAdvPerThread];  // local agent offset (0..3) | kind in bit 8
    int nev = 0;
    if (i0 < a.n) {
      const uchar4 st4 = *reinterpret_cast<const uchar4*>(a.state + i0);
      const uint8_t st[4] = {st4.x, st4.y, st4.z, st4.w};
      uint8_t nst[4] = {st[0], st[1], st[2], st[3]};
      bool state_changed = false;
      if (st4.x | st4.y | st4.z | st4.w) {
        const double2 d01 = *reinterpret_cast<const double2*>(a.dist + i0);
        const double2 d23 = *reinterpret_cast<const double2*>(a.dist + i0 + 2);
        const double2 s01 = *reinterpret_cast<const double2*>(a.speed + i0);
        const double2 s23 = *reinterpret_cast<const double2*>(a.speed + i0 + 2);
        const double2 e01 = *reinterpret_cast<const double2*>(a.elen + i0);
        const double2 e23 = *reinterpret_cast<const double2*>(a.elen + i0 + 2);
        double d[4] = {d01.x, d01.y, d23.x, d23.y};
        const double sp[4] = {s01.x, s01.y, s23.x, s23.y};
        const double el[4] = {e01.x, e01.y, e23.x, e23.y};
#pragma unroll
        for (int j = 0; j < kAdvPerThread; ++j) {
          if (st[j] == kMoving) {
            // two roundings, never contracted to an FMA (agent.py:197-198)
            const double nd = __dadd_rn(d[j], __dmul_rn(sp[j], delta));
            if (nd >= el[j]) {  // agent.py:200-203: at most one link per tick, overshoot dropped
              const int64_t i = i0 + j;
              const int32_t ri = a.route_idx[i] + 1;
              a.route_idx[i] = ri;
              d[j] = 0.0;
              ev_link[nev] = a.cur_link[i];
              ev_kind_agent[nev] = j;  // leave
              ++nev;
              if (ri < a.route_len[i]) {
                const int32_t nl = arena[a.route_off[i] + ri];
                a.cur_link[i] = nl;
                ev_link[nev] = nl;
                ev_kind_agent[nev] = j | 256;  // enter
                ++nev;
              } else {  // agent.py:107-116: arrived
                a.cur_link[i] = -1;
                nst[j] = kWaiting;
                state_changed = true;
                const unsigned long long slot = atomicAdd(&ctr->n_arrivals, 1ULL);
                if (slot - rg.arr_drained < rg.arr_cap) {
                  rg.arr_agent[slot % rg.arr_cap] = (int32_t)i;
                  rg.arr_tick[slot % rg.arr_cap] = tick;
                } else {
                  atomicOr(&ctr->overflow, kOvfArrivals);
                }
              }
            } else {
              d[j] = nd;
            }
          }
        }
        *reinterpret_cast<double2*>(a.dist + i0) = make_double2(d[0], d[1]);
        *reinterpret_cast<double2*>(a.dist + i0 + 2) = make_double2(d[2], d[3]);
      }
      if (DEMAND) {
#pragma unroll
        for (int j = 0; j < kAdvPerThread; ++j) {
          const int64_t i = i0 + j;
          if (st[j] != kWaiting || i >= a.n) continue;  // an agent arriving this tick draws from the next tick on
          const int cnt = a.chain_cnt[i];
          if (cnt == 0) continue;  // no destination available (chain exhausted)
          const double u = philox_uniform53(dm.seed, a.base + (uint32_t)i, (uint32_t)tick);
          if (!(u < dm.p_dep)) continue;  // agent.py:180
          const int32_t ns = a.next_status[i];
          if (ns == 0) {  // no prefetched route: routed inside this tick (rare)
            dm.dep_req[atomicAdd(&ctr->n_dep_req, 1)] = (int32_t)i;
            continue;
          }
          // Agent.start_new_journey (agent.py:46-104) with the prefetched route
          const int head = a.chain_head[i];
          const int32_t src = a.cur_node[i], dst = a.chain_dst[i * a.chain_cap + head];
          a.chain_head[i] = (uint8_t)((head + 1) % a.chain_cap);
          a.chain_cnt[i] = (uint8_t)(cnt - 1);
          a.trip_count[i] += 1;
          a.cur_node[i] = dst;
          const int64_t noff = a.next_off[i];
          const int32_t nlen = a.next_len[i];
          a.next_status[i] = a.next_status[a.npad + i];  // shift the second prefetched route up
          a.next_off[i] = a.next_off[a.npad + i];
          a.next_len[i] = a.next_len[a.npad + i];
          a.next_status[a.npad + i] = 0;
          const unsigned long long slot = atomicAdd(&ctr->n_departures, 1ULL);
          if (slot - dm.log_drained < dm.log_cap) {
            const unsigned long long k = slot % dm.log_cap;
            dm.log_agent[k] = (int32_t)i;
            dm.log_tick[k] = tick;
            dm.log_src[k] = src;
            dm.log_dst[k] = dst;
            dm.log_status[k] = ns - 1;
          } else {
            atomicOr(&ctr->overflow, kOvfDepLog);
          }
          if (ns == 1 + DRIFT_ROUTE_OK) {
            const int32_t fl = arena[noff];
            a.route_off[i] = noff;
            a.route_len[i] = nlen;
            a.route_idx[i] = 0;
            a.dist[i] = 0.0;
            a.cur_link[i] = fl;
            nst[j] = kMoving;
            state_changed = true;
            ev_link[nev] = fl;
            ev_kind_agent[nev] = j | 256;
            ++nev;
          }
        }
      }
      if (state_changed) *reinterpret_cast<uchar4*>(a.state + i0) = make_uchar4(nst[0], nst[1], nst[2], nst[3]);
    }
    const int slot0 = block_alloc(nev, &ctr->n_events, &s_cnt, &s_base);
    if (nev) {
      if (slot0 + nev <= ev_cap) {
        for (int k = 0; k < nev; ++k) {
          Event e;
          e.link = ev_link[k];
          e.agent = a.base + (uint32_t)(i0 + (ev_kind_agent[k] & 255));
          e.kind = (ev_kind_agent[k] & 256) ? 1 : -1;
          e.pos = 0;
          events[slot0 + k] = e;
        }
      } else {
        atomicOr(&ctr->overflow, kOvfEvents);
      }
    }
  }
}

// Agent.start_new_journey tail (agent.py:101-104) for the routed pending departures: the agent
// becomes 'moving' on the first link of its route and emits the enter event of that link.
__global__ void k_inject_departures(AgentArrays a, const int32_t* __restrict__ arena, Counters* ctr,
                                    const int32_t* __restrict__ pend_agent, const int64_t* __restrict__ pend_off,
                                    const int32_t* __restrict__ pend_len, const int32_t* __restrict__ n_pending_ptr,
                                    Event* __restrict__ events, int32_t ev_cap) {
  __shared__ int s_cnt, s_base;
  const int n_pending = *n_pending_ptr;
  for (int p0 = blockIdx.x * blockDim.x; p0 < n_pending; p0 += gridDim.x * blockDim.x) {
    const int p = p0 + threadIdx.x;
    int nev = 0;
    int32_t link = -1;
    int32_t i = -1;
    if (p < n_pending && pend_len[p] > 0) {
      i = pend_agent[p];
      link = arena[pend_off[p]];
      a.state[i] = kMoving;
      a.route_off[i] = pend_off[p];
      a.route_len[i] = pend_len[p];
      a.route_idx[i] = 0;
      a.dist[i] = 0.0;
      a.cur_link[i] = link;
      nev = 1;
    }
    const int slot = block_alloc(nev, &ctr->n_events, &s_cnt, &s_base);
    if (nev) {
      if (slot < ev_cap) {
        Event e;
        e.link = link;
        e.agent = a.base + (uint32_t)i;
        e.kind = 1;
        e.pos = 0;
        events[slot] = e;
      } else {
        atomicOr(&ctr->overflow, kOvfEvents);
      }
    }
  }
}

// K3 step 1: slot of every event inside its link (arrival order, only used as an address)
// and the list of touched links.
__global__ void k_count(Event* __restrict__ events, const int32_t* __restrict__ n_events_ptr,
                        int32_t* __restrict__ evcnt, int32_t* __restrict__ touched, Counters* ctr) {
  __shared__ int s_cnt, s_base;
  const int n = *n_events_ptr;
  for (int e0 = blockIdx.x * blockDim.x; e0 < n; e0 += gridDim.x * blockDim.x) {
    const int e = e0 + threadIdx.x;
    int first = 0;
    int32_t link = -1;
    if (e < n) {
      link = events[e].link;
      const int pos = atomicAdd(&evcnt[link], 1);
      events[e].pos = pos;
      first = (pos == 0);
    }
    const int slot = block_alloc(first, &ctr->n_touched, &s_cnt, &s_base);
    if (first) touched[slot] = link;
  }
}

// K3 step 2: one contiguous segment per touched link; snapshot of its start-of-tick occupancy.
__global__ void k_segment(const int32_t* __restrict__ touched, Counters* ctr, int32_t* __restrict__ evcnt,
                          int32_t* __restrict__ segid, Segment* __restrict__ segs,
                          const int32_t* __restrict__ occ) {
  __shared__ int s_cnt, s_base;
  const int n = ctr->n_touched;
  for (int t0 = blockIdx.x * blockDim.x; t0 < n; t0 += gridDim.x * blockDim.x) {
    const int t = t0 + threadIdx.x;
    int m = 0;
    int32_t link = -1;
    if (t < n) {
      link = touched[t];
      m = evcnt[link];
      evcnt[link] = 0;  // ready for the next tick
    }
    const int base = block_alloc(m, &ctr->seg_cursor, &s_cnt, &s_base);
    if (t < n) {
      Segment s;
      s.link = link;
      s.base = base;
      s.count = m;
      s.occ0 = occ[link];
      segs[t] = s;
      segid[link] = t;
    }
  }
}

__global__ void k_scatter(const Event* __restrict__ events, const int32_t* __restrict__ n_events_ptr,
                          const int32_t* __restrict__ segid, const Segment* __restrict__ segs,
                          SortedEvent* __restrict__ sorted) {
  const int n = *n_events_ptr;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const Event ev = events[e];
    const Segment s = segs[segid[ev.link]];
    SortedEvent o;
    o.agent = ev.agent;
    o.kind = ev.kind;
    sorted[s.base + ev.pos] = o;
  }
}

// K3 step 3: ordered running count -> BPR factor -> entry speed (agent.py:154-160), and the
// end-of-tick occupancy of every touched link.
__global__ void k_resolve(const Event* __restrict__ events, const int32_t* __restrict__ n_events_ptr,
                          const int32_t* __restrict__ segid, const Segment* __restrict__ segs,
                          const SortedEvent* __restrict__ sorted, LinkArrays lk, AgentArrays a, double kph_to_mps,
                          int32_t tick, Counters* ctr, Rings rg) {
  const int n = *n_events_ptr;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const Event ev = events[e];
    const Segment s = segs[segid[ev.link]];
    int before = 0;  // sum of deltas of events ordered before this one (Gauss-Seidel order)
    int net = 0;
    for (int j = 0; j < s.count; ++j) {
      const SortedEvent o = sorted[s.base + j];
      net += o.kind;
      if (o.agent < ev.agent) before += o.kind;
    }
    if (ev.pos == 0) lk.occ[ev.link] = s.occ0 + net;
    if (ev.kind > 0) {
      const uint32_t li = ev.agent - a.base;
      if (ev.agent >= a.base && (int64_t)li < a.n) {
        const int32_t count = s.occ0 + before + 1;  // includes the entering agent itself
        const double f = bpr_factor_lookup(lk, ev.link, count);
        const double sp = __dmul_rn(__dmul_rn(lk.v0[ev.link], f), kph_to_mps);
        a.speed[li] = sp;
        a.elen[li] = lk.len[ev.link];
        if (rg.ent_cap) {
          const unsigned long long slot = atomicAdd(&ctr->n_entries, 1ULL);
          if (slot - rg.ent_drained < rg.ent_cap) {
            const unsigned long long k = slot % rg.ent_cap;
            rg.ent_tick[k] = tick;
            rg.ent_agent[k] = (int32_t)li;
            rg.ent_link[k] = ev.link;
            rg.ent_speed[k] = sp;
          } else {
            atomicOr(&ctr->overflow, kOvfEntries);
          }
        }
      }
    }
  }
}

// K3': per-link volume recount = histogram of cur_link over moving agents
// (traffic_manager.py:122-135: total_agents_on_roads / per-link list lengths).
__global__ void k_recount(AgentArrays a, int32_t* __restrict__ hist) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    if (a.state[i] == kMoving) {
      const int32_t l = a.cur_link[i];
      if (l >= 0) atomicAdd(&hist[l], 1);
    }
  }
}

// K4 (extension): tt[l] = t0[l] / factor(occ[l]) -- the BPR travel-time form of traffic_manager.py:92-107.
__global__ void k_travel_times(LinkArrays lk) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < lk.L; l += gridDim.x * blockDim.x) {
    const int32_t c = lk.occ[l];
    const double f = (c == 0) ? 1.0 : bpr_factor_lookup(lk, l, c);
    lk.tt[l] = __ddiv_rn(lk.t0[l], f);
  }
}

// Top up every agent's chain ring from a block of host-generated candidate destinations
// (k per agent, row-major); candidates that do not fit are dropped.
__global__ void k_refill_chains(AgentArrays a, const int32_t* __restrict__ cand, int32_t k) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    int cnt = a.chain_cnt[i];
    const int head = a.chain_head[i];
    for (int j = 0; j < k && cnt < a.chain_cap; ++j, ++cnt)
      a.chain_dst[i * a.chain_cap + (head + cnt) % a.chain_cap] = cand[i * k + j];
    a.chain_cnt[i] = (uint8_t)cnt;
  }
}

// per-CSR-entry routing weights from the per-link travel times (extension mode)
__global__ void k_refresh_weights(int32_t E, const double* __restrict__ tt, const int32_t* __restrict__ fwd_link,
                                  const int32_t* __restrict__ bwd_link, double* __restrict__ fwd_wc,
                                  double* __restrict__ bwd_wc) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    fwd_wc[e] = tt[fwd_link[e]];
    bwd_wc[e] = tt[bwd_link[e]];
  }
}

__global__ void k_stats(AgentArrays a, LinkArrays lk, Counters* ctr) {
  long long mv = 0, on = 0, lw = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += stride) mv += a.state[i] == kMoving;
  for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < lk.L; l += stride) {
    const int32_t c = lk.occ[l];
    on += c;
    lw += c > 0;
  }
  for (int o = 16; o; o >>= 1) {
    mv += __shfl_down_sync(0xffffffffu, mv, o);
    on += __shfl_down_sync(0xffffffffu, on, o);
    lw += __shfl_down_sync(0xffffffffu, lw, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (mv) atomicAdd((unsigned long long*)&ctr->moving, (unsigned long long)mv);
    if (on) atomicAdd((unsigned long long*)&ctr->on_roads, (unsigned long long)on);
    if (lw) atomicAdd((unsigned long long*)&ctr->links_with_traffic, (unsigned long long)lw);
  }
}

}  // namespace drift

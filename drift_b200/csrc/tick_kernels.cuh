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
//
// A tick is three phases separated by grid-wide barriers:
//   A  advance every agent, emit + link the events of crossers and of prefetched departures
//   A2 (only when needed) inject host-committed departures, route + inject departures that had
//      no prefetched route
//   C  resolve: every event walks the event list of its link
// run either as separate kernels (multi-rank exchange between A2 and C, profiling) or inside one
// persistent cooperative kernel that loops over many ticks (k_tick_loop).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "route_kernels.cuh"

namespace drift {

namespace cg = cooperative_groups;

constexpr int kAdvThreads = 256;
constexpr int kAdvPerThread = 4;

struct TickArgs {
  AgentArrays a;
  LinkArrays lk;
  GraphArrays g;
  Rings rg;
  DemandArgs dm;
  RouteScratch fsc;         // scratch of the in-tick fallback router
  const double* w_fwd;      // routing weights of the fallback router
  const double* w_bwd;
  const double* link_cost;
  int32_t* arena;
  unsigned long long arena_cap;
  Counters* ctr;
  Event* events;
  int32_t ev_cap;      // capacity of one sub-buffer; sub-buffer b starts at events + b * ev_cap
  const int32_t* pend_agent;
  const int64_t* pend_off;
  const int32_t* pend_len;
  int32_t pend_cap;         // entries of the pend_* arrays; ctr->n_pending may exceed it (flagged kOvfPending)
  double delta;
  double kph_to_mps;
  // A tile = 1024 agents handled four per thread + `tile_singles` agents (a multiple of 32, <= 256) handled one per
  // thread by the first threads of the block.  0 unless that is what lets every tile of the persistent kernel run in
  // ONE wave: a tick is latency-bound between grid barriers, and a block with a second tile makes the whole grid wait
  // a second tile time (1 M agents = 977 plain tiles on 888 resident blocks; 869 tiles of 1024 + 128).
  int32_t tile_singles;
  int32_t demand;           // device-driven departures on/off
  int32_t debug;            // timing probes only (tools/tick_probe.py): 1 = no departure draws, 2 = crossers stay, 4 = events not emitted
};

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

// Write event `slot` and push it on the event list of its link.  The first event of a link in
// this tick snapshots the link's start-of-tick occupancy (occ is only written in phase C).
// `tag` identifies the list generation: the tick number for local lists, tick | 2^31 for the
// lists rebuilt over the events gathered from all ranks.
__device__ __forceinline__ void emit_event(const LinkArrays& lk, Event* __restrict__ events, int slot, int32_t link,
                                           uint32_t agent, int32_t kind, uint32_t tag) {
  const unsigned long long mine = ((unsigned long long)tag << 32) | (uint32_t)slot;
  // issued alongside the exchange: only the first event needs it.  Never negative in a consistent state; the clamp
  // keeps the aux encoding (and so the list walks of phase C) well-formed whatever the state
  const int32_t occ0 = max(lk.occ[link], 0);
  const unsigned long long prev = atomicExch(&lk.head[link], mine);
  Event e;
  e.link = link;
  e.agent = agent;
  e.kind = kind;
  e.aux = ((uint32_t)(prev >> 32) == tag) ? (int32_t)(prev & 0xffffffffu) : -occ0 - 1;
  events[slot] = e;
}

// ---- phase A ------------------------------------------------------------------------------------
// K1+K2: advance every agent by one tick, emit the link events of the crossers; with device-driven
// demand waiting agents also draw their Bernoulli departure (agent.py:176-183) and leave on their
// prefetched route.
//
// The hot loop touches state (1 B) for every agent and dist/speed/elen (24 B read, 8 B written) for
// moving ones, 4 agents per thread with 16-byte loads.  Crossings and departures are rare: they run
// in out-of-line handlers that stage their events in shared memory; at the end of a tile the block
// reserves the slots with ONE global atomic (single-address L2 atomics serialise at ~1 op/cycle
// chip-wide) and all threads write + link the staged events.
constexpr int kAdvWarps = kAdvThreads / 32;
constexpr int kAdvTileQuads = kAdvThreads * kAdvPerThread;  // agents of a tile that are processed four per thread
constexpr int kAdvMaxSingles = kAdvThreads;                 // + at most one more agent per thread (TickArgs::tile_singles)
constexpr int kStageCap = 2 * 32 * (kAdvPerThread + 1);     // every agent of a warp's slice emits <= 2 events

struct AdvSmem {  // per-warp staging: warps do not wait for each other while they process agents
  int n[kAdvWarps];
  int base;
  uint32_t link_kind[kAdvWarps][kStageCap];  // link | (enter ? 2^31 : 0)
  uint32_t agent[kAdvWarps][kStageCap];
};

// ---- multi-GPU over NVLink peer memory (k_tick_loop_p2p) -------------------------------------------
// Links are owned round-robin (owner = link % world).  A rank writes every event of its agents straight
// into the owner's inbox (peer store), the owner orders the events of its links and writes the entry
// speed straight into the home rank's agent arrays: the exchange is part of the tick kernel, the work
// per rank does not grow with the number of ranks, occupancy is held by the owner only.
constexpr int kMaxRanks = 16;
// Entry speed not yet delivered by the link's owner: the home rank writes this when it emits the enter
// event and the next advance of that agent waits for a real (positive) speed -- per-agent hand-over
// instead of a second all-rank barrier per tick.
constexpr double kSpeedPending = -1.0;
// Same, for an agent that CANNOT cross its new link in its first tick on it even at free-flow speed (the
// common case: links are longer than one tick of travel): its first advance produces no event, so the home rank
// does not wait for the owner at all -- it skips that advance (marking dist = -0.0) and applies it together with
// the next one, when the owner's speed is guaranteed to have landed (the owner resolved tick k before it published
// its events of tick k+1, which the home rank acquired before its own tick k+2).  Bit-exact: the two advances are
// the same two fp64 operations, only later.
constexpr double kSpeedDeferred = -2.0;
constexpr long long kNegZeroBits = (long long)0x8000000000000000ULL;

__device__ __forceinline__ bool owes_advance(double d) { return __double_as_longlong(d) == kNegZeroBits; }

struct P2PArgs {
  int32_t rank, world;
  int32_t cap;                            // events one rank may send to one owner per tick
  uint32_t shard_base, shard_rem;         // agents per rank: n_total / world (+1 for the first n_total % world ranks)
  Event* inbox[kMaxRanks];                // inbox of every rank (peer-mapped): [2 tick parities][world][cap], region r written by rank r
  unsigned long long* flags[kMaxRanks];   // of every rank: [3][kMaxRanks]; rows 0-1 (tick parity): tick << 32 | events
                                          // sent by rank r, row 2: last tick << 32 once rank r has finished a launch
  double* speed[kMaxRanks];               // agent arrays of every rank
  double* elen[kMaxRanks];
  int32_t* out_cnt;                       // local [kMaxRanks]: events written to each owner in this tick
};

// An event record travels to its owner WITHOUT a fence: its 4th word carries a check value over (tick, link, agent,
// kind), and the owner polls each expected record until the check matches (a system-scope fence after the peer
// stores costs ~6 us per block and tick on NVLink -- more than a quarter of the local tick).  A stale record of the
// same slot (two ticks old, its 4th word since overwritten by the list link) or a torn one fails the check.
__device__ __forceinline__ int32_t event_check(uint32_t tick, int32_t link, uint32_t agent, int32_t kind) {
  return (int32_t)(tick * 2654435761u ^ (uint32_t)link * 0x85EBCA6Bu ^ agent * 0xC2B2AE35u ^ (uint32_t)kind ^ 0x5bd1e995u);
}

constexpr int kP2PSorted = 640;  // keeps the kernel at 6 blocks per SM (shared memory)

struct P2PSmem {
  int ocnt[kMaxRanks], obase[kMaxRanks], ooff[kMaxRanks + 1];
  uint16_t idx[kAdvWarps][kStageCap];
  uint2 sorted[kP2PSorted];  // the tile's events grouped by owner: (link | enter bit, agent); a tile with more events
                             // than this (rare: every agent of the tile crossing at once) is written record by record
};

__device__ __noinline__ double p2p_wait_speed(const double* p) {
  unsigned long long t0 = 0;
  for (;;) {
    unsigned long long bits;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(bits) : "l"(p) : "memory");
    const double v = __longlong_as_double((long long)bits);
    if (v >= 0.0) return v;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (!t0) t0 = now;
    if (now - t0 > 60000000000ULL) {
      printf("drift_b200: an entry speed was not delivered within 60 s\n");
      __trap();
    }
    __nanosleep(32);
  }
}

// kSpeedDeferred when even the free-flow speed (factor 1.0, the largest the BPR table holds) cannot carry the agent
// over the whole link in one tick, else kSpeedPending.  Rounding is monotone, so the bound holds for every factor.
__device__ __forceinline__ double first_tick_speed_marker(const LinkRec& r, double delta, double kph_to_mps) {
  const double smax = __dmul_rn(__dmul_rn(r.v0, 1.0), kph_to_mps);
  return __dadd_rn(0.0, __dmul_rn(smax, delta)) < r.len ? kSpeedDeferred : kSpeedPending;
}

__device__ __forceinline__ void stage_event(AdvSmem& sm, int32_t link, uint32_t agent, bool enter) {
  const int w = threadIdx.x >> 5;
  const int k = atomicAdd(&sm.n[w], 1);
  sm.link_kind[w][k] = (uint32_t)link | (enter ? 0x80000000u : 0u);
  sm.agent[w][k] = agent;
}

// Agent crosses to the next link of its route or arrives (agent.py:200-203, 107-116).  Returns true on arrival.
__device__ __noinline__ bool cross_agent(const TickArgs& p, int64_t i, int32_t tick, AdvSmem& sm, bool own_next) {
  const AgentArrays& a = p.a;
  const int32_t old_link = a.cur_link[i], nl = a.next_link[i];
  const int32_t ri = a.route_idx[i] + 1;
  a.route_idx[i] = ri;
  stage_event(sm, old_link, a.base + (uint32_t)i, false);
  if (nl >= 0) {  // next_link is refreshed in phase C when the enter event is resolved ...
    a.cur_link[i] = nl;
    // ... unless the link is resolved on another GPU (k_tick_loop_p2p): then the home rank does it here
    if (own_next) {
      a.next_link[i] = ri + 1 < a.route_len[i] ? p.arena[a.route_off[i] + ri + 1] : -1;
      const LinkRec r = p.lk.rec[nl];  // link tables are replicated: only the speed has to come from the owner
      a.elen[i] = r.len;
      a.speed[i] = first_tick_speed_marker(r, p.delta, p.kph_to_mps);
    }
    stage_event(sm, nl, a.base + (uint32_t)i, true);
    return false;
  }
  a.cur_link[i] = -1;
  const unsigned long long slot = atomicAdd(&p.ctr->n_arrivals, 1ULL);
  if (slot - p.rg.arr_drained < p.rg.arr_cap) {
    p.rg.arr_agent[slot % p.rg.arr_cap] = (int32_t)i;
    p.rg.arr_tick[slot % p.rg.arr_cap] = tick;
  } else {
    atomicOr(&p.ctr->overflow, kOvfArrivals);
  }
  return true;
}

// Agent.start_new_journey (agent.py:46-104) with the prefetched route.  Returns true if the agent is now moving.
// Rare (about one agent in a thousand per tick) but on the critical path of its block: all loads are
// issued up front in two dependent levels, then the stores -- written naively (load, store, load ...)
// the ~20 accesses serialise and the block reaches the grid barrier ~10 us late.
__device__ __noinline__ bool depart_agent(const TickArgs& p, int64_t i, int cnt, int32_t tick, AdvSmem& sm, bool own_next) {
  const AgentArrays& a = p.a;
  Counters* ctr = p.ctr;
  // level 1
  const int32_t ns = a.next_status[i];
  const int head = a.chain_head[i];
  const int32_t src = a.cur_node[i];
  const int64_t noff = a.next_off[i];
  const int32_t nlen = a.next_len[i];
  const int32_t ns2 = a.next_status[a.npad + i];
  const int64_t noff2 = a.next_off[a.npad + i];
  const int32_t nlen2 = a.next_len[a.npad + i];
  const int32_t trips = a.trip_count[i];
  if (ns == 0) {  // no prefetched route: routed in phase A2 of this tick (rare)
    p.dm.dep_req[atomicAdd(&ctr->n_dep_req[tick & 1], 1)] = (int32_t)i;
    return false;
  }
  // level 2
  const bool go = ns == 1 + DRIFT_ROUTE_OK;
  const int32_t dst = a.chain_dst[i * a.chain_cap + head];
  const int32_t fl = go ? p.arena[noff] : -1;
  const int32_t nl = (go && own_next && nlen > 1) ? p.arena[noff + 1] : -1;
  const unsigned long long slot = atomicAdd(&ctr->n_departures, 1ULL);
  // stores
  a.chain_head[i] = (uint8_t)((head + 1) % a.chain_cap);
  a.chain_cnt[i] = (uint8_t)(cnt - 1);
  a.trip_count[i] = trips + 1;
  a.cur_node[i] = dst;
  a.next_status[i] = ns2;  // shift the second prefetched route up
  a.next_off[i] = noff2;
  a.next_len[i] = nlen2;
  a.next_status[a.npad + i] = 0;
  if (slot - p.dm.log_drained < p.dm.log_cap) {
    const unsigned long long k = slot % p.dm.log_cap;
    p.dm.log_agent[k] = (int32_t)i;
    p.dm.log_tick[k] = tick;
    p.dm.log_src[k] = src;
    p.dm.log_dst[k] = dst;
    p.dm.log_status[k] = ns - 1;
    p.dm.log_off[k] = noff;
    p.dm.log_len[k] = go ? nlen : 0;
    p.dm.log_epoch[k] = p.dm.epoch;
  } else {
    atomicOr(&ctr->overflow, kOvfDepLog);
  }
  if (!go) return false;
  a.route_off[i] = noff;
  a.route_len[i] = nlen;
  a.route_idx[i] = 0;
  a.dist[i] = 0.0;
  a.cur_link[i] = fl;
  if (own_next) {
    const LinkRec r = p.lk.rec[fl];
    a.next_link[i] = nl;
    a.elen[i] = r.len;
    a.speed[i] = first_tick_speed_marker(r, p.delta, p.kph_to_mps);
  }
  stage_event(sm, fl, a.base + (uint32_t)i, true);
  return true;
}

// Peer exchange: an agent that entered a link in the previous tick is advanced after everything else of
// its tile, when the owner of that link has had the time of a whole tile to deliver the entry speed.
__device__ __noinline__ bool advance_late(const TickArgs& p, int64_t i, int32_t tick, AdvSmem& sm) {
  const AgentArrays& a = p.a;
  double d = a.dist[i];
  const double el = __ldcg(a.elen + i);
  const double sp = p2p_wait_speed(a.speed + i);
  if (owes_advance(d)) d = __dadd_rn(0.0, __dmul_rn(sp, p.delta));  // the advance skipped last tick (cannot cross)
  const double nd = __dadd_rn(d, __dmul_rn(sp, p.delta));  // agent.py:197-198
  if (nd >= el) {
    a.dist[i] = 0.0;
    return cross_agent(p, i, tick, sm, true);
  }
  a.dist[i] = nd;
  return false;
}

template <bool DEMAND, bool P2P = false>
__device__ __forceinline__ void phase_advance(const TickArgs& p, int32_t tick, double p_dep, AdvSmem& sm,
                                              const P2PArgs* x = nullptr, P2PSmem* xs = nullptr) {
  const AgentArrays& a = p.a;
  const int par = tick & 1;
  const unsigned long long dep_thresh = DEMAND ? departure_threshold(p_dep) : 0ULL;
  const int64_t tile_agents = (int64_t)kAdvTileQuads + p.tile_singles;
  const int64_t tiles = (a.n + tile_agents - 1) / tile_agents;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    if (lane == 0) sm.n[w] = 0;
    if (P2P && threadIdx.x < kMaxRanks) xs->ocnt[threadIdx.x] = 0;
    __syncwarp();
    const int64_t i0 = tile * tile_agents + (int64_t)threadIdx.x * kAdvPerThread;
    // the thread's single agent of this tile (tile_singles > 0 only), handled after the quad
    const int64_t is = (int)threadIdx.x < p.tile_singles ? tile * tile_agents + kAdvTileQuads + threadIdx.x : a.n;
    if (i0 < a.n) {
      const uint32_t st4 = *reinterpret_cast<const uint32_t*>(a.state + i0);
      uint32_t cn4 = 0;
      if (DEMAND) cn4 = *reinterpret_cast<const uint32_t*>(a.chain_cnt + i0);  // independent of the state load
      uint32_t nst4 = st4;
      uint32_t pend = 0;  // P2P: agents waiting for their entry speed
      if (st4) {  // at least one of the four agents is moving
        const double2 d01 = *reinterpret_cast<const double2*>(a.dist + i0);
        const double2 d23 = *reinterpret_cast<const double2*>(a.dist + i0 + 2);
        // speed / elen may have been written by another GPU: read them past the (non-coherent) L1
        const double2 s01 = P2P ? __ldcg(reinterpret_cast<const double2*>(a.speed + i0))
                                : *reinterpret_cast<const double2*>(a.speed + i0);
        const double2 s23 = P2P ? __ldcg(reinterpret_cast<const double2*>(a.speed + i0 + 2))
                                : *reinterpret_cast<const double2*>(a.speed + i0 + 2);
        const double2 e01 = P2P ? __ldcg(reinterpret_cast<const double2*>(a.elen + i0))
                                : *reinterpret_cast<const double2*>(a.elen + i0);
        const double2 e23 = P2P ? __ldcg(reinterpret_cast<const double2*>(a.elen + i0 + 2))
                                : *reinterpret_cast<const double2*>(a.elen + i0 + 2);
        double d[4] = {d01.x, d01.y, d23.x, d23.y};
        double sp[4] = {s01.x, s01.y, s23.x, s23.y};
        const double el[4] = {e01.x, e01.y, e23.x, e23.y};
        uint32_t skip = 0;  // P2P: first advance on a new link deferred to the next tick (kSpeedDeferred)
        if (P2P) {  // entered a link last tick and the owner's speed has not landed yet
#pragma unroll
          for (int j = 0; j < kAdvPerThread; ++j)
            if (((st4 >> (8 * j)) & 0xffu) && sp[j] < 0.0) {
              if (sp[j] == kSpeedDeferred && !owes_advance(d[j]))
                skip |= 1u << j;  // cannot cross this tick: no event, nothing to wait for
              else
                pend |= 1u << j;  // may cross at once (or already skipped once): advanced last (below)
            }
        }
        uint32_t cross = 0;
#pragma unroll
        for (int j = 0; j < kAdvPerThread; ++j) {
          if (((st4 >> (8 * j)) & 0xffu) && !(((pend | skip) >> j) & 1u)) {
            // two roundings, never contracted to an FMA (agent.py:197-198)
            double dj = d[j];
            if (P2P && owes_advance(dj)) dj = __dadd_rn(0.0, __dmul_rn(sp[j], p.delta));  // the skipped advance first
            const double nd = __dadd_rn(dj, __dmul_rn(sp[j], p.delta));
            const bool c = nd >= el[j];  // agent.py:200-203: at most one link per tick, overshoot dropped
            d[j] = c ? 0.0 : nd;
            cross |= (c ? 1u : 0u) << j;
          } else if (P2P && ((skip >> j) & 1u)) {
            d[j] = __longlong_as_double(kNegZeroBits);  // owes this tick's advance
          }
        }
        *reinterpret_cast<double2*>(a.dist + i0) = make_double2(d[0], d[1]);
        *reinterpret_cast<double2*>(a.dist + i0 + 2) = make_double2(d[2], d[3]);
        if (p.debug & 2) cross = 0;
        while (cross) {
          const int j = __ffs(cross) - 1;
          cross &= cross - 1;
          if (cross_agent(p, i0 + j, tick, sm, P2P)) nst4 &= ~(0xffu << (8 * j));  // arrived: waiting again
        }
      }
      if (DEMAND) {
        // agents that draw this tick: waiting at tick start (an agent arriving this tick draws from
        // the next tick on) with a trip in their ring
        uint32_t want = 0;
#pragma unroll
        for (int j = 0; j < kAdvPerThread; ++j) {
          const int cnt = (cn4 >> (8 * j)) & 0xffu;
          if (((st4 >> (8 * j)) & 0xffu) == kWaiting && cnt != 0 && i0 + j < a.n) want |= 1u << j;
        }
        if (p.debug & 1) want = 0;
        uint32_t dep = 0;
        if (want) {
          const uint32_t gid0 = a.base + (uint32_t)i0;
          if ((gid0 & 1u) == 0) {  // one Philox block serves the two agents of a pair
#pragma unroll
            for (int h = 0; h < kAdvPerThread / 2; ++h) {
              if (!(want & (3u << (2 * h)))) continue;
              uint32_t o[4];
              philox4x32_10(p.dm.seed, (gid0 >> 1) + (uint32_t)h, (uint32_t)tick, o);
              if (bits53(o[0], o[1]) < dep_thresh) dep |= 1u << (2 * h);
              if (bits53(o[2], o[3]) < dep_thresh) dep |= 2u << (2 * h);
            }
            dep &= want;
          } else {
#pragma unroll
            for (int j = 0; j < kAdvPerThread; ++j)
              if (((want >> j) & 1u) && philox_bits53(p.dm.seed, gid0 + (uint32_t)j, (uint32_t)tick) < dep_thresh)
                dep |= 1u << j;
          }
        }
        while (dep) {  // agent.py:180
          const int j = __ffs(dep) - 1;
          dep &= dep - 1;
          if (depart_agent(p, i0 + j, (cn4 >> (8 * j)) & 0xffu, tick, sm, P2P)) nst4 |= (uint32_t)kMoving << (8 * j);
        }
      }
      if (P2P) {  // by now the owner has usually delivered: the wait overlaps with the rest of the tile
        while (pend) {
          const int j = __ffs(pend) - 1;
          pend &= pend - 1;
          if (advance_late(p, i0 + j, tick, sm)) nst4 &= ~(0xffu << (8 * j));
        }
      }
      if (nst4 != st4) *reinterpret_cast<uint32_t*>(a.state + i0) = nst4;
    }
    if (is < a.n) {  // the same steps for the thread's single agent
      const uint32_t st1 = a.state[is];
      const uint32_t cn1 = DEMAND ? a.chain_cnt[is] : 0u;
      uint32_t nst1 = st1;
      bool pend1 = false;
      if (st1) {
        const double d1 = a.dist[is];
        const double s1 = P2P ? __ldcg(a.speed + is) : a.speed[is];
        const double e1 = P2P ? __ldcg(a.elen + is) : a.elen[is];
        bool skip1 = false;
        if (P2P && s1 < 0.0) {
          if (s1 == kSpeedDeferred && !owes_advance(d1)) skip1 = true; else pend1 = true;
        }
        if (skip1) {
          a.dist[is] = __longlong_as_double(kNegZeroBits);
        } else if (!pend1) {
          double dj = d1;
          if (P2P && owes_advance(dj)) dj = __dadd_rn(0.0, __dmul_rn(s1, p.delta));
          const double nd = __dadd_rn(dj, __dmul_rn(s1, p.delta));  // agent.py:197-198
          const bool c = nd >= e1;
          a.dist[is] = c ? 0.0 : nd;
          if (c && !(p.debug & 2) && cross_agent(p, is, tick, sm, P2P)) nst1 = kWaiting;
        }
      }
      if (DEMAND && st1 == kWaiting && cn1 != 0 && !(p.debug & 1) &&
          philox_bits53(p.dm.seed, a.base + (uint32_t)is, (uint32_t)tick) < dep_thresh) {  // agent.py:180
        if (depart_agent(p, is, (int)cn1, tick, sm, P2P)) nst1 = kMoving;
      }
      if (P2P && pend1 && advance_late(p, is, tick, sm)) nst1 = kWaiting;
      if (nst1 != st1) a.state[is] = (uint8_t)nst1;
    }
    __syncthreads();  // all warps of the block have staged their events
    int n_blk = 0, my_off = 0;
#pragma unroll
    for (int k = 0; k < kAdvWarps; ++k) {
      if (k == w) my_off = n_blk;
      n_blk += sm.n[k];
    }
    if (P2P) {
      if (n_blk) {  // bucket the staged events by owner: one local atomic per block, tile and owner
        const int n = sm.n[w];
        for (int k = lane; k < n; k += 32)
          xs->idx[w][k] = (uint16_t)atomicAdd(&xs->ocnt[(sm.link_kind[w][k] & 0x7fffffffu) % (uint32_t)x->world], 1);
        __syncthreads();
        if (threadIdx.x < x->world && xs->ocnt[threadIdx.x])
          xs->obase[threadIdx.x] = atomicAdd(&x->out_cnt[threadIdx.x], xs->ocnt[threadIdx.x]);
        if (threadIdx.x == 0) {
          int acc = 0;
          for (int o = 0; o < x->world; ++o) {
            xs->ooff[o] = acc;
            acc += xs->ocnt[o];
          }
          xs->ooff[x->world] = acc;
        }
        __syncthreads();
        // group the events by owner in shared memory, then the whole block writes each owner's run with consecutive
        // threads: the peer stores of a tile are contiguous 16-byte records (full 128-byte NVLink writes) instead of
        // one scattered record per lane
        const bool grouped = n_blk <= kP2PSorted;
        if (grouped) {
          for (int k = lane; k < n; k += 32) {
            const uint32_t lkd = sm.link_kind[w][k];
            const int o = (int)((lkd & 0x7fffffffu) % (uint32_t)x->world);
            xs->sorted[xs->ooff[o] + xs->idx[w][k]] = make_uint2(lkd, sm.agent[w][k]);
          }
          __syncthreads();
        }
        // grouped: thread t writes the t-th record of the grouped tile; else every lane writes its own staged records
        for (int t = grouped ? (int)threadIdx.x : lane; t < (grouped ? n_blk : n); t += grouped ? (int)blockDim.x : 32) {
          int o = 0;
          uint2 rec;
          int pos;
          if (grouped) {
            while (xs->ooff[o + 1] <= t) ++o;
            rec = xs->sorted[t];
            pos = xs->obase[o] + (t - xs->ooff[o]);
          } else {
            rec = make_uint2(sm.link_kind[w][t], sm.agent[w][t]);
            o = (int)((rec.x & 0x7fffffffu) % (uint32_t)x->world);
            pos = xs->obase[o] + xs->idx[w][t];
          }
          const int32_t link = (int32_t)(rec.x & 0x7fffffffu);
          if (pos < x->cap) {
            Event e;
            e.link = link;
            e.agent = rec.y;
            e.kind = (rec.x >> 31) ? 1 : -1;
            e.aux = event_check((uint32_t)tick, link, rec.y, e.kind);
            *reinterpret_cast<int4*>(x->inbox[o] + ((int64_t)par * x->world + x->rank) * x->cap + pos) =
              *reinterpret_cast<const int4*>(&e);
          } else {
            // The owner never sees this event.  For an entry nobody would deliver the speed and the agent's next
            // advance would spin on kSpeedPending: give it the free-flow speed so that the launch ends and the host
            // can raise DRIFT_EOVERFLOW (the run is invalid from here on; raise the exchange capacity).
            atomicOr(&p.ctr->overflow, kOvfExchange);
            if (rec.x >> 31) a.speed[rec.y - a.base] = __dmul_rn(p.lk.v0[link], p.kph_to_mps);
          }
        }
        // no fence here: ONE thread per block issues the system-scope fence after the block's last tile (see
        // k_tick_loop_p2p) -- a membar.sys per warp and tile costs more than the whole local advance
      }
    } else if (n_blk && !(p.debug & 4)) {  // one global atomic per block and tile, spread over kEvSub counters
      const int sub = (int)(tile % kEvSub);
      if (threadIdx.x == 0) sm.base = atomicAdd(&p.ctr->n_events[par][sub], n_blk);
      __syncthreads();
      const int base = sm.base + my_off;
      const int n = sm.n[w];
      if (sm.base + n_blk <= p.ev_cap) {
        for (int k = lane; k < n; k += 32) {
          const uint32_t lkd = sm.link_kind[w][k];
          emit_event(p.lk, p.events, sub * p.ev_cap + base + k, (int32_t)(lkd & 0x7fffffffu), sm.agent[w][k],
                     (lkd >> 31) ? 1 : -1, (uint32_t)tick);
        }
      } else if (threadIdx.x == 0) {
        atomicOr(&p.ctr->overflow, kOvfEvents);
      }
    }
    __syncthreads();  // staging and sm.base are reused by the next tile
  }
}

// ---- phase A2 -----------------------------------------------------------------------------------
__device__ __forceinline__ void start_on_route(const TickArgs& p, int32_t i, int64_t off, int32_t len, int32_t tick) {
  const AgentArrays& a = p.a;
  if (a.state[i] == kMoving) return;  // committed while en route (caller error): the trip in progress wins
  const int32_t link = p.arena[off];
  a.state[i] = kMoving;
  a.route_off[i] = off;
  a.route_len[i] = len;
  a.route_idx[i] = 0;
  a.dist[i] = 0.0;
  a.cur_link[i] = link;
  const int sub = i % kEvSub;
  const int slot = atomicAdd(&p.ctr->n_events[tick & 1][sub], 1);
  if (slot < p.ev_cap)
    emit_event(p.lk, p.events, sub * p.ev_cap + slot, link, a.base + (uint32_t)i, 1, (uint32_t)tick);
  else
    atomicOr(&p.ctr->overflow, kOvfEvents);
}

// Agent.start_new_journey tail (agent.py:101-104) for host-committed departures, and the whole of
// it (selector output = next ring entry, NetworkX-exact route, one thread per agent) for device
// departures that had no prefetched route.
__device__ __noinline__ void phase_inject(const TickArgs& p, int32_t tick, int n_pending, int n_urgent) {
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = gtid; q < n_pending; q += gsz)
    if (p.pend_len[q] > 0) start_on_route(p, p.pend_agent[q], p.pend_off[q], p.pend_len[q], tick);
  if (n_urgent > 0 && gtid == 0) p.ctr->n_fallback += (unsigned long long)n_urgent;
  if (n_urgent > 0 && gtid < p.fsc.workers) {
    const AgentArrays& a = p.a;
    uint32_t epoch = p.fsc.epoch[gtid];
    for (int r = (int)gtid; r < n_urgent; r += p.fsc.workers) {
      const int32_t i = p.dm.dep_req[r];
      const int head = a.chain_head[i];
      const int32_t s = a.cur_node[i];
      const int32_t t = a.chain_dst[(int64_t)i * a.chain_cap + head];
      a.chain_head[i] = (uint8_t)((head + 1) % a.chain_cap);
      a.chain_cnt[i] -= 1;
      a.trip_count[i] += 1;
      a.cur_node[i] = t;  // agent.py:51: target is overwritten even if routing fails
      const RouteResult rr = route_nx_one(p.g, p.w_fwd, p.w_bwd, s, t, p.fsc, (int)gtid, epoch, p.arena, p.arena_cap,
                                          p.ctr, p.link_cost);
      const unsigned long long slot = atomicAdd(&p.ctr->n_departures, 1ULL);
      if (slot - p.dm.log_drained < p.dm.log_cap) {
        const unsigned long long k = slot % p.dm.log_cap;
        p.dm.log_agent[k] = i;
        p.dm.log_tick[k] = tick;
        p.dm.log_src[k] = s;
        p.dm.log_dst[k] = t;
        p.dm.log_status[k] = rr.status;
        p.dm.log_off[k] = rr.off;
        p.dm.log_len[k] = rr.n_links;
        p.dm.log_epoch[k] = p.dm.epoch;
      } else {
        atomicOr(&p.ctr->overflow, kOvfDepLog);
      }
      if (rr.status == DRIFT_ROUTE_OK) start_on_route(p, i, rr.off, rr.n_links, tick);
    }
    p.fsc.epoch[gtid] = epoch;
  }
}

// ---- phase C ------------------------------------------------------------------------------------
// K3: ordered running count -> BPR factor -> entry speed (agent.py:154-160), and the end-of-tick
// occupancy of every touched link.  Each event walks the list of its link.
__device__ __forceinline__ void resolve_one(const LinkArrays& lk, const AgentArrays& a, Counters* ctr,
                                            const Event* __restrict__ events, int e, int32_t tick, double kph_to_mps,
                                            const Rings& rg, const int32_t* __restrict__ arena);

// contiguous list [e0, e0 + n_events)
__device__ __forceinline__ void phase_resolve(const LinkArrays& lk, const AgentArrays& a, Counters* ctr,
                                              const Event* __restrict__ events, int e0, int n_events, int32_t tick,
                                              double kph_to_mps, const Rings& rg, const int32_t* __restrict__ arena) {
  for (int e = e0 + blockIdx.x * blockDim.x + threadIdx.x; e < e0 + n_events; e += gridDim.x * blockDim.x)
    resolve_one(lk, a, ctr, events, e, tick, kph_to_mps, rg, arena);
}

// the kEvSub local sub-buffers of this tick, flattened into one index space; s_pref: kEvSub + 1 ints of
// shared memory (exclusive prefix of the sub-buffer counts)
__device__ __forceinline__ void phase_resolve_local(const TickArgs& p, int32_t tick, int* s_pref) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < kEvSub; ++k) {
      s_pref[k] = acc;
      acc += min(p.ctr->n_events[tick & 1][k], p.ev_cap);
    }
    s_pref[kEvSub] = acc;
  }
  __syncthreads();
  const int total = s_pref[kEvSub];
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    int sub = 0;
#pragma unroll
    for (int step = kEvSub / 2; step; step >>= 1)
      if (s_pref[sub + step] <= f) sub += step;
    resolve_one(p.lk, p.a, p.ctr, p.events, sub * p.ev_cap + (f - s_pref[sub]), tick, p.kph_to_mps, p.rg, p.arena);
  }
}

__device__ __forceinline__ void resolve_one(const LinkArrays& lk, const AgentArrays& a, Counters* ctr,
                                            const Event* __restrict__ events, int e, int32_t tick, double kph_to_mps,
                                            const Rings& rg, const int32_t* __restrict__ arena) {
  const Event ev = events[e];
  // Everything that depends only on the event itself is loaded before the list walk, so that the walk
  // (a chain of dependent event loads) overlaps with the link / route lookups instead of preceding them.
  const int first = (int)(lk.head[ev.link] & 0xffffffffu);
  const uint32_t li = ev.agent - a.base;
  const bool mine = ev.kind > 0 && ev.agent >= a.base && (int64_t)li < a.n;  // entry of a local agent
  double v0 = 0.0, len = 0.0;
  int64_t f_lo = 0, f_hi = 1;
  int32_t next = -1;
  if (mine) {
    const LinkRec r = lk.rec[ev.link];  // one sector: v0, len, factor range
    v0 = r.v0;
    len = r.len;
    f_lo = r.f_lo;
    f_hi = r.f_hi;
    const int32_t ri = a.route_idx[li] + 1;
    const int32_t rlen = a.route_len[li];
    const int64_t roff = a.route_off[li];
    if (ri < rlen) next = arena[roff + ri];
  }
  int before = 0;  // sum of the deltas of events ordered before this one (agent-list order)
  int net = 0;
  int occ0 = 0;
  int j = first;
  for (;;) {
    const Event o = events[j];
    net += o.kind;
    if (o.agent < ev.agent) before += o.kind;
    if (o.aux < 0) {
      occ0 = -o.aux - 1;
      break;
    }
    j = o.aux;
  }
  if (first == e) lk.occ[ev.link] = occ0 + net;
  if (mine) {
    const int32_t count = max(occ0 + before + 1, 0);  // includes the entering agent itself (>= 1 in a consistent state)
    // host-tabulated max(0.1, 1/(1+alpha*(count/cap)**beta)) (traffic_manager.py:88-107), saturated at the floor
    int64_t idx = f_lo + (int64_t)count;
    if (idx >= f_hi) idx = f_hi - 1;
    const double f = lk.factor[idx];
    const double sp = __dmul_rn(__dmul_rn(v0, f), kph_to_mps);
    a.speed[li] = sp;
    a.elen[li] = len;
    a.next_link[li] = next;
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

// ---- standalone kernels (multi-rank exchange between A2 and C; per-kernel profiling) -----------------
template <bool DEMAND>
__global__ void __launch_bounds__(kAdvThreads, 6) k_advance(const __grid_constant__ TickArgs p, int32_t tick, double p_dep) {
  __shared__ AdvSmem sm;
  phase_advance<DEMAND>(p, tick, p_dep, sm);
}

__global__ void __launch_bounds__(256) k_inject(const __grid_constant__ TickArgs p, int32_t tick) {
  const int par = tick & 1;
  phase_inject(p, tick, min(p.ctr->n_pending, p.pend_cap), p.ctr->n_dep_req[par]);
}

// Events that arrived from other ranks (link, agent, kind): push them on the link lists.
__global__ void k_link_events(LinkArrays lk, Event* __restrict__ events, int n_events, int32_t tick) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_events; e += gridDim.x * blockDim.x) {
    const Event ev = events[e];
    emit_event(lk, events, e, ev.link, ev.agent, ev.kind, (uint32_t)tick | 0x80000000u);
  }
}

// n_events < 0: the local sub-buffers of this tick; otherwise `events` is one contiguous list.
__global__ void k_resolve(const __grid_constant__ TickArgs p, const Event* __restrict__ events, int n_events, int32_t tick) {
  const int par = tick & 1;
  __shared__ int s_pref[kEvSub + 1];
  if (n_events < 0) {
    phase_resolve_local(p, tick, s_pref);
  } else {
    phase_resolve(p.lk, p.a, p.ctr, events, 0, n_events, tick, p.kph_to_mps, p.rg, p.arena);
  }
}

// counters of the next tick (after k_resolve: nobody reads the other parity any more)
__global__ void k_reset_counters(Counters* ctr, int32_t tick) {
  const int par = tick & 1;
  if (threadIdx.x < kEvSub) ctr->n_events[par ^ 1][threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    ctr->n_dep_req[par ^ 1] = 0;
    ctr->n_pending = 0;
  }
}

// Gather the local sub-buffers into one contiguous list (what the ranks exchange); total in *n_out.
__global__ void k_compact_events(const Event* __restrict__ events, int32_t ev_cap, const Counters* ctr, int32_t tick,
                                 Event* __restrict__ out, int32_t* __restrict__ n_out) {
  const int par = tick & 1;
  int off = 0;
  for (int sub = 0; sub < kEvSub; ++sub) {
    const int n = ctr->n_events[par][sub];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
      out[off + e] = events[(int64_t)sub * ev_cap + e];
    off += n;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = off;
}

// ---- multi-rank exchange without host round trips -------------------------------------------------
// Send buffer of one rank: row 0 = header (link field = number of events), rows 1..cap = events.
__global__ void k_pack_events(const Event* __restrict__ events, int32_t ev_cap, Counters* ctr, int32_t tick,
                              Event* __restrict__ out, int32_t cap) {
  const int par = tick & 1;
  int off = 0;
  for (int sub = 0; sub < kEvSub; ++sub) {
    const int n = min(ctr->n_events[par][sub], ev_cap);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
      if (off + e < cap) out[1 + off + e] = events[(int64_t)sub * ev_cap + e];
    off += n;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    Event h;
    h.link = min(off, cap);
    h.agent = 0;
    h.kind = 0;
    h.aux = off;
    out[0] = h;
    if (off > cap) atomicOr(&ctr->overflow, kOvfExchange);
    atomicMax(&ctr->max_events, off);
  }
}

// Gathered buffers of all ranks: [world][cap + 1] records.  Flattened index -> (rank, event).
__device__ __forceinline__ int gathered_index(const Event* __restrict__ recv, int world, int cap, int f, int* s_pref) {
  int r = 0;
  while (r + 1 < world && s_pref[r + 1] <= f) ++r;
  return r * (cap + 1) + 1 + (f - s_pref[r]);
}

__global__ void k_link_gathered(LinkArrays lk, Event* __restrict__ recv, int world, int cap, int32_t tick) {
  __shared__ int s_pref[17];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < world; ++r) {
      s_pref[r] = acc;
      acc += recv[(int64_t)r * (cap + 1)].link;
    }
    s_pref[world] = acc;
  }
  __syncthreads();
  const int total = s_pref[world];
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    const int e = gathered_index(recv, world, cap, f, s_pref);
    const Event ev = recv[e];
    emit_event(lk, recv, e, ev.link, ev.agent, ev.kind, (uint32_t)tick | 0x80000000u);
  }
}

__global__ void k_resolve_gathered(const __grid_constant__ TickArgs p, const Event* __restrict__ recv, int world,
                                   int cap, int32_t tick) {
  __shared__ int s_pref[17];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < world; ++r) {
      s_pref[r] = acc;
      acc += recv[(int64_t)r * (cap + 1)].link;
    }
    s_pref[world] = acc;
  }
  __syncthreads();
  const int total = s_pref[world];
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x)
    resolve_one(p.lk, p.a, p.ctr, recv, gathered_index(recv, world, cap, f, s_pref), tick, p.kph_to_mps, p.rg, p.arena);
}

// ---- the persistent tick loop ---------------------------------------------------------------------
// One cooperative launch runs `n_ticks` ticks: 2 grid barriers per tick (3 when phase A2 has work).
__global__ void __launch_bounds__(kAdvThreads, 6) k_tick_loop(const __grid_constant__ TickArgs p, int32_t tick0, int32_t n_ticks, double p_dep) {
  __shared__ AdvSmem sm;
  __shared__ int s_pref[kEvSub + 1];
  cg::grid_group grid = cg::this_grid();
  Counters* ctr = p.ctr;
  const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long t0 = 0;
  for (int32_t k = 0; k < n_ticks; ++k) {
    const int32_t tick = tick0 + k;
    const int par = tick & 1;
    if (lead) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (p.demand)
      phase_advance<true>(p, tick, p_dep, sm);
    else
      phase_advance<false>(p, tick, p_dep, sm);
    grid.sync();
    if (lead) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      ctr->phase_ns[0] += t1 - t0;
      t0 = t1;
    }
    const int n_pending = min(ctr->n_pending, p.pend_cap), n_urgent = ctr->n_dep_req[par];
    if (n_pending | n_urgent) {  // uniform across the grid
      phase_inject(p, tick, n_pending, n_urgent);
      grid.sync();
      if (lead) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        ctr->phase_ns[1] += t1 - t0;
        t0 = t1;
      }
    }
    phase_resolve_local(p, tick, s_pref);
    if (blockIdx.x == 0 && threadIdx.x < kEvSub) ctr->n_events[par ^ 1][threadIdx.x] = 0;
    if (lead) ctr->n_dep_req[par ^ 1] = 0;
    grid.sync();
    if (lead) {
      ctr->n_pending = 0;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      ctr->phase_ns[2] += t1 - t0;
    }
  }
}

// ---- the persistent tick loop over NVLink peer memory ---------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Every block waits until all ranks have published `tick` in this row of the local flags; the low
// words (event counts) go to s_val.  A rank that never shows up (crashed peer) traps after 60 s
// instead of hanging the GPU.
__device__ __forceinline__ void p2p_wait(const unsigned long long* row, int world, uint32_t tick, int* s_val) {
  if (threadIdx.x < world) {
    const unsigned long long t0 = global_ns();
    unsigned long long v;
    for (;;) {
      v = ld_acquire_sys(row + threadIdx.x);
      if ((uint32_t)(v >> 32) >= tick) break;
      if (global_ns() - t0 > 60000000000ULL) {
        printf("drift_b200: rank flag %d did not reach tick %u within 60 s\n", (int)threadIdx.x, tick);
        __trap();
      }
      __nanosleep(64);
    }
    if (s_val) s_val[threadIdx.x] = (int)(uint32_t)(v & 0xffffffffu);
  }
  __syncthreads();
}

__device__ __forceinline__ void p2p_home(const P2PArgs& x, uint32_t agent, int& home, uint32_t& local) {
  const uint32_t big = (x.shard_base + 1u) * x.shard_rem;  // agents held by the ranks with one extra agent
  if (agent < big) {
    home = (int)(agent / (x.shard_base + 1u));
    local = agent - (uint32_t)home * (x.shard_base + 1u);
  } else {
    const uint32_t r = agent - big;
    const uint32_t q = r / x.shard_base;
    home = (int)(x.shard_rem + q);
    local = r - q * x.shard_base;
  }
}

__device__ __forceinline__ void p2p_send_one(const TickArgs& p, const P2PArgs& x, int par, uint32_t tick, int32_t link,
                                             uint32_t agent, int32_t kind) {
  const int o = (int)((uint32_t)link % (uint32_t)x.world);
  const int pos = atomicAdd(&x.out_cnt[o], 1);
  if (pos < x.cap) {
    Event e;
    e.link = link;
    e.agent = agent;
    e.kind = kind;
    e.aux = event_check(tick, link, agent, kind);
    *reinterpret_cast<int4*>(x.inbox[o] + ((int64_t)par * x.world + x.rank) * x.cap + pos) = *reinterpret_cast<const int4*>(&e);
  } else {  // see phase_advance: a dropped entry gets the free-flow speed so that nobody waits for its owner
    atomicOr(&p.ctr->overflow, kOvfExchange);
    if (kind > 0) p.a.speed[agent - p.a.base] = __dmul_rn(p.lk.v0[link], p.kph_to_mps);
  }
}

// phase A2 with peer delivery (host-committed departures and in-tick fallback routes)
__device__ __noinline__ void phase_inject_p2p(const TickArgs& p, const P2PArgs& x, int32_t tick, int n_pending,
                                              int n_urgent) {
  const AgentArrays& a = p.a;
  auto start = [&](int32_t i, int64_t off, int32_t len) {
    if (a.state[i] == kMoving) return;
    const int32_t link = p.arena[off];
    a.state[i] = kMoving;
    a.route_off[i] = off;
    a.route_len[i] = len;
    a.route_idx[i] = 0;
    a.dist[i] = 0.0;
    a.cur_link[i] = link;
    a.next_link[i] = len > 1 ? p.arena[off + 1] : -1;
    a.elen[i] = p.lk.len[link];
    a.speed[i] = kSpeedPending;
    p2p_send_one(p, x, tick & 1, (uint32_t)tick, link, a.base + (uint32_t)i, 1);
  };
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = gtid; q < n_pending; q += gsz)
    if (p.pend_len[q] > 0) start(p.pend_agent[q], p.pend_off[q], p.pend_len[q]);
  if (n_urgent > 0 && gtid == 0) p.ctr->n_fallback += (unsigned long long)n_urgent;
  if (n_urgent > 0 && gtid < p.fsc.workers) {
    uint32_t epoch = p.fsc.epoch[gtid];
    for (int r = (int)gtid; r < n_urgent; r += p.fsc.workers) {
      const int32_t i = p.dm.dep_req[r];
      const int head = a.chain_head[i];
      const int32_t s = a.cur_node[i];
      const int32_t t = a.chain_dst[(int64_t)i * a.chain_cap + head];
      a.chain_head[i] = (uint8_t)((head + 1) % a.chain_cap);
      a.chain_cnt[i] -= 1;
      a.trip_count[i] += 1;
      a.cur_node[i] = t;
      const RouteResult rr = route_nx_one(p.g, p.w_fwd, p.w_bwd, s, t, p.fsc, (int)gtid, epoch, p.arena, p.arena_cap,
                                          p.ctr, p.link_cost);
      const unsigned long long slot = atomicAdd(&p.ctr->n_departures, 1ULL);
      if (slot - p.dm.log_drained < p.dm.log_cap) {
        const unsigned long long k = slot % p.dm.log_cap;
        p.dm.log_agent[k] = i;
        p.dm.log_tick[k] = tick;
        p.dm.log_src[k] = s;
        p.dm.log_dst[k] = t;
        p.dm.log_status[k] = rr.status;
        p.dm.log_off[k] = rr.off;
        p.dm.log_len[k] = rr.n_links;
        p.dm.log_epoch[k] = p.dm.epoch;
      } else {
        atomicOr(&p.ctr->overflow, kOvfDepLog);
      }
      if (rr.status == DRIFT_ROUTE_OK) start(i, rr.off, rr.n_links);
    }
    p.fsc.epoch[gtid] = epoch;
  }
}

// Owner side of phase C: ordered count on an owned link -> entry speed, written into the home
// rank's agent arrays (peer store); the list head writes the end-of-tick occupancy of the link.
__device__ __forceinline__ bool resolve_one_p2p(const TickArgs& p, const P2PArgs& x, const Event* __restrict__ events,
                                                int e) {
  const LinkArrays& lk = p.lk;
  const Event ev = events[e];
  const int first = (int)(lk.head[ev.link] & 0xffffffffu);
  double v0 = 0.0;
  int64_t f_lo = 0, f_hi = 1;
  if (ev.kind > 0) {  // loaded before the list walk: independent of it
    const LinkRec r = lk.rec[ev.link];
    v0 = r.v0;
    f_lo = r.f_lo;
    f_hi = r.f_hi;
  }
  int before = 0, net = 0, occ0 = 0;
  int j = first;
  for (;;) {
    const Event o = events[j];
    net += o.kind;
    if (o.agent < ev.agent) before += o.kind;
    if (o.aux < 0) {
      occ0 = -o.aux - 1;
      break;
    }
    j = o.aux;
  }
  if (first == e) lk.occ[ev.link] = occ0 + net;
  if (ev.kind > 0) {
    int home;
    uint32_t li;
    p2p_home(x, ev.agent, home, li);
    const int32_t count = max(occ0 + before + 1, 0);  // includes the entering agent itself
    int64_t idx = f_lo + (int64_t)count;
    if (idx >= f_hi) idx = f_hi - 1;
    x.speed[home][li] = __dmul_rn(__dmul_rn(v0, lk.factor[idx]), p.kph_to_mps);  // the home rank waits for it per agent
    return home != x.rank;
  }
  return false;
}

// One cooperative launch per rank runs `n_ticks` ticks; the ranks meet twice per tick through flags
// in peer memory (events delivered / links resolved), never through the host.
__global__ void __launch_bounds__(kAdvThreads, 6) k_tick_loop_p2p(const __grid_constant__ TickArgs p,
                                                                  const __grid_constant__ P2PArgs x, int32_t tick0,
                                                                  int32_t n_ticks, double p_dep) {
  __shared__ AdvSmem sm;
  __shared__ P2PSmem xs;
  __shared__ int s_in[kMaxRanks], s_pref[kMaxRanks + 1];
  cg::grid_group grid = cg::this_grid();
  Counters* ctr = p.ctr;
  const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
  Event* inbox = x.inbox[x.rank];
  unsigned long long* my_flags = x.flags[x.rank];
  unsigned long long t0 = 0;
  for (int32_t k = 0; k < n_ticks; ++k) {
    const int32_t tick = tick0 + k;
    const int par = tick & 1;
    unsigned long long d0 = 0;
    auto lap = [&](int slot) {  // block 0's own clock, for the breakdown drift_p2p_phase_detail reports
      if (lead) {
        const unsigned long long now = global_ns();
        ctr->p2p_ns[slot] += now - d0;
        d0 = now;
      }
    };
    if (lead) t0 = d0 = global_ns();
    if (p.demand)
      phase_advance<true, true>(p, tick, p_dep, sm, &x, &xs);
    else
      phase_advance<false, true>(p, tick, p_dep, sm, &x, &xs);
    lap(0);
    // no fence for the peer stores: the records are self-validating (event_check), the owner polls them
    lap(1);
    grid.sync();
    lap(2);
    const int n_pending = min(ctr->n_pending, p.pend_cap), n_urgent = ctr->n_dep_req[par];
    if (n_pending | n_urgent) {  // uniform across the grid
      phase_inject_p2p(p, x, tick, n_pending, n_urgent);
      grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x < x.world) {  // publish: this rank's events are in owner o's inbox
      const int o = threadIdx.x;
      int cnt = x.out_cnt[o];
      x.out_cnt[o] = 0;
      atomicMax(&ctr->max_events, cnt);
      if (cnt > x.cap) cnt = x.cap;  // overflow already flagged by the writers
      // the count only: a relaxed store, nothing is ordered behind it (the records carry their own check)
      st_relaxed_sys(x.flags[o] + par * kMaxRanks + x.rank, ((unsigned long long)(uint32_t)tick << 32) | (uint32_t)cnt);
    }
    if (lead) {
      ctr->n_dep_req[par ^ 1] = 0;
      const unsigned long long t1 = global_ns();
      ctr->phase_ns[0] += t1 - t0;
      t0 = t1;
    }
    lap(3);
    p2p_wait(my_flags + par * kMaxRanks, x.world, (uint32_t)tick, s_in);
    lap(4);
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int r = 0; r < x.world; ++r) {
        s_pref[r] = acc;
        acc += s_in[r];
      }
      s_pref[x.world] = acc;
    }
    __syncthreads();
    const int total = s_pref[x.world];
    // link the events of the owned links (all ranks' regions of the inbox)
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
      int r = 0;
      while (r + 1 < x.world && s_pref[r + 1] <= f) ++r;
      const int e = (par * x.world + r) * x.cap + (f - s_pref[r]);
      // written by a peer, without a fence: poll (past L1) until the record of THIS tick has landed whole
      int4 raw;
      unsigned long long w0 = 0;
      for (;;) {
        asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(inbox + e) : "memory");
        if (raw.w == event_check((uint32_t)tick, raw.x, (uint32_t)raw.y, raw.z) && (raw.z == 1 || raw.z == -1)) break;
        const unsigned long long now = global_ns();
        if (!w0) w0 = now;
        if (now - w0 > 60000000000ULL) {
          printf("drift_b200: an event record of tick %d did not arrive within 60 s\n", tick);
          __trap();
        }
      }
      emit_event(p.lk, inbox, e, raw.x, (uint32_t)raw.y, raw.z, (uint32_t)tick);
    }
    lap(5);
    grid.sync();
    lap(6);
    if (lead) {
      ctr->n_pending = 0;
      const unsigned long long t1 = global_ns();
      ctr->phase_ns[1] += t1 - t0;  // exchange wait + list building
      t0 = t1;
    }
    bool any_remote = false;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
      int r = 0;
      while (r + 1 < x.world && s_pref[r + 1] <= f) ++r;
      const bool remote = resolve_one_p2p(p, x, inbox, (par * x.world + r) * x.cap + (f - s_pref[r]));
      any_remote |= remote;
    }
    // the speeds stored into peers need no fence either: a home rank polls the value itself (kSpeedPending /
    // kSpeedDeferred are not speeds).  Only the last tick of a launch fences them (one thread per block, after a block
    // barrier), so that they have landed before the launch-end barrier below lets any rank return to its host
    if (k == n_ticks - 1 && __syncthreads_or(any_remote ? 1 : 0) && threadIdx.x == 0) __threadfence_system();
    if (lead) ctr->phase_ns[2] += global_ns() - t0;
    lap(7);
    // No all-rank barrier here: the next advance waits per agent for a pending entry speed, the inbox and
    // the event flags are double-buffered by tick parity, and a rank cannot run two ticks ahead of a peer
    // (it needs that peer's events of the next tick first).
  }
  // launch end: every rank has delivered the entry speeds of the last tick before any rank returns to its host
  grid.sync();
  const uint32_t last = (uint32_t)(tick0 + n_ticks - 1);
  if (blockIdx.x == 0 && threadIdx.x < x.world)
    st_release_sys(x.flags[threadIdx.x] + 2 * kMaxRanks + x.rank, (unsigned long long)last << 32);
  p2p_wait(my_flags + 2 * kMaxRanks, x.world, last, nullptr);
  // every entry speed has landed: apply the advances still owed (deferred at the last tick), so that the state a
  // host reads between launches is the reference's
  {
    const AgentArrays& a = p.a;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x)
      if (a.state[i] == kMoving && owes_advance(a.dist[i]))
        a.dist[i] = __dadd_rn(0.0, __dmul_rn(__ldcg(a.speed + i), p.delta));
  }
}

// ---- auxiliary passes -------------------------------------------------------------------------------
// K3': per-link volume recount = histogram of cur_link over moving agents
// (traffic_manager.py:122-135: total_agents_on_roads / per-link list lengths).  A tile of 1024 agents is staged in
// a shared-memory histogram (open addressing on the link id) before it touches HBM: lanes of a warp that sit on
// the same link are merged with __match_any_sync and only the group leader inserts, the block then issues ONE
// global atomic per distinct link of the tile.  Agents on a hot link (a hub approach, a gridlocked arterial)
// therefore cost a shared-memory add instead of a same-address L2 atomic, which serialises chip-wide.
constexpr int kRecThreads = 256;
constexpr int kRecPerThread = 4;
constexpr int kRecSlots = 2048;  // >= 2 x agents of a tile: load factor <= 0.5

__global__ void __launch_bounds__(kRecThreads) k_recount(AgentArrays a, int32_t* __restrict__ hist) {
  __shared__ int32_t s_key[kRecSlots];
  __shared__ int32_t s_cnt[kRecSlots];
  const int64_t tile_agents = (int64_t)kRecThreads * kRecPerThread;
  const int64_t tiles = (a.n + tile_agents - 1) / tile_agents;
  const int lane = threadIdx.x & 31;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    for (int i = threadIdx.x; i < kRecSlots; i += kRecThreads) {
      s_key[i] = -1;
      s_cnt[i] = 0;
    }
    __syncthreads();
    const int64_t i0 = (tile * kRecThreads + threadIdx.x) * kRecPerThread;  // arrays are padded to whole tiles
    uint32_t st4 = 0;
    int4 l4 = make_int4(-1, -1, -1, -1);
    if (i0 < a.n) {
      st4 = *reinterpret_cast<const uint32_t*>(a.state + i0);
      if (st4) l4 = *reinterpret_cast<const int4*>(a.cur_link + i0);
    }
    const int32_t lk[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
    for (int j = 0; j < kRecPerThread; ++j) {
      const bool on = ((st4 >> (8 * j)) & 0xffu) == kMoving && lk[j] >= 0 && i0 + j < a.n;
      const unsigned active = __ballot_sync(0xffffffffu, on);
      if (!on) continue;
      const unsigned peers = __match_any_sync(active, lk[j]);
      if ((peers & ((1u << lane) - 1u)) != 0) continue;  // not the lowest lane of its group
      int slot = (int)(((uint32_t)lk[j] * 2654435761u) >> 21);  // 11 bits
      for (;;) {
        const int32_t k = s_key[slot];
        if (k == lk[j]) break;
        if (k == -1) {
          const int32_t old = atomicCAS(&s_key[slot], -1, lk[j]);
          if (old == -1 || old == lk[j]) break;
        }
        slot = (slot + 1) & (kRecSlots - 1);
      }
      atomicAdd(&s_cnt[slot], __popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kRecSlots; i += kRecThreads)
      if (s_key[i] >= 0) atomicAdd(&hist[s_key[i]], s_cnt[i]);
    __syncthreads();
  }
}

__global__ void k_pack_linkrec(LinkArrays lk, LinkRec* __restrict__ rec) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < lk.L; l += gridDim.x * blockDim.x) {
    const int32_t c = lk.cls[l];
    LinkRec r;
    r.v0 = lk.v0[l];
    r.len = lk.len[l];
    r.f_lo = (int32_t)lk.cls_off[c];
    r.f_hi = (int32_t)lk.cls_off[c + 1];
    r.pad[0] = r.pad[1] = 0;
    rec[l] = r;
  }
}

// K4 (extension): tt[l] = t0[l] / factor(occ[l]) -- the BPR travel-time form of traffic_manager.py:92-107, one
// fused elementwise pass over the link arrays: 4 links per thread (one 16-byte load of the volumes, two of the
// free-flow times, two 16-byte stores), the factor table is only touched for links that hold traffic
// (an empty link has factor 1.0: traffic_manager.py:88-89).  24 B per link: occ 4 R + cap class 4 R + t0 8 R + tt 8 W.
__global__ void __launch_bounds__(256) k_travel_times(LinkArrays lk) {
  const int64_t quads = ((int64_t)lk.L + 3) / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l0 = q * 4;
    if (l0 + 4 <= lk.L) {
      const int4 c = *reinterpret_cast<const int4*>(lk.occ + l0);
      const double2 t01 = *reinterpret_cast<const double2*>(lk.t0 + l0);
      const double2 t23 = *reinterpret_cast<const double2*>(lk.t0 + l0 + 2);
      const double f0 = c.x == 0 ? 1.0 : bpr_factor_lookup(lk, (int32_t)l0, c.x);
      const double f1 = c.y == 0 ? 1.0 : bpr_factor_lookup(lk, (int32_t)l0 + 1, c.y);
      const double f2 = c.z == 0 ? 1.0 : bpr_factor_lookup(lk, (int32_t)l0 + 2, c.z);
      const double f3 = c.w == 0 ? 1.0 : bpr_factor_lookup(lk, (int32_t)l0 + 3, c.w);
      *reinterpret_cast<double2*>(lk.tt + l0) = make_double2(__ddiv_rn(t01.x, f0), __ddiv_rn(t01.y, f1));
      *reinterpret_cast<double2*>(lk.tt + l0 + 2) = make_double2(__ddiv_rn(t23.x, f2), __ddiv_rn(t23.y, f3));
    } else {
      for (int64_t l = l0; l < lk.L; ++l) {
        const int32_t c = lk.occ[l];
        lk.tt[l] = __ddiv_rn(lk.t0[l], c == 0 ? 1.0 : bpr_factor_lookup(lk, (int32_t)l, c));
      }
    }
  }
}

// Top up every agent's chain ring from a block of host-generated candidate destinations (k per agent,
// row-major).  The k candidates of an agent are one unit (e.g. an out leg and its return leg): they are
// taken together when the ring has room for all of them, else the whole row is dropped -- a partial
// take would splice chains (out, out, ...) that no s-t model produces.
__global__ void k_refill_chains(AgentArrays a, const int32_t* __restrict__ cand, int32_t k) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    int cnt = a.chain_cnt[i];
    if (cnt + k > a.chain_cap) continue;
    const int head = a.chain_head[i];
    for (int j = 0; j < k; ++j, ++cnt) a.chain_dst[i * a.chain_cap + (head + cnt) % a.chain_cap] = cand[i * k + j];
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

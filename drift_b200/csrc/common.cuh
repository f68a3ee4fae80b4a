// drift_b200 -- shared device/host definitions (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/drift_b200.h"

namespace drift {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs
constexpr uint8_t kWaiting = 0;
constexpr uint8_t kMoving = 1;

// One link event of a tick: an agent leaving or entering a link.  16 bytes so that a warp
// writes 512 contiguous bytes.  The events of one link form a singly linked list threaded through
// `aux` (head in LinkArrays::head); list order is arbitrary, the Gauss-Seidel order of the
// reference (agent-list order, lib/simulation_thread.py:299-303) is recovered from `agent`.
struct __align__(16) Event {
  int32_t link;
  uint32_t agent;  // global agent id
  int32_t kind;    // +1 enter, -1 leave
  int32_t aux;     // >= 0: next event of the same link; < 0: last one, carries -(occ_start) - 1
};

struct HeapItem {  // (dist, push counter, node): nx:weighted.py heap entries
  double dist;
  uint32_t cnt;
  int32_t node;
};

// The event buffer is split into kEvSub sub-buffers with their own counters so that the warps of
// phase A do not all hit one address when they reserve slots.
constexpr int kEvSub = 16;

struct Counters {
  int32_t n_events[2][kEvSub];  // [tick parity][sub-buffer]: tick k resets the slots of tick k+1
  int32_t n_dep_req[2];   // departures of this tick that found no prefetched route
  int32_t n_pending;      // pending departures (committed routes) for the next tick
  int32_t overflow;       // bit flags, see kOvf*
  int32_t n_plan;         // route requests of the current planning batch
  int32_t max_events;     // largest per-tick event count of this rank since the last reset (exchange sizing)
  int32_t pad0;
  unsigned long long n_arrivals;    // total appended to the arrival ring
  unsigned long long n_departures;  // total appended to the departure log ring
  unsigned long long n_entries;     // total appended to the entry trace
  unsigned long long arena_cursor;  // route arena bump pointer (link ids)
  unsigned long long stage_cursor;
  long long moving;                 // reductions for drift_stats
  long long on_roads;
  long long links_with_traffic;
  long long total_capacity;
  unsigned long long phase_ns[4];   // persistent kernel: accumulated %globaltimer time per phase
  unsigned long long tree_relaxed;  // tree builder: edges relaxed / relaxations that improved / nodes expanded
  unsigned long long tree_improved;
  unsigned long long tree_expanded;
  unsigned long long tree_jobs;
  unsigned long long tree_iters;      // near-list waves (block barriers) over all jobs
  unsigned long long tree_far_visits; // far-list entries scanned by threshold advances
  unsigned long long n_fallback;      // departures routed inside the tick because no planned route was ready
  unsigned long long label_entries;   // hub-label router: label entries read by the queries (roofline accounting)
  unsigned long long label_queries;
  unsigned long long n_tie_fallback;  // batched router: queries re-run by the NetworkX-exact router (possible ties)
  // k_tick_loop_p2p, block 0's clock per tick: advance | block fence | grid barrier | publish | wait for the peers'
  // flags | list building | grid barrier | resolve + fence
  unsigned long long p2p_ns[8];
};

constexpr int kOvfEvents = 1;
constexpr int kOvfArrivals = 2;
constexpr int kOvfArena = 4;
constexpr int kOvfHeap = 8;
constexpr int kOvfDepLog = 16;
constexpr int kOvfEntries = 32;
constexpr int kOvfStage = 64;
constexpr int kOvfExchange = 128;
constexpr int kOvfPending = 256;

struct AgentArrays {
  uint8_t* state;       // [Npad] kWaiting / kMoving
  double* dist;         // [Npad] distance_on_edge
  double* speed;        // [Npad] m/s, fixed while on a link
  double* elen;         // [Npad] edge_length of the current link
  int32_t* cur_link;    // [Npad] -1 when not on a link
  int32_t* next_link;   // [Npad] link after cur_link on the route, -1 = cur_link is the last one
  int32_t* route_idx;   // [Npad] path_index
  int32_t* route_len;   // [Npad] hops of the current route
  int64_t* route_off;   // [Npad] start of the route in the arena
  int32_t* trip_count;  // [Npad]
  int32_t* cur_node;    // [Npad] Agent.target: where the next trip starts
  // host-generated trip chain: a small ring of upcoming destinations per agent, topped up by
  // drift_push_trips; entry k of agent i is chain_dst[i * chain_cap + (chain_head[i] + k) % chain_cap]
  int32_t* chain_dst;    // [N][chain_cap]
  uint8_t* chain_head;   // [Npad]
  uint8_t* chain_cnt;    // [Npad] destinations available
  int32_t chain_cap;
  // prefetched routes of the next two trips (device-driven demand): slot k at [k * npad + i]
  int64_t* next_off;     // [2][Npad] arena offset
  int32_t* next_len;     // [2][Npad] hops
  int32_t* next_status;  // [2][Npad] 0 = none, 1 + DRIFT_ROUTE_* otherwise
  int64_t npad;
  int64_t n;
  uint32_t base;  // global id of local agent 0
};

// What resolving an entry needs to know about a link, in one 32-byte sector: free speed, length and the
// range of the link's capacity class in the BPR factor table (instead of cls -> cls_off -> factor).
struct __align__(32) LinkRec {
  double v0;
  double len;
  int32_t f_lo;  // first entry of the class in `factor`
  int32_t f_hi;  // one past its last entry (the saturated 10 % floor)
  int32_t pad[2];
};

struct LinkArrays {
  const double* len;
  const double* v0;
  const double* t0;
  const int32_t* cap;
  const int32_t* u;
  const int32_t* v;
  int32_t* occ;
  unsigned long long* head;  // [L] (tick << 32 | event index): head of the link's event list this tick
  double* tt;  // congestion-updated travel time (extension)
  const int32_t* cls;
  const double* factor;     // concatenated per-class tables
  const int64_t* cls_off;   // [n_class+1]
  const LinkRec* rec;       // [L] packed copy of v0 / len / factor range (built by drift_set_bpr)
  int32_t L;
};

struct DemandArgs {  // device-driven demand (drift_run)
  uint64_t seed;
  double p_dep;                // hourly[h] * 4 * delta / 3600 (agent.py:180,259)
  int32_t* dep_req;            // departures that found no prefetched route (routed within the tick)
  int32_t* log_agent;          // departure log ring
  int32_t* log_tick;
  int32_t* log_src;
  int32_t* log_dst;
  int32_t* log_status;
  long long* log_off;          // where the departure's route sits in the arena (valid until the next compaction) ...
  int32_t* log_len;            // ... its hops ...
  int32_t* log_epoch;          // ... and the number of compactions so far (trip records: drift_export_logged_routes)
  int32_t epoch;
  unsigned long long log_cap;
  unsigned long long log_drained;
};

struct Rings {
  int32_t* arr_agent;
  int32_t* arr_tick;
  unsigned long long arr_cap;
  unsigned long long arr_drained;
  int32_t* ent_tick;
  int32_t* ent_agent;
  int32_t* ent_link;
  double* ent_speed;
  unsigned long long ent_cap;  // 0 = trace off
  unsigned long long ent_drained;
};

// ---- Philox4x32-10 counter-based generator (Salmon et al., SC'11; public algorithm) ----
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0;
  c1 = n1;
  c2 = n2;
  c3 = n3;
}

// One Philox4x32-10 block = 128 random bits = the draws of TWO agents: the counter is
// (agent >> 1, tick), the even agent takes words 0-1, the odd one words 2-3.
__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint32_t pair, uint32_t tick, uint32_t out[4]) {
  uint32_t c0 = pair, c1 = tick, c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// 53 random bits the way MT19937's genrand_res53 (Python's random.random()) builds them:
// (a >> 5, b >> 6) -> a * 2^26 + b; the uniform double is that integer * 2^-53.
__device__ __forceinline__ unsigned long long bits53(uint32_t a, uint32_t b) {
  return ((unsigned long long)(a >> 5) << 26) | (unsigned long long)(b >> 6);
}

// u < p for u = x * 2^-53  <=>  x < ceil(p * 2^53) (scaling by 2^53 is exact): the per-tick
// Bernoulli test (lib/agent.py:180) as one integer compare.
__device__ __forceinline__ unsigned long long departure_threshold(double p) {
  if (!(p > 0.0)) return 0ULL;
  const double s = ceil(__dmul_rn(p, 9007199254740992.0));
  return s >= 18446744073709551615.0 ? ~0ULL : (unsigned long long)s;
}

__device__ __forceinline__ unsigned long long philox_bits53(uint64_t seed, uint32_t agent, uint32_t tick) {
  uint32_t o[4];
  philox4x32_10(seed, agent >> 1, tick, o);
  return (agent & 1u) ? bits53(o[2], o[3]) : bits53(o[0], o[1]);
}

__device__ __forceinline__ double philox_uniform53(uint64_t seed, uint32_t agent, uint32_t tick) {
  return __dmul_rn((double)philox_bits53(seed, agent, tick), 1.0 / 9007199254740992.0);
}

}  // namespace drift

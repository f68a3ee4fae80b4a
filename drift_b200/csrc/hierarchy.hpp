// drift_b200 -- contraction hierarchy of the static road graph (host-side preprocessing).
//
// The reference routes every departure with nx.shortest_path on the `travel_time` attribute
// (lib/agent.py:94), which the loader writes once and never updates (lib/graph_loader.py:759-766): the
// weights are static, so the graph can be preprocessed once per run.  A contraction hierarchy (Geisberger,
// Sanders, Schultes, Delling: "Contraction Hierarchies", WEA 2008; public algorithm, restated here) orders
// the nodes, removes them one by one and inserts a shortcut (u, w) wherever the removed node was the only
// shortest connection between two of its remaining neighbours.  A shortest path then always exists as an
// "up-down" path in the node order, so a query only climbs: a few hundred to a few thousand labels instead of
// a corridor of the whole graph -- small enough to live in the shared memory of one thread block
// (route_kernels.cuh: k_ch_query).
//
// Output layout (what the device holds):
//   up[v]    arcs v -> w with rank[w] > rank[v]          (forward search climbs these)
//   down[v]  arcs u -> v with rank[u] > rank[v], stored at v (backward search climbs these)
//   per arc: other end, fp64 weight, hops (original links below it) and either the original link id
//   (b == -1) or its two children: a = position in down[mid] of (u -> mid), b = position in up[mid] of
//   (mid -> w) -- a shortcut's middle node is contracted before both ends, so its first half is always a
//   down arc of the middle node and its second half an up arc of it.
// Shortcut weights are fp64 sums in contraction order; the route handed to the tick kernels is the unpacked
// link sequence and its cost is re-summed left to right, so wherever the shortest path is unique (continuous
// weights) links and cost equal NetworkX's bit for bit.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <queue>
#include <vector>

namespace drift {

struct HierArcs {               // one CSR side (up or down)
  std::vector<int32_t> rowptr;  // [V+1]
  std::vector<int32_t> other;   // head of an up arc / tail of a down arc
  std::vector<double> w;
  std::vector<int32_t> a, b;    // b == -1: a = original link; else children (see above)
  std::vector<int32_t> hops;
};

struct Hierarchy {
  int32_t V = 0;
  std::vector<int32_t> rank;   // contraction order of every node
  std::vector<int32_t> level;  // longest chain of contracted neighbours below the node
  HierArcs up, down;
  int64_t shortcuts = 0;
  int32_t max_level = 0;
};

namespace hier_detail {

struct Adj {
  int32_t to;
  int32_t arc;
  double w;
};

struct ArcRec {
  int32_t u, v;
  double w;
  int32_t a, b;  // b == -1: original link a; else child arc ids (u -> mid, mid -> v)
  int32_t hops;
};

struct Witness {  // bounded Dijkstra in the remaining graph, time-stamped labels
  std::vector<double> dist;
  std::vector<uint32_t> stamp;
  uint32_t epoch = 0;
  typedef std::pair<double, int32_t> QE;
  std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
  void init(int32_t V) {
    dist.assign(V, 0.0);
    stamp.assign(V, 0u);
  }
  // labels of everything within `limit` of `src`, never through `skip`; at most `max_settled` nodes
  void run(const std::vector<std::vector<Adj>>& out, int32_t src, int32_t skip, double limit, int max_settled) {
    ++epoch;
    while (!pq.empty()) pq.pop();
    dist[src] = 0.0;
    stamp[src] = epoch;
    pq.push(QE(0.0, src));
    int settled = 0;
    while (!pq.empty()) {
      const QE top = pq.top();
      pq.pop();
      const int32_t x = top.second;
      if (top.first > dist[x]) continue;
      if (top.first > limit) break;
      if (++settled > max_settled) break;
      for (const Adj& e : out[x]) {
        if (e.to == skip) continue;
        const double nd = top.first + e.w;
        if (stamp[e.to] != epoch || nd < dist[e.to]) {
          stamp[e.to] = epoch;
          dist[e.to] = nd;
          pq.push(QE(nd, e.to));
        }
      }
    }
  }
  bool reached(int32_t x, double need) const { return stamp[x] == epoch && dist[x] <= need; }
};

struct Shortcut {
  int32_t u, w;
  double wt;
  int32_t arc_in, arc_out;
};

}  // namespace hier_detail

// Build the hierarchy of the graph given as a forward CSR ((u, v) pairs are unique; self loops are ignored).
// `witness_settled` bounds each witness search (a smaller bound is faster and only adds unnecessary shortcuts).
struct HierParams {
  int witness_settled = 100;  // bound of each witness search
  int w_ed = 2, w_cn = 1, w_level = 2;  // node priority = w_ed * edge difference + w_cn * contracted neighbours + w_level * level
  double time_limit_s = 0;  // > 0: give up after this many seconds (grid-like graphs have no hierarchy to find)
};

// Returns false when the time limit was hit (H is then unusable).
inline bool build_hierarchy(int32_t V, const int32_t* rowptr, const int32_t* col, const double* w, const int32_t* link,
                            const HierParams& prm, Hierarchy* H) {
  const int witness_settled = prm.witness_settled;
  using namespace hier_detail;
  std::vector<std::vector<Adj>> out(V), in(V);
  std::vector<ArcRec> arcs;
  arcs.reserve((size_t)rowptr[V] * 2);
  for (int32_t u = 0; u < V; ++u)
    for (int32_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {
      const int32_t v = col[e];
      if (v == u) continue;
      const int32_t id = (int32_t)arcs.size();
      arcs.push_back(ArcRec{u, v, w[e], link[e], -1, 1});
      out[u].push_back(Adj{v, id, w[e]});
      in[v].push_back(Adj{u, id, w[e]});
    }
  std::vector<uint8_t> contracted(V, 0);
  std::vector<int32_t> level(V, 0), cn(V, 0), ed(V, 0), rank(V, -1);
  Witness ws;
  ws.init(V);
  std::vector<Shortcut> sc;

  // shortcuts that contracting v would need right now
  auto simulate = [&](int32_t v, std::vector<Shortcut>& res) {
    res.clear();
    if (out[v].empty() || in[v].empty()) return;
    double wmax = 0.0;
    for (const Adj& o : out[v]) wmax = std::max(wmax, o.w);
    for (const Adj& i : in[v]) {
      bool any = false;
      for (const Adj& o : out[v]) any |= o.to != i.to;
      if (!any) continue;
      ws.run(out, i.to, v, i.w + wmax, witness_settled);
      for (const Adj& o : out[v]) {
        if (o.to == i.to) continue;
        const double need = i.w + o.w;
        if (!ws.reached(o.to, need)) res.push_back(Shortcut{i.to, o.to, need, i.arc, o.arc});
      }
    }
  };
  auto priority = [&](int32_t v) {
    return (int64_t)prm.w_ed * ed[v] + (int64_t)prm.w_cn * cn[v] + (int64_t)prm.w_level * level[v];
  };

  typedef std::pair<int64_t, int32_t> PE;
  std::priority_queue<PE, std::vector<PE>, std::greater<PE>> pq;
  std::vector<int64_t> cur(V);
  for (int32_t v = 0; v < V; ++v) {
    simulate(v, sc);
    ed[v] = (int32_t)sc.size() - (int32_t)(out[v].size() + in[v].size());
    cur[v] = priority(v);
    pq.push(PE(cur[v], v));
  }
  auto erase_to = [](std::vector<Adj>& lst, int32_t node) {
    for (size_t k = 0; k < lst.size(); ++k)
      if (lst[k].to == node) {
        lst[k] = lst.back();
        lst.pop_back();
        return;
      }
  };
  std::vector<std::vector<int32_t>> up_of(V), down_of(V);  // final arcs stored at their lower end
  int32_t next_rank = 0;
  const auto t_begin = std::chrono::steady_clock::now();
  while (!pq.empty()) {
    const PE top = pq.top();
    pq.pop();
    const int32_t v = top.second;
    if (contracted[v] || top.first != cur[v]) continue;
    if (prm.time_limit_s > 0 && (next_rank & 1023) == 0 &&
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() > prm.time_limit_s)
      return false;
    simulate(v, sc);
    ed[v] = (int32_t)sc.size() - (int32_t)(out[v].size() + in[v].size());
    const int64_t p = priority(v);
    if (!pq.empty() && p > pq.top().first) {  // lazy update: somebody else is cheaper now
      cur[v] = p;
      pq.push(PE(p, v));
      continue;
    }
    // contract v
    contracted[v] = 1;
    rank[v] = next_rank++;
    for (const Adj& i : in[v]) {
      down_of[v].push_back(i.arc);
      erase_to(out[i.to], v);
    }
    for (const Adj& o : out[v]) {
      up_of[v].push_back(o.arc);
      erase_to(in[o.to], v);
    }
    for (const Shortcut& s : sc) {
      // one arc per (u, w) in the remaining graph: a lighter shortcut replaces what is there
      Adj* ex = nullptr;
      for (Adj& e : out[s.u])
        if (e.to == s.w) ex = &e;
      if (ex && ex->w <= s.wt) continue;
      const int32_t id = (int32_t)arcs.size();
      arcs.push_back(ArcRec{s.u, s.w, s.wt, s.arc_in, s.arc_out, arcs[s.arc_in].hops + arcs[s.arc_out].hops});
      H->shortcuts += 1;
      if (ex) {
        ex->w = s.wt;
        ex->arc = id;
        for (Adj& e : in[s.w])
          if (e.to == s.u) {
            e.w = s.wt;
            e.arc = id;
          }
      } else {
        out[s.u].push_back(Adj{s.w, id, s.wt});
        in[s.w].push_back(Adj{s.u, id, s.wt});
      }
    }
    auto touch = [&](int32_t x) {
      if (contracted[x]) return;
      level[x] = std::max(level[x], level[v] + 1);
      cn[x] += 1;
      const int64_t q = priority(x);
      if (q != cur[x]) {
        cur[x] = q;
        pq.push(PE(q, x));
      }
    };
    for (const Adj& i : in[v]) touch(i.to);
    for (const Adj& o : out[v]) touch(o.to);
    std::vector<Adj>().swap(in[v]);
    std::vector<Adj>().swap(out[v]);
  }
  // final CSR sides; children are referenced by their position inside the side they live in
  H->V = V;
  H->rank = rank;
  H->level = level;
  H->max_level = 0;
  for (int32_t v = 0; v < V; ++v) H->max_level = std::max(H->max_level, level[v]);
  std::vector<int32_t> pos(arcs.size(), -1);
  auto layout = [&](const std::vector<std::vector<int32_t>>& lists, HierArcs& side) {
    side.rowptr.assign(V + 1, 0);
    for (int32_t v = 0; v < V; ++v) side.rowptr[v + 1] = side.rowptr[v] + (int32_t)lists[v].size();
    const size_t n = (size_t)side.rowptr[V];
    side.other.resize(n);
    side.w.resize(n);
    side.a.resize(n);
    side.b.resize(n);
    side.hops.resize(n);
    for (int32_t v = 0; v < V; ++v)
      for (size_t k = 0; k < lists[v].size(); ++k) pos[lists[v][k]] = side.rowptr[v] + (int32_t)k;
  };
  layout(up_of, H->up);
  layout(down_of, H->down);
  auto fill = [&](const std::vector<std::vector<int32_t>>& lists, HierArcs& side, bool is_up) {
    for (int32_t v = 0; v < V; ++v)
      for (size_t k = 0; k < lists[v].size(); ++k) {
        const ArcRec& r = arcs[lists[v][k]];
        const size_t p = (size_t)side.rowptr[v] + k;
        side.other[p] = is_up ? r.v : r.u;
        side.w[p] = r.w;
        side.hops[p] = r.hops;
        if (r.b < 0) {
          side.a[p] = r.a;
          side.b[p] = -1;
        } else {
          side.a[p] = pos[r.a];  // (u -> mid): a down arc of mid
          side.b[p] = pos[r.b];  // (mid -> v): an up arc of mid
        }
      }
  };
  fill(up_of, H->up, true);
  fill(down_of, H->down, false);
  return true;
}

// Number of nodes a search climbing from `v` can reach on one side (= size of v's label on that side).
inline int32_t closure_size(const HierArcs& side, int32_t v, std::vector<uint32_t>& stamp, uint32_t epoch,
                            std::vector<int32_t>& stack) {
  stack.clear();
  stack.push_back(v);
  stamp[v] = epoch;
  int32_t n = 0;
  while (!stack.empty()) {
    const int32_t x = stack.back();
    stack.pop_back();
    ++n;
    for (int32_t e = side.rowptr[x]; e < side.rowptr[x + 1]; ++e) {
      const int32_t y = side.other[e];
      if (stamp[y] != epoch) {
        stamp[y] = epoch;
        stack.push_back(y);
      }
    }
  }
  return n;
}

}  // namespace drift

// Native form of the consensus cluster decisions for ONE connected component of the instance
// graph (empanada/consensus.py:35-142: create_graph_of_clusters + merge_clusters on
// graph.subgraph(comp)), called for the components whose edges do not all pass the IoU cut.
//
// Cluster ids and merge decisions of the reference depend on container iteration orders of its
// un-vendored dependencies: networkx's dict-of-dict adjacency (insertion order; Graph.copy
// re-inserts edges node by node), the BFS set of connected_components, and - through
// `graph.subgraph(comp)`'s FilterAtlas - the iteration order of a CPython `set` of node ids.
// empanada-napari_b200/consensus.py holds the same logic in Python (`_Graph`, itself checked
// against networkx); this file restates it with CPython 3.x's open-addressing set
// (Objects/setobject.c: linear probes 9, perturb shift 5, growth at 3/5 fill to 4x used) so that
// iteration orders are reproduced without the interpreter. tests/test_consensus_host.py compares
// both on hundreds of thousands of random components.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" int be_set_error(const char* msg);

namespace {

// CPython set of non-negative ints (hash(i) == i), insert-only.
struct PySet {
  std::vector<long long> key;   // -1 = empty
  size_t mask = 7, fill = 0;
  PySet() : key(8, -1) {}
  static constexpr int LINEAR_PROBES = 9;
  static void insert_clean(std::vector<long long>& table, size_t mask, long long k) {
    size_t perturb = static_cast<size_t>(k), i = static_cast<size_t>(k) & mask;
    while (true) {
      if (table[i] < 0) { table[i] = k; return; }
      if (i + LINEAR_PROBES <= mask)
        for (int j = 1; j <= LINEAR_PROBES; ++j)
          if (table[i + j] < 0) { table[i + j] = k; return; }
      perturb >>= 5;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }
  void resize(size_t minused) {
    size_t newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    std::vector<long long> nt(newsize, -1);
    for (long long k : key)
      if (k >= 0) insert_clean(nt, newsize - 1, k);
    key.swap(nt);
    mask = newsize - 1;
  }
  void add(long long k) {
    size_t perturb = static_cast<size_t>(k), i = static_cast<size_t>(k) & mask;
    size_t slot = 0;
    bool found = false;
    while (!found) {
      const int probes = (i + LINEAR_PROBES <= mask) ? LINEAR_PROBES : 0;
      for (int j = 0; j <= probes; ++j) {
        if (key[i + j] < 0) { slot = i + j; found = true; break; }
        if (key[i + j] == k) return;   // already present
      }
      if (found) break;
      perturb >>= 5;
      i = (i * 5 + 1 + perturb) & mask;
    }
    key[slot] = k;
    ++fill;
    if (fill * 5 < mask * 3) return;
    resize(fill > 50000 ? fill * 2 : fill * 4);
  }
  template <class F> void for_each(F f) const { for (long long k : key) if (k >= 0) f(static_cast<int>(k)); }
  std::vector<int> items() const { std::vector<int> v; for_each([&](int k) { v.push_back(k); }); return v; }
};

struct EdgeAttr { double iou; double overlap; };

// networkx.Graph container semantics on LOCAL node indices 0..n-1 (node order kept separately)
struct Graph {
  std::vector<int> order;                       // node insertion order (local ids)
  std::vector<char> alive;
  std::vector<std::vector<int>> adj;            // neighbour insertion order
  std::vector<std::vector<EdgeAttr>> attr;      // parallel to adj
  explicit Graph(int n) : alive(n, 0), adj(n), attr(n) {}
  void add_node(int u) { if (!alive[u]) { alive[u] = 1; order.push_back(u); } }
  int find(int u, int v) const {
    for (size_t k = 0; k < adj[u].size(); ++k) if (adj[u][k] == v) return static_cast<int>(k);
    return -1;
  }
  void add_edge(int u, int v, EdgeAttr a) {
    add_node(u); add_node(v);
    int k = find(u, v);
    if (k >= 0) { attr[u][k] = a; attr[v][find(v, u)] = a; return; }
    adj[u].push_back(v); attr[u].push_back(a);
    if (u != v) { adj[v].push_back(u); attr[v].push_back(a); }
  }
  void remove_edge(int u, int v) {
    int k = find(u, v);
    adj[u].erase(adj[u].begin() + k); attr[u].erase(attr[u].begin() + k);
    if (u != v) { k = find(v, u); adj[v].erase(adj[v].begin() + k); attr[v].erase(attr[v].begin() + k); }
  }
  void remove_node(int u) {
    for (int v : std::vector<int>(adj[u])) if (v != u) { int k = find(v, u); adj[v].erase(adj[v].begin() + k); attr[v].erase(attr[v].begin() + k); }
    adj[u].clear(); attr[u].clear();
    alive[u] = 0;
    order.erase(std::find(order.begin(), order.end(), u));
  }
  Graph copy() const {          // Graph.copy(): nodes in order, then edges node by node
    Graph g(static_cast<int>(adj.size()));
    for (int u : order) g.add_node(u);
    for (int u : order)
      for (size_t k = 0; k < adj[u].size(); ++k) g.add_edge(u, adj[u][k], attr[u][k]);
    return g;
  }
  size_t n_edges() const { size_t s = 0; for (int u : order) s += adj[u].size(); return s / 2; }
};

// _plain_bfs on a graph of local ids; `ids` maps local -> global node id (the set holds GLOBAL ids)
static PySet plain_bfs(const Graph& g, const std::vector<int>& ids, size_t n_total, int source,
                       std::vector<int>* visit_order = nullptr) {
  PySet seen;
  std::vector<char> in(g.adj.size(), 0);
  seen.add(ids[source]); in[source] = 1;
  if (visit_order) visit_order->push_back(source);
  std::vector<int> next{source};
  size_t count = 1;
  while (!next.empty()) {
    std::vector<int> level;
    level.swap(next);
    for (int v : level) {
      for (int w : g.adj[v])
        if (!in[w]) { in[w] = 1; seen.add(ids[w]); ++count; next.push_back(w); if (visit_order) visit_order->push_back(w); }
      if (count == n_total) return seen;
    }
  }
  return seen;
}

}  // namespace


namespace {
// One component: members [n] (ascending global node ids), edges [m] in the global graph's
// edge-insertion order. Appends the clusters (cluster graph node order) to sizes_out / members_out
// - clusters may share nodes: a cluster that is pushed into several larger neighbours is copied
// into each of them (consensus.py:113-119). Returns the number of clusters or a negative code.
// float_sum: 1 = Python >= 3.12 builtin sum() on floats (Neumaier compensated), 0 = plain.
int component_clusters(int n, const int* members, int m, const int* ea, const int* eb,
                       const double* eiou, const long long* eov, long long n_nodes_total,
                       double cluster_iou_thr, double min_iou, double min_overlap, int float_sum,
                       std::vector<int>& sizes_out, std::vector<int>& members_out) {
  if (n <= 0) return 0;
  std::vector<int> ids(members, members + n);
  auto local = [&](int gid) { return static_cast<int>(std::lower_bound(ids.begin(), ids.end(), gid) - ids.begin()); };
  // adjacency of the full graph restricted to the component (edge-insertion order), for the BFS set
  Graph g0(n);
  for (int i = 0; i < n; ++i) g0.add_node(i);
  for (int k = 0; k < m; ++k) g0.add_edge(local(ea[k]), local(eb[k]), EdgeAttr{eiou[k], static_cast<double>(eov[k])});
  // node order of graph.subgraph(comp): the set rebuilt from the BFS set when it is "shorter"
  std::vector<int> node_seq;
  if (2LL * n < n_nodes_total) {
    const PySet bfs = plain_bfs(g0, ids, static_cast<size_t>(n_nodes_total), 0);
    PySet again;
    bfs.for_each([&](int gid) { again.add(gid); });
    again.for_each([&](int gid) { node_seq.push_back(local(gid)); });
  } else {
    for (int i = 0; i < n; ++i) node_seq.push_back(i);
  }
  if (static_cast<int>(node_seq.size()) != n) return be_set_error("component is not connected");
  Graph sub(n);
  for (int u : node_seq) sub.add_node(u);
  for (int k = 0; k < m; ++k) sub.add_edge(local(ea[k]), local(eb[k]), EdgeAttr{eiou[k], static_cast<double>(eov[k])});

  // ---- create_graph_of_clusters
  Graph H = sub.copy();
  {
    std::vector<char> seen(n, 0);     // G.edges(): (u, v) with v not yet finished
    std::vector<std::pair<int, int>> cut;
    for (int u : sub.order) {
      for (size_t k = 0; k < sub.adj[u].size(); ++k) {
        const int v = sub.adj[u][k];
        if (!seen[v] && sub.attr[u][k].iou <= cluster_iou_thr) cut.emplace_back(u, v);
      }
      seen[u] = 1;
    }
    for (auto& e : cut) H.remove_edge(e.first, e.second);
  }
  std::vector<std::vector<int>> cluster_iter;   // members of every initial cluster in SET iteration order (local ids)
  std::vector<int> owner(n, -1);
  {
    std::vector<char> done(n, 0);
    const size_t hn = H.order.size();
    for (int v : H.order) {
      if (done[v]) continue;
      std::vector<int> visited;
      const PySet c = plain_bfs(H, ids, hn, v, &visited);
      for (int w : visited) done[w] = 1;
      std::vector<int> it;
      c.for_each([&](int gid) { it.push_back(local(gid)); });
      for (int w : it) owner[w] = static_cast<int>(cluster_iter.size());
      cluster_iter.push_back(it);
    }
  }
  const int nc = static_cast<int>(cluster_iter.size());
  std::vector<std::pair<int, int>> linked;
  {
    std::vector<char> seen(n, 0);
    for (int u : sub.order) {
      for (int v : sub.adj[u])
        if (!seen[v]) {
          const int a = owner[u], b = owner[v];
          if (a != b) linked.emplace_back(std::min(a, b), std::max(a, b));
        }
      seen[u] = 1;
    }
    std::sort(linked.begin(), linked.end());
    linked.erase(std::unique(linked.begin(), linked.end()), linked.end());
  }
  Graph CG(nc);
  for (int i = 0; i < nc; ++i) CG.add_node(i);
  for (auto& pr : linked) {
    const std::vector<int>&c1 = cluster_iter[pr.first], &c2 = cluster_iter[pr.second];
    // Python: sum(list of floats and int zeros) / len(list). The ints are zeros (no effect); the
    // floats are added left to right, with Neumaier's compensation from Python 3.12 on
    // (Python/bltinmodule.c); the overlaps are integers (exact).
    double si = 0.0, comp = 0.0, so = 0.0;
    bool first = true;
    for (int a : c1)
      for (int b : c2) {
        const int k = sub.find(a, b);
        if (k < 0) continue;
        const double x = sub.attr[a][k].iou;
        if (first || !float_sum) { si += x; first = false; }
        else {
          const double t = si + x;
          if (std::fabs(si) >= std::fabs(x)) comp += (si - t) + x; else comp += (x - t) + si;
          si = t;
        }
        so += sub.attr[a][k].overlap;
      }
    if (float_sum && comp != 0.0 && std::isfinite(comp)) si += comp;
    const double cnt = static_cast<double>(c1.size() * c2.size());
    const double iw = si / cnt, ow = so / cnt;
    if (iw > min_iou || ow > min_overlap) CG.add_edge(pr.first, pr.second, EdgeAttr{iw, ow});
  }

  // ---- merge_clusters
  Graph M = CG.copy();
  std::vector<std::vector<int>> cl(cluster_iter);    // membership only (order irrelevant from here on)
  auto unite_into = [&](int dst, int src) {
    for (int w : cl[src]) if (std::find(cl[dst].begin(), cl[dst].end(), w) == cl[dst].end()) cl[dst].push_back(w);
  };
  while (M.n_edges() > 0) {
    int mc = -1;
    size_t best = 0;
    for (int u : M.order) if (mc < 0 || M.adj[u].size() > best) { mc = u; best = M.adj[u].size(); }
    std::vector<int> nbrs(M.adj[mc]);
    std::stable_sort(nbrs.begin(), nbrs.end(), [&](int a, int b) { return cl[a].size() > cl[b].size(); });
    if (cl[nbrs[0]].size() > cl[mc].size()) {
      for (int nb : nbrs) { unite_into(nb, mc); M.remove_edge(mc, nb); }
      M.remove_node(mc);
    } else {
      for (int nb : nbrs) {
        unite_into(mc, nb);
        M.remove_edge(nb, mc);
        // (the reference re-adds the edge mc-nb for neighbours of nb that mc does not touch and
        // then removes nb: no lasting effect)
        M.remove_node(nb);
      }
    }
  }
  int out = 0;
  for (int u : M.order) {
    sizes_out.push_back(static_cast<int>(cl[u].size()));
    for (int w : cl[u]) members_out.push_back(ids[w]);
    ++out;
  }
  return out;
}

thread_local std::vector<int> g_sizes, g_members;
}  // namespace

extern "C" {

// n_comp components in CSR layout (node_off / edge_off [n_comp + 1]). n_clusters [n_comp] out;
// totals[0] = clusters, totals[1] = cluster members over all components; the lists themselves are
// kept in a thread-local buffer and copied out by be_components_clusters_fetch (cluster_sizes
// [totals[0]], members [totals[1]], back to back in component / cluster order).
int be_components_clusters(int n_comp, const int* node_off, const int* nodes, const int* edge_off,
                           const int* ea, const int* eb, const double* eiou, const long long* eov,
                           long long n_nodes_total, double cluster_iou_thr, double min_iou,
                           double min_overlap, int float_sum, int* n_clusters, long long* totals) {
  g_sizes.clear();
  g_members.clear();
  for (int c = 0; c < n_comp; ++c) {
    const int n0 = node_off[c], n = node_off[c + 1] - n0, e0 = edge_off[c], m = edge_off[c + 1] - e0;
    const int k = component_clusters(n, nodes + n0, m, ea + e0, eb + e0, eiou + e0, eov + e0, n_nodes_total,
                                     cluster_iou_thr, min_iou, min_overlap, float_sum, g_sizes, g_members);
    if (k < 0) return k;
    n_clusters[c] = k;
  }
  totals[0] = static_cast<long long>(g_sizes.size());
  totals[1] = static_cast<long long>(g_members.size());
  return 0;
}

int be_components_clusters_fetch(int* cluster_sizes, int* members) {
  if (!g_sizes.empty()) std::memcpy(cluster_sizes, g_sizes.data(), g_sizes.size() * sizeof(int));
  if (!g_members.empty()) std::memcpy(members, g_members.data(), g_members.size() * sizeof(int));
  return 0;
}

}  // extern "C"

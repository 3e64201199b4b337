// Host replay of the reference's sequential instance tracker on the sparse tables the GPU
// kernels produce (per-slice component areas/boxes + adjacent-slice overlap counts).
// Decision logic restated from:
//   RLEMatcher.__call__ / rle_matcher        empanada/inference/matcher.py:136-326
//   forward_matching / backward_matching     empanada/inference/patterns.py:55-121
//   InstanceTracker.update (boxes, order)    empanada/inference/tracker.py:11-100
// Geometry never enters here: IoU/IoA of merged objects are sums over their (disjoint)
// components, so the overlap table is sufficient. The Hungarian step is SciPy's
// linear_sum_assignment(maximize=True) (un-vendored dependency of the reference,
// matcher.py:213): the same shortest-augmenting-path algorithm (Crouse 2016) is run here on
// each connected block of the sparse IoU matrix, rows/columns kept in matrix order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

extern "C" int be_set_error(const char* msg);

namespace {

// Objects of one slice in CSR form (no per-object heap allocation).
struct SliceObjs {
  std::vector<int> label;
  std::vector<long long> area;
  std::vector<int> box;     // 4 per object: y0 x0 y1 x1 (half open)
  std::vector<int> cc_off;  // size n + 1
  std::vector<int> cc;      // component ids (1-based) of this slice, grouped by object
  int size() const { return static_cast<int>(label.size()); }
  void clear() { label.clear(); area.clear(); box.clear(); cc_off.assign(1, 0); cc.clear(); }
};

// Rectangular LAP, minimisation, shortest augmenting paths (Crouse, "On implementing 2D
// rectangular assignment algorithms", 2016) - the algorithm behind scipy.optimize.linear_sum_assignment.
// cost is nr x nc row-major with nr <= nc. Returns col4row.
// The duals stay in the workspace: cost[i][j] - u[i] - v[j] >= 0, = 0 on the assignment.
struct LapWork {
  std::vector<double> u, v, shortest;
  std::vector<int> path, row4col, remaining;
  std::vector<char> SR, SC;
};

static void lap_min(int nr, int nc, const std::vector<double>& cost, std::vector<int>& col4row, LapWork& W) {
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double>&u = W.u, &v = W.v, &shortest = W.shortest;
  std::vector<int>&path = W.path, &row4col = W.row4col, &remaining = W.remaining;
  std::vector<char>&SR = W.SR, &SC = W.SC;
  u.assign(nr, 0.0); v.assign(nc, 0.0); shortest.resize(nc);
  path.assign(nc, -1); row4col.assign(nc, -1); remaining.resize(nc);
  SR.resize(nr); SC.resize(nc);
  col4row.assign(nr, -1);
  for (int cur = 0; cur < nr; ++cur) {
    double min_val = 0.0;
    int i = cur;
    int num_remaining = nc;
    for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
    std::fill(SR.begin(), SR.end(), 0);
    std::fill(SC.begin(), SC.end(), 0);
    std::fill(shortest.begin(), shortest.end(), INF);
    int sink = -1;
    while (sink == -1) {
      int index = -1;
      double lowest = INF;
      SR[i] = 1;
      for (int it = 0; it < num_remaining; ++it) {
        const int j = remaining[it];
        const double r = min_val + cost[static_cast<size_t>(i) * nc + j] - u[i] - v[j];
        if (r < shortest[j]) { path[j] = i; shortest[j] = r; }
        if (shortest[j] < lowest || (shortest[j] == lowest && row4col[j] == -1)) {
          lowest = shortest[j];
          index = it;
        }
      }
      min_val = lowest;
      if (min_val == INF) return;  // infeasible (cannot happen for finite costs)
      const int j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = 1;
      remaining[index] = remaining[--num_remaining];
    }
    u[cur] += min_val;
    for (int r = 0; r < nr; ++r)
      if (SR[r] && r != cur) u[r] += min_val - shortest[col4row[r]];
    for (int j = 0; j < nc; ++j)
      if (SC[j]) v[j] -= min_val - shortest[j];
    int j = sink;
    while (true) {
      const int r = path[j];
      row4col[j] = r;
      std::swap(col4row[r], j);
      if (r == cur) break;
    }
  }
}

struct Entry { int row, col; long long inter; };
constexpr int kNoLabel = std::numeric_limits<int>::min();
constexpr double kTieEps = 1e-12;  // reduced costs below this count as ties (IoUs are O(1))
// diagnostics of the last be_match_replay call: matcher steps, steps with a multi-entry block,
// steps replayed on the full matrix
static thread_local long long g_stats[3] = {0, 0, 0};   // per calling thread: replays of two planes may overlap

struct Scratch {
  std::vector<Entry> agg, bycol;
  std::vector<double> iou;
  std::vector<float> ioa;
  std::vector<int> parent, matched_row, ioa_arg, rows, cols, col_start, row_deg, col_deg, multi, root;
  std::vector<float> ioa_max;
  std::vector<double> dense, cost;
  std::vector<int> col4row;
  LapWork lap;
  // merge_by_label
  std::vector<int> gid, gcount, gstart, members, hkey, hval;
  // star blocks
  std::vector<char> row_star, col_star;
  std::vector<int> row_best, col_best;
  std::vector<double> row_second, col_second;
  // uniqueness test of a block optimum
  std::vector<char> alt_adj, alt_free, alt_zero, alt_seen;
  std::vector<int> alt_stack;
};

// Does the block have a second optimal assignment? By complementary slackness every optimal
// assignment uses only TIGHT pairs (reduced cost cost[a][j] - u[a] - v[j] = 0) and covers every
// column with v[j] < 0. Against the assignment found, another one differs by alternating cycles
// (row a takes the column of row b, ... back to a) or by an alternating path that ends in a free
// column and frees a column whose dual is zero. Rows are nodes; a -> b when (a, column of b) is
// tight. Conservative under rounding: anything within kTieEps counts as tight / zero.
static bool block_has_alternative(int nr, int nc, const std::vector<double>& cost, const std::vector<int>& col4row,
                                  const LapWork& W, Scratch& S) {
  S.alt_adj.assign(static_cast<size_t>(nr) * nr, 0);
  S.alt_free.assign(nr, 0);
  S.alt_zero.assign(nr, 0);
  bool any_edge = false;
  for (int a = 0; a < nr; ++a) {
    S.alt_zero[a] = W.v[col4row[a]] >= -kTieEps;
    for (int j = 0; j < nc; ++j) {
      if (j == col4row[a] || cost[static_cast<size_t>(a) * nc + j] - W.u[a] - W.v[j] > kTieEps) continue;
      const int b = W.row4col[j];
      if (b < 0) { S.alt_free[a] = 1; if (S.alt_zero[a]) return true; }
      else { S.alt_adj[static_cast<size_t>(a) * nr + b] = 1; any_edge = true; }
    }
  }
  if (!any_edge) return false;
  for (int start = 0; start < nr; ++start) {     // blocks are small: one search per row
    S.alt_seen.assign(nr, 0);
    S.alt_stack.clear();
    S.alt_stack.push_back(start);
    while (!S.alt_stack.empty()) {
      const int a = S.alt_stack.back();
      S.alt_stack.pop_back();
      for (int b = 0; b < nr; ++b) {
        if (!S.alt_adj[static_cast<size_t>(a) * nr + b]) continue;
        if (b == start) return true;                              // alternating cycle
        if (S.alt_seen[b]) continue;
        S.alt_seen[b] = 1;
        if (S.alt_zero[start] && S.alt_free[b]) return true;       // alternating path
        S.alt_stack.push_back(b);
      }
    }
  }
  return false;
}

// One matcher step: relabel `match` objects against `target` objects.
// entries: sparse intersections (row = target index, col = match index), duplicates allowed.
// Everything here is linear in the number of entries for the common case (an object that
// overlaps exactly one object of the neighbouring slice, and vice versa): entries are bucketed
// by column, duplicates summed, and only the entries that share a row or a column with another
// one go through the union-find / assignment machinery - an isolated positive entry is a
// connected block of the IoU matrix by itself and always part of the optimum.
static void match_step(const SliceObjs& target, const SliceObjs& match, std::vector<Entry>& entries,
                       double iou_thr, float ioa_thr, bool assign_new, int& next_label,
                       std::vector<int>& new_labels, Scratch& S) {
  const int n = target.size(), m = match.size();
  ++g_stats[0];
  new_labels.assign(m, 0);
  S.matched_row.assign(m, -1);
  S.ioa_max.assign(m, 0.0f);
  S.ioa_arg.assign(m, 0);
  std::vector<Entry>& agg = S.agg;
  agg.clear();
  if (n > 0 && m > 0 && !entries.empty()) {
    // stable counting sort by column, rows ascending inside a column, duplicates summed
    std::vector<int>& cs = S.col_start;
    cs.assign(m + 1, 0);
    for (const Entry& e : entries) ++cs[e.col + 1];
    for (int c = 0; c < m; ++c) cs[c + 1] += cs[c];
    S.bycol.resize(entries.size());
    S.col_deg.assign(m, 0);  // used as the fill cursor first
    for (const Entry& e : entries) S.bycol[cs[e.col] + S.col_deg[e.col]++] = e;
    S.row_deg.assign(n, 0);
    bool shared = false;   // does any row or column hold more than one entry?
    for (int c = 0; c < m; ++c) {
      Entry* b = S.bycol.data() + cs[c];
      const int len = cs[c + 1] - cs[c];
      if (len > 16) {
        std::sort(b, b + len, [](const Entry& x, const Entry& y) { return x.row < y.row; });
      } else {
        for (int i = 1; i < len; ++i) {
          const Entry e = b[i];
          int j = i - 1;
          while (j >= 0 && b[j].row > e.row) { b[j + 1] = b[j]; --j; }
          b[j + 1] = e;
        }
      }
      int deg = 0;
      for (int i = 0; i < len; ++i) {
        if (deg > 0 && agg.back().row == b[i].row) agg.back().inter += b[i].inter;
        else { agg.push_back(b[i]); ++deg; if (++S.row_deg[b[i].row] > 1) shared = true; }
      }
      S.col_deg[c] = deg;
      if (deg > 1) shared = true;
    }
    const size_t na = agg.size();
    S.iou.resize(na);
    S.ioa.resize(na);
    S.multi.clear();
    // Star blocks - one row against several columns that touch nothing else, or one column
    // against several such rows (an object that splits in two, or two that merge) - are the
    // bulk of the non-isolated entries: their assignment is the largest IoU of the star, unique
    // exactly when the runner-up is smaller (the same dual certificate as for general blocks:
    // u = -best, v = 0, reduced cost of the others = best - iou).
    if (shared) {
      S.row_star.assign(n, 1);
      S.col_star.assign(m, 1);
      for (size_t k = 0; k < na; ++k) {
        const Entry& e = agg[k];
        if (S.col_deg[e.col] != 1) S.row_star[e.row] = 0;
        if (S.row_deg[e.row] != 1) S.col_star[e.col] = 0;
      }
      S.row_best.assign(n, -1);
      S.col_best.assign(m, -1);
      S.row_second.assign(n, -1.0);
      S.col_second.assign(m, -1.0);
    }
    bool tie_risk = false;
    for (size_t k = 0; k < na; ++k) {
      const Entry& e = agg[k];
      const long long inter = e.inter;
      const long long uni = target.area[e.row] + match.area[e.col] - inter;
      S.iou[k] = static_cast<double>(inter) / static_cast<double>(uni);
      S.ioa[k] = static_cast<float>(static_cast<double>(inter) / static_cast<double>(match.area[e.col]));
      // per-column IoA maximum (float32 matrix semantics: zeros everywhere else, first max row
      // wins - rows ascend inside a column, so ">" keeps the first one)
      if (S.ioa[k] > S.ioa_max[e.col]) { S.ioa_max[e.col] = S.ioa[k]; S.ioa_arg[e.col] = e.row; }
      const int rd = S.row_deg[e.row], cd = S.col_deg[e.col];
      if (rd == 1 && cd == 1) {
        if (S.iou[k] >= iou_thr) S.matched_row[e.col] = e.row;
      } else if (cd > 1 && S.col_star[e.col]) {       // rd == 1 for every entry of this column
        int& best = S.col_best[e.col];
        double& second = S.col_second[e.col];
        if (best < 0 || S.iou[k] > S.iou[best]) { if (best >= 0) second = std::max(second, S.iou[best]); best = static_cast<int>(k); }
        else second = std::max(second, S.iou[k]);
      } else if (rd > 1 && S.row_star[e.row]) {       // cd == 1 for every entry of this row
        int& best = S.row_best[e.row];
        double& second = S.row_second[e.row];
        if (best < 0 || S.iou[k] > S.iou[best]) { if (best >= 0) second = std::max(second, S.iou[best]); best = static_cast<int>(k); }
        else second = std::max(second, S.iou[k]);
      } else {
        S.multi.push_back(static_cast<int>(k));
      }
    }
    for (int c = 0; shared && c < m; ++c) {
      const int best = S.col_best[c];
      if (best < 0) continue;
      if (S.iou[best] - S.col_second[c] <= kTieEps) tie_risk = true;
      else if (S.iou[best] >= iou_thr) S.matched_row[c] = agg[best].row;
    }
    for (int r = 0; shared && r < n; ++r) {
      const int best = S.row_best[r];
      if (best < 0) continue;
      if (S.iou[best] - S.row_second[r] <= kTieEps) tie_risk = true;
      else if (S.iou[best] >= iou_thr) S.matched_row[agg[best].col] = r;
    }
    if (!S.multi.empty()) {
      ++g_stats[1];
      // connected blocks of the bipartite graph (union-find over rows [0,n) and cols [n,n+m))
      std::vector<int>& parent = S.parent;
      parent.resize(n + m);
      for (int k : S.multi) { parent[agg[k].row] = agg[k].row; parent[n + agg[k].col] = n + agg[k].col; }
      auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
      for (int k : S.multi) {
        const int a = find(agg[k].row), b = find(n + agg[k].col);
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
      }
      S.root.resize(na);
      for (int k : S.multi) S.root[k] = find(agg[k].row);
      std::sort(S.multi.begin(), S.multi.end(), [&](int a, int b) {
        return S.root[a] != S.root[b] ? S.root[a] < S.root[b] : a < b;
      });
      size_t g0 = 0;
      const size_t ne = S.multi.size();
      while (g0 < ne) {
        const int root = S.root[S.multi[g0]];
        size_t g1 = g0 + 1;
        while (g1 < ne && S.root[S.multi[g1]] == root) ++g1;
        std::vector<int>&rows = S.rows, &cols = S.cols;
        rows.clear(); cols.clear();
        for (size_t k = g0; k < g1; ++k) { rows.push_back(agg[S.multi[k]].row); cols.push_back(agg[S.multi[k]].col); }
        std::sort(rows.begin(), rows.end()); rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
        std::sort(cols.begin(), cols.end()); cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        const int br = static_cast<int>(rows.size()), bc = static_cast<int>(cols.size());
        S.dense.assign(static_cast<size_t>(br) * bc, 0.0);
        for (size_t k = g0; k < g1; ++k) {
          const Entry& e = agg[S.multi[k]];
          const int r = static_cast<int>(std::lower_bound(rows.begin(), rows.end(), e.row) - rows.begin());
          const int c = static_cast<int>(std::lower_bound(cols.begin(), cols.end(), e.col) - cols.begin());
          S.dense[static_cast<size_t>(r) * bc + c] = S.iou[S.multi[k]];
        }
        // scipy: maximise == minimise(-cost); tall matrices are transposed first
        const bool transpose = bc < br;
        const int nr = transpose ? bc : br, nc = transpose ? br : bc;
        S.cost.resize(static_cast<size_t>(nr) * nc);
        for (int r = 0; r < br; ++r)
          for (int c = 0; c < bc; ++c) {
            const double val = -S.dense[static_cast<size_t>(r) * bc + c];
            if (transpose) S.cost[static_cast<size_t>(c) * nc + r] = val; else S.cost[static_cast<size_t>(r) * nc + c] = val;
          }
        lap_min(nr, nc, S.cost, S.col4row, S.lap);
        for (int a = 0; a < nr; ++a) {
          if (S.col4row[a] < 0) continue;
          const int r = transpose ? S.col4row[a] : a, c = transpose ? a : S.col4row[a];
          if (S.dense[static_cast<size_t>(r) * bc + c] >= iou_thr) S.matched_row[cols[c]] = rows[r];
        }
        // Is this optimum the only one? Cross-block entries being zero, SciPy's run on the full
        // matrix must pick the same positive pairs when it is. Otherwise the choice among equal
        // optima depends on SciPy's scan order over the WHOLE matrix: that run is replayed
        // literally (below).
        if (!tie_risk && block_has_alternative(nr, nc, S.cost, S.col4row, S.lap, S)) tie_risk = true;
        g0 = g1;
      }
    }
    if (tie_risk) {
      ++g_stats[2];
      // scipy.optimize.linear_sum_assignment(iou, maximize=True) on the full n x m matrix
      // (matcher.py:213): negate, transpose when there are fewer columns than rows
      const bool transpose = m < n;
      const int nr = transpose ? m : n, nc = transpose ? n : m;
      S.cost.assign(static_cast<size_t>(nr) * nc, 0.0);
      for (size_t k = 0; k < na; ++k) {
        const Entry& e = agg[k];
        S.cost[transpose ? static_cast<size_t>(e.col) * nc + e.row : static_cast<size_t>(e.row) * nc + e.col] = -S.iou[k];
      }
      lap_min(nr, nc, S.cost, S.col4row, S.lap);
      S.matched_row.assign(m, -1);
      for (int a = 0; a < nr; ++a) {
        if (S.col4row[a] < 0) continue;
        const int r = transpose ? S.col4row[a] : a, c = transpose ? a : S.col4row[a];
        if (-S.cost[static_cast<size_t>(a) * nc + S.col4row[a]] >= iou_thr) S.matched_row[c] = r;
      }
    }
  }
  for (int i = 0; i < m; ++i) {
    if (S.matched_row[i] >= 0) {
      new_labels[i] = target.label[S.matched_row[i]];
    } else if (n > 0 && m > 0 && S.ioa_max[i] >= ioa_thr) {
      new_labels[i] = target.label[S.ioa_arg[i]];
    } else if (assign_new) {
      new_labels[i] = next_label++;
    } else {
      new_labels[i] = match.label[i];
    }
  }
}

// Objects with equal new label are merged; output order = first appearance (dict order), members
// of a merged object in index order. Labels are grouped through a small open-addressing table;
// when every label is distinct (the common case) the slice is copied with its new labels.
static void merge_by_label(const SliceObjs& match, const std::vector<int>& new_labels, SliceObjs& out,
                           Scratch& S) {
  const int m = match.size();
  int cap = 16;
  while (cap < 2 * m) cap <<= 1;
  S.hkey.assign(cap, kNoLabel);
  S.hval.resize(cap);
  S.gid.resize(m);
  S.gcount.clear();
  int ng = 0;
  for (int i = 0; i < m; ++i) {
    const int lab = new_labels[i];
    unsigned h = (static_cast<unsigned>(lab) * 2654435761u) & (cap - 1);
    while (S.hkey[h] != kNoLabel && S.hkey[h] != lab) h = (h + 1) & (cap - 1);
    if (S.hkey[h] == kNoLabel) { S.hkey[h] = lab; S.hval[h] = ng++; S.gcount.push_back(0); }
    S.gid[i] = S.hval[h];
    ++S.gcount[S.gid[i]];
  }
  if (ng == m) {  // nothing merges
    out.label = new_labels;
    out.area = match.area;
    out.box = match.box;
    out.cc_off = match.cc_off;
    out.cc = match.cc;
    return;
  }
  S.gstart.assign(ng + 1, 0);
  for (int g = 0; g < ng; ++g) S.gstart[g + 1] = S.gstart[g] + S.gcount[g];
  S.members.resize(m);
  std::fill(S.gcount.begin(), S.gcount.end(), 0);
  for (int i = 0; i < m; ++i) S.members[S.gstart[S.gid[i]] + S.gcount[S.gid[i]]++] = i;
  out.label.resize(ng); out.area.resize(ng); out.box.resize(4 * ng); out.cc_off.resize(ng + 1);
  out.cc.resize(match.cc.size());
  out.cc_off[0] = 0;
  int w = 0;
  for (int g = 0; g < ng; ++g) {
    long long area = 0;
    int b0 = 0x7fffffff, b1 = 0x7fffffff, b2 = -1, b3 = -1;
    for (int k = S.gstart[g]; k < S.gstart[g + 1]; ++k) {
      const int i = S.members[k];
      area += match.area[i];
      b0 = std::min(b0, match.box[4 * i]); b1 = std::min(b1, match.box[4 * i + 1]);
      b2 = std::max(b2, match.box[4 * i + 2]); b3 = std::max(b3, match.box[4 * i + 3]);
      for (int c = match.cc_off[i]; c < match.cc_off[i + 1]; ++c) out.cc[w++] = match.cc[c];
    }
    out.label[g] = new_labels[S.members[S.gstart[g]]];
    out.area[g] = area;
    out.box[4 * g] = b0; out.box[4 * g + 1] = b1; out.box[4 * g + 2] = b2; out.box[4 * g + 3] = b3;
    out.cc_off[g + 1] = w;
  }
}

}  // namespace

extern "C" {

// out[0..2]: matcher steps / steps with a multi-entry block / steps replayed on the full matrix,
// of the last be_match_replay call of the calling thread (diagnostics)
int be_match_replay_stats(long long* out) {
  for (int i = 0; i < 3; ++i) out[i] = g_stats[i];
  return 0;
}

// n_cc        [n_slices]              components per slice
// cc_table    [n_slices][cap][5]      area, y0, x0, y1, x1
// pair_keys   slice(24)|prev_cc(20)|cur_cc(20), pair_vals = overlapping pixels (slice vs slice-1)
// axis        0 xy, 1 xz, 2 yz        (tracker.py:11-23 box lifting)
// lut         [n_slices][lut_stride]  out: component id -> final tracked label
// inst_*      out, tracker insertion order (first arrival in the backward sweep)
int be_match_replay(int n_slices, const int* n_cc, const int* cc_table, int cap,
                    const unsigned long long* pair_keys, const int* pair_vals, long long n_pairs,
                    int class_id, int label_divisor, double iou_thr, double ioa_thr, int axis,
                    int* lut, int lut_stride, int* inst_labels, long long* inst_sizes,
                    int* inst_boxes, int max_inst, int* n_inst) {
  g_stats[0] = g_stats[1] = g_stats[2] = 0;
  if (n_slices <= 0) { *n_inst = 0; return 0; }
  // With a threshold <= 0 the reference also keeps the ZERO-overlap pairs SciPy happens to assign;
  // the sparse formulation has no such pairs (the engines fix both thresholds at 0.25).
  if (!(iou_thr > 0.0)) return be_set_error("be_match_replay: iou_thr must be positive");
  for (int s = 0; s < n_slices; ++s)
    if (n_cc[s] > cap || n_cc[s] >= lut_stride || n_cc[s] >= (1 << 20))
      return be_set_error("component count exceeds table capacity; re-run with a larger cap");
  // bucket pairs by slice
  std::vector<long long> pstart(n_slices + 1, 0);
  for (long long k = 0; k < n_pairs; ++k) {
    const int s = static_cast<int>(pair_keys[k] >> 40);
    if (s < 0 || s >= n_slices) return be_set_error("overlap table: slice index out of range");
    ++pstart[s + 1];
  }
  for (int s = 0; s < n_slices; ++s) pstart[s + 1] += pstart[s];
  std::vector<long long> porder(n_pairs), fill(pstart.begin(), pstart.end() - 1);
  for (long long k = 0; k < n_pairs; ++k) porder[fill[pair_keys[k] >> 40]++] = k;

  const int base_label = class_id * label_divisor;
  const float ioa_thr_f = static_cast<float>(ioa_thr);
  Scratch S;
  auto slice_ccs = [&](int s, SliceObjs& out) {
    out.clear();
    const int* t = cc_table + static_cast<size_t>(s) * cap * 5;
    const int n = n_cc[s];
    out.label.resize(n); out.area.resize(n); out.box.resize(4 * n); out.cc.resize(n); out.cc_off.resize(n + 1);
    for (int c = 0; c < n; ++c) {
      out.label[c] = base_label + c + 1;
      out.area[c] = t[c * 5];
      out.box[4 * c] = t[c * 5 + 1]; out.box[4 * c + 1] = t[c * 5 + 2];
      out.box[4 * c + 2] = t[c * 5 + 3]; out.box[4 * c + 3] = t[c * 5 + 4];
      out.cc[c] = c + 1;
      out.cc_off[c + 1] = c + 1;
    }
  };
  auto fill_owner = [](const SliceObjs& o, int ncc, std::vector<int>& owner) {
    owner.assign(ncc + 1, -1);
    for (int i = 0; i < o.size(); ++i)
      for (int k = o.cc_off[i]; k < o.cc_off[i + 1]; ++k) owner[o.cc[k]] = i;
  };

  // ---------------- forward pass (assign_new = True)
  std::vector<SliceObjs> fwd(n_slices);
  int next_label = base_label + 1;
  std::vector<int> owner;  // component id (slice s-1) -> index into fwd[s-1]
  std::vector<Entry> entries;
  std::vector<int> new_labels;
  SliceObjs cur;
  for (int s = 0; s < n_slices; ++s) {
    slice_ccs(s, cur);
    if (s == 0) {
      fwd[0] = cur;
      if (cur.size() > 0) next_label = cur.label.back() + 1;  // max(labels) + 1, labels ascend
    } else {
      const SliceObjs& tgt = fwd[s - 1];
      fill_owner(tgt, n_cc[s - 1], owner);
      entries.clear();
      for (long long k = pstart[s]; k < pstart[s + 1]; ++k) {
        const unsigned long long key = pair_keys[porder[k]];
        const int q = static_cast<int>((key >> 20) & 0xFFFFF), c = static_cast<int>(key & 0xFFFFF);
        if (q < 1 || q > n_cc[s - 1] || c < 1 || c > n_cc[s]) return be_set_error("overlap table: component id out of range");
        entries.push_back({owner[q], c - 1, pair_vals[porder[k]]});
      }
      match_step(tgt, cur, entries, iou_thr, ioa_thr_f, true, next_label, new_labels, S);
      merge_by_label(cur, new_labels, fwd[s], S);
    }
  }

  // ---------------- backward pass (assign_new = False) + tracker
  // every label of the backward sweep was issued by the forward pass: base < label < next_label
  std::vector<int> inst_pos(static_cast<size_t>(std::max(1, next_label - base_label)), -1);
  int ninst = 0;
  SliceObjs bwd_next, bwd_cur;
  std::vector<int> owner_next, owner_cur;
  int dummy_next = 0;
  for (int s = n_slices - 1; s >= 0; --s) {
    if (s == n_slices - 1) {
      bwd_cur = fwd[s];
    } else {
      const SliceObjs& match = fwd[s];
      fill_owner(match, n_cc[s], owner_cur);
      fill_owner(bwd_next, n_cc[s + 1], owner_next);
      entries.clear();
      for (long long k = pstart[s + 1]; k < pstart[s + 2]; ++k) {
        const unsigned long long key = pair_keys[porder[k]];
        const int q = static_cast<int>((key >> 20) & 0xFFFFF), c = static_cast<int>(key & 0xFFFFF);
        entries.push_back({owner_next[c], owner_cur[q], pair_vals[porder[k]]});
      }
      match_step(bwd_next, match, entries, iou_thr, ioa_thr_f, false, dummy_next, new_labels, S);
      merge_by_label(match, new_labels, bwd_cur, S);
    }
    // tracker.update(rle_seg, s): dict order of bwd_cur
    int* l = lut + static_cast<size_t>(s) * lut_stride;
    std::memset(l, 0, sizeof(int) * lut_stride);
    for (int i = 0; i < bwd_cur.size(); ++i) {
      const int label = bwd_cur.label[i];
      const int* ob = &bwd_cur.box[4 * i];
      for (int k = bwd_cur.cc_off[i]; k < bwd_cur.cc_off[i + 1]; ++k) l[bwd_cur.cc[k]] = label;
      int b3[6];
      if (axis == 0) { b3[0] = s; b3[1] = ob[0]; b3[2] = ob[1]; b3[3] = s + 1; b3[4] = ob[2]; b3[5] = ob[3]; }
      else if (axis == 1) { b3[0] = ob[0]; b3[1] = s; b3[2] = ob[1]; b3[3] = ob[2]; b3[4] = s + 1; b3[5] = ob[3]; }
      else { b3[0] = ob[0]; b3[1] = ob[1]; b3[2] = s; b3[3] = ob[2]; b3[4] = ob[3]; b3[5] = s + 1; }
      if (label <= base_label || label >= next_label) return be_set_error("tracker label outside the issued range");
      int& pos = inst_pos[label - base_label];
      if (pos < 0) {
        if (ninst >= max_inst) return be_set_error("instance table capacity exceeded; re-run with a larger max_inst");
        pos = ninst;
        inst_labels[ninst] = label;
        inst_sizes[ninst] = bwd_cur.area[i];
        std::memcpy(inst_boxes + ninst * 6, b3, sizeof(b3));
        ++ninst;
      } else {
        const int p = pos;
        inst_sizes[p] += bwd_cur.area[i];
        int* bb = inst_boxes + p * 6;
        for (int d = 0; d < 3; ++d) { bb[d] = std::min(bb[d], b3[d]); bb[d + 3] = std::max(bb[d + 3], b3[d + 3]); }
      }
    }
    std::swap(bwd_next, bwd_cur);
  }
  *n_inst = ninst;
  return 0;
}

}  // extern "C"

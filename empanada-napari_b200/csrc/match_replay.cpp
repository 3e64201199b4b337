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
#include <unordered_map>
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
static void lap_min(int nr, int nc, const std::vector<double>& cost, std::vector<int>& col4row) {
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(nr, 0.0), v(nc, 0.0), shortest(nc);
  std::vector<int> path(nc, -1), row4col(nc, -1), remaining(nc);
  std::vector<char> SR(nr), SC(nc);
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

struct Scratch {
  std::vector<Entry> agg;
  std::vector<double> iou;
  std::vector<float> ioa;
  std::vector<int> parent, matched_row, ioa_arg, order, rows, cols;
  std::vector<float> ioa_max;
  std::vector<size_t> eorder;
  std::vector<double> dense, cost;
  std::vector<int> col4row;
  std::vector<std::pair<int, int>> groups;  // (first index, start in order)
};

// One matcher step: relabel `match` objects against `target` objects.
// entries: sparse intersections (row = target index, col = match index), duplicates allowed.
static void match_step(const SliceObjs& target, const SliceObjs& match, std::vector<Entry>& entries,
                       double iou_thr, float ioa_thr, bool assign_new, int& next_label,
                       std::vector<int>& new_labels, Scratch& S) {
  const int n = target.size(), m = match.size();
  new_labels.assign(m, 0);
  S.matched_row.assign(m, -1);
  // aggregate duplicate (row, col)
  std::sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& b) {
    return a.row != b.row ? a.row < b.row : a.col < b.col;
  });
  std::vector<Entry>& agg = S.agg;
  agg.clear();
  for (const Entry& e : entries) {
    if (!agg.empty() && agg.back().row == e.row && agg.back().col == e.col) agg.back().inter += e.inter;
    else agg.push_back(e);
  }
  S.iou.resize(agg.size());
  S.ioa.resize(agg.size());
  for (size_t k = 0; k < agg.size(); ++k) {
    const long long inter = agg[k].inter;
    const long long uni = target.area[agg[k].row] + match.area[agg[k].col] - inter;
    S.iou[k] = static_cast<double>(inter) / static_cast<double>(uni);
    S.ioa[k] = static_cast<float>(static_cast<double>(inter) / static_cast<double>(match.area[agg[k].col]));
  }
  if (n > 0 && m > 0 && !agg.empty()) {
    // connected blocks of the bipartite graph (union-find over rows [0,n) and cols [n,n+m))
    std::vector<int>& parent = S.parent;
    parent.resize(n + m);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    for (const Entry& e : agg) {
      const int a = find(e.row), b = find(n + e.col);
      if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }
    std::vector<size_t>& eo = S.eorder;
    eo.resize(agg.size());
    std::iota(eo.begin(), eo.end(), 0);
    std::stable_sort(eo.begin(), eo.end(), [&](size_t a, size_t b) { return find(agg[a].row) < find(agg[b].row); });
    size_t g0 = 0;
    while (g0 < eo.size()) {
      const int root = find(agg[eo[g0]].row);
      size_t g1 = g0 + 1;
      while (g1 < eo.size() && find(agg[eo[g1]].row) == root) ++g1;
      if (g1 - g0 == 1) {  // isolated positive entry: always part of the optimum
        const Entry& e = agg[eo[g0]];
        if (S.iou[eo[g0]] >= iou_thr) S.matched_row[e.col] = e.row;
      } else {
        std::vector<int>&rows = S.rows, &cols = S.cols;
        rows.clear(); cols.clear();
        for (size_t k = g0; k < g1; ++k) { rows.push_back(agg[eo[k]].row); cols.push_back(agg[eo[k]].col); }
        std::sort(rows.begin(), rows.end()); rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
        std::sort(cols.begin(), cols.end()); cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        const int br = static_cast<int>(rows.size()), bc = static_cast<int>(cols.size());
        S.dense.assign(static_cast<size_t>(br) * bc, 0.0);
        for (size_t k = g0; k < g1; ++k) {
          const int r = static_cast<int>(std::lower_bound(rows.begin(), rows.end(), agg[eo[k]].row) - rows.begin());
          const int c = static_cast<int>(std::lower_bound(cols.begin(), cols.end(), agg[eo[k]].col) - cols.begin());
          S.dense[static_cast<size_t>(r) * bc + c] = S.iou[eo[k]];
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
        lap_min(nr, nc, S.cost, S.col4row);
        for (int a = 0; a < nr; ++a) {
          if (S.col4row[a] < 0) continue;
          const int r = transpose ? S.col4row[a] : a, c = transpose ? a : S.col4row[a];
          if (S.dense[static_cast<size_t>(r) * bc + c] >= iou_thr) S.matched_row[cols[c]] = rows[r];
        }
      }
      g0 = g1;
    }
  }
  // per-column IoA maximum (float32 matrix semantics: zeros everywhere else, first max row wins)
  S.ioa_max.assign(m, 0.0f);
  S.ioa_arg.assign(m, 0);
  if (n > 0 && m > 0) {
    for (size_t k = 0; k < agg.size(); ++k) {  // agg is sorted by row, so ">" keeps the first max
      const int c = agg[k].col;
      if (S.ioa[k] > S.ioa_max[c]) { S.ioa_max[c] = S.ioa[k]; S.ioa_arg[c] = agg[k].row; }
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

// Objects with equal new label are merged; output order = first appearance (dict order).
static void merge_by_label(const SliceObjs& match, const std::vector<int>& new_labels, SliceObjs& out,
                           Scratch& S) {
  const int m = match.size();
  out.clear();
  std::vector<int>& order = S.order;
  order.resize(m);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return new_labels[a] < new_labels[b]; });
  S.groups.clear();
  for (int g0 = 0; g0 < m;) {
    int g1 = g0 + 1;
    while (g1 < m && new_labels[order[g1]] == new_labels[order[g0]]) ++g1;
    S.groups.emplace_back(order[g0], g0);  // stable sort: order[g0] is the first appearance
    g0 = g1;
  }
  std::sort(S.groups.begin(), S.groups.end());
  for (const auto& gr : S.groups) {
    const int lab = new_labels[gr.first];
    long long area = 0;
    int b0 = 0x7fffffff, b1 = 0x7fffffff, b2 = -1, b3 = -1;
    for (int k = gr.second; k < m && new_labels[order[k]] == lab; ++k) {
      const int i = order[k];
      area += match.area[i];
      b0 = std::min(b0, match.box[4 * i]); b1 = std::min(b1, match.box[4 * i + 1]);
      b2 = std::max(b2, match.box[4 * i + 2]); b3 = std::max(b3, match.box[4 * i + 3]);
      out.cc.insert(out.cc.end(), match.cc.begin() + match.cc_off[i], match.cc.begin() + match.cc_off[i + 1]);
    }
    out.label.push_back(lab);
    out.area.push_back(area);
    out.box.push_back(b0); out.box.push_back(b1); out.box.push_back(b2); out.box.push_back(b3);
    out.cc_off.push_back(static_cast<int>(out.cc.size()));
  }
}

}  // namespace

extern "C" {

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
  if (n_slices <= 0) { *n_inst = 0; return 0; }
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
  std::unordered_map<int, int> inst_pos;
  inst_pos.reserve(4096);
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
      auto it = inst_pos.find(label);
      if (it == inst_pos.end()) {
        if (ninst >= max_inst) return be_set_error("instance table capacity exceeded; re-run with a larger max_inst");
        inst_pos[label] = ninst;
        inst_labels[ninst] = label;
        inst_sizes[ninst] = bwd_cur.area[i];
        std::memcpy(inst_boxes + ninst * 6, b3, sizeof(b3));
        ++ninst;
      } else {
        const int p = it->second;
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
